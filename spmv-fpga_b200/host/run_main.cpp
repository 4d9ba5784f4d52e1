// Command-line driver over the drop-in host API (include/spmv_fpga_compat.h): what an existing user script expects from
// the reference's run.elf (src/main.cpp:16-99) - same argument, same report lines, same exit status - with the SpMV on
// the B200 engine.  Build with the reference's macros: -DCU= -DVF= -DDOUBLE=.
//   run_cu<C>_vf<V>_d<D>.elf <matrix-file>
// The gold SpMV and the comparison are the caller's self-check (csr.cpp:184-194, csr_hw.cpp:1571-1590 in the
// reference); spmv_hw never touches them.
#include <cstdio>
#include <iostream>

#include "spmv_fpga_compat.h"

namespace {

struct Stopwatch {  // the reference reports milliseconds from its microsecond timestamps
  double start = getTimestamp();
  double ms() const { return (getTimestamp() - start) / 1000; }
};

struct Problem {
  csr_matrix *A = nullptr;
  csr_vector *x = nullptr, *y_gold = nullptr, *y_engine = nullptr;
  ~Problem() {
    if (A) delete_csr_matrix(A);
    for (csr_vector *v : {x, y_gold, y_engine})
      if (v) delete_csr_vector(v);
  }
};

// matrix file -> CSR, x = the reference's seeded random vector, two result vectors; nullptr message on success
const char *load(char *path, Problem &p) {  // char *: the reference's reader signature (csr.h)
  csr_header h;
  if (read_csr_header(&h, path) != 0) return "Error reading matrix header\n";
  p.A = create_csr_matrix(h);
  if (read_csr_matrix(p.A, path) != 0) return "Error reading matrix\n";
  p.x = create_csr_vector(h.nr_cols);
  init_vector_rand(p.x, 1);
  p.y_gold = create_csr_vector(h.nr_rows);
  p.y_engine = create_csr_vector(h.nr_rows);
  return nullptr;
}

struct Engine {  // the three handles of the reference API and their release order
  csr_hw_matrix **pieces = nullptr;
  bool **bitmap = nullptr;
  csr_hw_vector *x = nullptr;
  int blocks = 0;
  void build(const Problem &p) {
    create_csr_hw_matrix(p.A, &pieces, &bitmap);
    blocks = pieces[0]->blocks;
    create_csr_hw_x_vector(&x, p.x, blocks, pieces[0]->nr_cols);
  }
  double megabytes() const {
    double mb = 0;
    for (int cu = 0; cu < ComputeUnits; cu++) mb += storage_overhead(pieces[cu]);
    return mb;
  }
  ~Engine() {
    if (pieces) delete_csr_hw_matrix(pieces);
    if (bitmap) delete_empty_rows_bitmap(bitmap, blocks);
    if (x) delete_csr_hw_x_vector(x);
  }
};

double csr_megabytes(const csr_matrix *A) {
  const double bits = ((double)A->nr_rows + 1) * INDEX_TYPE_BIT_WIDTH +
                      (double)A->nr_nzeros * (INDEX_TYPE_BIT_WIDTH + VALUE_TYPE_BIT_WIDTH);
  return bits / (8.0 * 1024 * 1024);
}

}  // namespace

int main(int argc, char **argv) {
  std::cout << "Welcome to SpMV (Compute Units : " << ComputeUnits << ", Vectorization Factor : " << VectFactor << ", "
            << (DOUBLE ? "double" : "single") << "-precision arithmetic)\n";
  if (argc != 2) {
    printf("please enter the input file name  \n");
    return 1;
  }
  Problem p;
  if (const char *err = load(argv[1], p)) {
    std::cout << err;
    return 1;
  }
  {
    Stopwatch t;
    spmv_gold(p.A, p.x->values, p.y_gold->values);
    printf("Software execution time : %.6f ms elapsed\n", t.ms());
  }
  Engine e;
  {
    Stopwatch t;
    e.build(p);
    printf("Matrix read time        : %.6f ms elapsed\n", t.ms());
  }
  spmv_hw(e.pieces, e.x, p.y_engine, e.bitmap);
  const int status = verification(p.y_gold->nr_values, p.y_gold->values, p.y_engine->values, 0);
  std::cout << (status == 0 ? "Verification PASSED!\n" : "Verification FAILED!\n");

  const double csr_mb = csr_megabytes(p.A), ours_mb = e.megabytes();
  std::cout << "CSR representation : " << csr_mb << " MB. Our representation : " << ours_mb
            << " MB. Storage Overhead : " << (ours_mb - csr_mb) / csr_mb * 100 << " %\n";
  return status;
}
