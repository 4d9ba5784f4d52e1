// run.elf analogue: the call sequence of the reference's src/main.cpp:16-99 against the B200 engine, through the
// reference's own host API (include/spmv_fpga_compat.h).  Build with the reference's macros: -DCU= -DVF= -DDOUBLE=.
//   run_cu<C>_vf<V>_d<D>.elf <matrix-file>
// spmv_gold and verification are the caller's self-check from the drop-in header (the reference's main() does the same
// with csr.cpp:184-194 and csr_hw.cpp:1571-1590); spmv_hw itself never touches them.
#include <math.h>

#include <iostream>
#include <string>

#include "spmv_fpga_compat.h"

int main(int argc, char **argv) {
  std::cout << "Welcome to SpMV (Compute Units : " << ComputeUnits << ", Vectorization Factor : " << VectFactor << ", "
            << (DOUBLE ? "double" : "single") << "-precision arithmetic)\n";
  if (argc != 2) {
    printf("please enter the input file name  \n");
    return 1;
  }
  csr_header hdr;
  if (read_csr_header(&hdr, argv[1]) != 0) { std::cout << "Error reading matrix header\n"; return 1; }
  csr_matrix *matrix = create_csr_matrix(hdr);
  if (read_csr_matrix(matrix, argv[1]) != 0) { std::cout << "Error reading matrix\n"; return 1; }
  csr_vector *x = create_csr_vector(hdr.nr_cols);
  init_vector_rand(x, 1);
  csr_vector *y = create_csr_vector(hdr.nr_rows);
  double t0 = getTimestamp();
  spmv_gold(matrix, x->values, y->values);
  printf("Software execution time : %.6f ms elapsed\n", (getTimestamp() - t0) / 1000);

  bool **empty_rows_bitmap;
  csr_hw_matrix **hw_matrix;
  csr_hw_vector *hw_x;
  t0 = getTimestamp();
  create_csr_hw_matrix(matrix, &hw_matrix, &empty_rows_bitmap);
  create_csr_hw_x_vector(&hw_x, x, hw_matrix[0]->blocks, hw_matrix[0]->nr_cols);
  printf("Matrix read time        : %.6f ms elapsed\n", (getTimestamp() - t0) / 1000);

  csr_vector *y_fpga = create_csr_vector(hdr.nr_rows);
  spmv_hw(hw_matrix, hw_x, y_fpga, empty_rows_bitmap);
  int status = verification(y->nr_values, y->values, y_fpga->values, 0);
  std::cout << (status == 0 ? "Verification PASSED!\n" : "Verification FAILED!\n");

  double csr_mem = (((double)matrix->nr_rows + 1) * INDEX_TYPE_BIT_WIDTH +
                    (double)matrix->nr_nzeros * (INDEX_TYPE_BIT_WIDTH + VALUE_TYPE_BIT_WIDTH)) / (8.0 * 1024 * 1024);
  double mem = 0;
  for (int i = 0; i < ComputeUnits; i++) mem += storage_overhead(hw_matrix[i]);
  std::cout << "CSR representation : " << csr_mem << " MB. Our representation : " << mem
            << " MB. Storage Overhead : " << (mem - csr_mem) / csr_mem * 100 << " %\n";

  const int blocks = hw_matrix[0]->blocks;
  delete_csr_matrix(matrix);
  delete_csr_vector(x);
  delete_csr_vector(y);
  delete_csr_hw_matrix(hw_matrix);
  delete_empty_rows_bitmap(empty_rows_bitmap, blocks);
  delete_csr_hw_x_vector(hw_x);
  delete_csr_vector(y_fpga);
  return status;
}
