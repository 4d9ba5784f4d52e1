// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// extern "C" driver linked together with the UNMODIFIED reference sources
// (/root/reference/src/{util,csr,csr_hw,csr_hw_wrapper,spmv}.cpp, compiled where
// they lie by oracle/Makefile with -DCU -DVF -DDOUBLE, outputs only in
// oracle/_ref/).  It calls the reference API in the order main() does
// (src/main.cpp:46-96) and exposes the resulting hw_matrix pieces, bitmap,
// hw_x and y so that tests can compare the oracle restatement and the CUDA
// engine against the real reference.  One shared object per (CU,VF,DOUBLE).
#include <unistd.h>
#include <fcntl.h>
#include <cstdint>
#include <cstring>
#include <exception>

#include "util.h"
#include "csr.h"
#include "csr_hw.h"
#include "csr_hw_wrapper.h"

namespace {
struct Quiet {  // the reference printf()s timings from inside its API
  int saved;
  Quiet() {
    fflush(stdout); std::cout.flush();
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1); close(nul);
  }
  ~Quiet() { fflush(stdout); std::cout.flush(); dup2(saved, 1); close(saved); }
};
struct RefHandle {
  csr_matrix *m;
  csr_hw_matrix **hw;
  bool **bitmap;
  csr_hw_vector *hw_x;
};
}  // namespace

extern "C" {

int ref_cu() { return ComputeUnits; }
int ref_vf() { return VectFactor; }
int ref_double() { return DOUBLE; }
int ref_cols_div_blocks() { return COLS_DIV_BLOCKS; }
int ref_value_bytes() { return (int)sizeof(ValueType); }

// values: ValueType[nnz] (double or float according to the build)
void *ref_build(uint32_t rows, uint32_t cols, uint32_t nnz, const uint32_t *row_ptr,
                const uint32_t *col_ind, const void *values) {
  Quiet q;
  csr_header hdr;
  hdr.nr_rows = rows; hdr.nr_cols = cols; hdr.nr_nzeros = nnz; hdr.blocks = 0;
  RefHandle *h = new RefHandle();
  h->m = create_csr_matrix(hdr);
  for (uint32_t i = 0; i <= rows; i++) h->m->row_ptr[i] = row_ptr[i];
  for (uint32_t i = 0; i < nnz; i++) h->m->col_ind[i] = col_ind[i];
  memcpy(h->m->values, values, (size_t)nnz * sizeof(ValueType));
  h->hw_x = NULL;
  try {
    create_csr_hw_matrix(h->m, &h->hw, &h->bitmap);
  } catch (const std::exception &e) {
    fprintf(stderr, "ref_build: %s\n", e.what());
    return NULL;
  }
  return h;
}

int ref_blocks(void *hv) { return ((RefHandle *)hv)->hw[0]->blocks; }

// out[5] = nr_rows, nr_cols, nr_nzeros, nr_ci, nr_val of piece (cu, block)
void ref_piece_info(void *hv, int cu, int block, uint32_t *out) {
  csr_hw_matrix *p = ((RefHandle *)hv)->hw[cu];
  out[0] = p->nr_rows[block]; out[1] = p->nr_cols[block]; out[2] = p->nr_nzeros[block];
  out[3] = p->nr_ci[block];   out[4] = p->nr_val[block];
}
const void *ref_piece_words(void *hv, int cu, int block) {
  return ((RefHandle *)hv)->hw[cu]->submatrix[block];
}
const uint8_t *ref_bitmap_row(void *hv, int block) {
  return (const uint8_t *)((RefHandle *)hv)->bitmap[block];
}

// Packs x like src/main.cpp:69; copies block b's words to out (nr_values/RATIO_v words)
int ref_make_hw_x(void *hv, const void *x, uint32_t n) {
  RefHandle *h = (RefHandle *)hv;
  Quiet q;
  csr_vector xv; xv.values = (ValueType *)x; xv.nr_values = n;
  if (h->hw_x) delete_csr_hw_x_vector(h->hw_x);
  create_csr_hw_x_vector(&h->hw_x, &xv, h->hw[0]->blocks, h->hw[0]->nr_cols);
  return 0;
}
uint32_t ref_hw_x_nr_values(void *hv, int block) { return ((RefHandle *)hv)->hw_x->nr_values[block]; }
const void *ref_hw_x_words(void *hv, int block) { return ((RefHandle *)hv)->hw_x->values[block]; }

// y (ValueType[rows]) is accumulated into, exactly like spmv_hw (src/csr_hw_wrapper.cpp:277-281).
// returns 0 ok, 1 if the emulated FIFOs under-ran
int ref_spmv_hw(void *hv, void *y, uint32_t rows) {
  RefHandle *h = (RefHandle *)hv;
  Quiet q;
  csr_vector yv; yv.values = (ValueType *)y; yv.nr_values = rows;
  try {
    spmv_hw(h->hw, h->hw_x, &yv, h->bitmap);
  } catch (const std::exception &e) {
    fprintf(stderr, "ref_spmv_hw: %s\n", e.what());
    return 1;
  }
  return 0;
}

void ref_spmv_gold_h(void *hv, const void *x, void *y) {
  spmv_gold(((RefHandle *)hv)->m, (ValueType *)x, (ValueType *)y);
}

// stand-alone gold on caller arrays (CPU baseline timing; src/csr.cpp:184-194)
void ref_spmv_gold(uint32_t rows, uint32_t cols, uint32_t nnz, uint32_t *row_ptr, uint32_t *col_ind,
                   void *values, const void *x, void *y) {
  csr_matrix m;
  m.row_ptr = (IndexType *)row_ptr; m.col_ind = (IndexType *)col_ind; m.values = (ValueType *)values;
  m.nr_rows = rows; m.nr_cols = cols; m.nr_nzeros = nnz; m.Filename = NULL;
  spmv_gold(&m, (ValueType *)x, (ValueType *)y);
}

int ref_verification(uint32_t n, void *sw, void *hw) {
  Quiet q;
  return verification(n, (ValueType *)sw, (ValueType *)hw, 0);
}

double ref_storage_overhead_mb(void *hv) {
  RefHandle *h = (RefHandle *)hv;
  double mem = 0;
  for (int i = 0; i < ComputeUnits; i++) mem += storage_overhead(h->hw[i]);
  return mem;
}

// Reads a matrix file with the reference reader (src/csr.cpp:10-46,87-136); returns handle without layout
int ref_read_matrix_file(const char *path, uint32_t *hdr_out /*rows, cols, nnz, blocks*/, uint32_t *row_ptr,
                         uint32_t *col_ind, void *values) {
  Quiet q;
  csr_header hdr;
  int rc = read_csr_header(&hdr, (char *)path);
  if (rc) return rc;
  hdr_out[0] = hdr.nr_rows; hdr_out[1] = hdr.nr_cols; hdr_out[2] = hdr.nr_nzeros; hdr_out[3] = hdr.blocks;
  if (!row_ptr) return 0;
  csr_matrix *m = create_csr_matrix(hdr);
  rc = read_csr_matrix(m, (char *)path);
  if (rc) return 10 + rc;
  for (uint32_t i = 0; i <= hdr_out[0]; i++) row_ptr[i] = m->row_ptr[i];
  for (uint32_t i = 0; i < hdr_out[2]; i++) col_ind[i] = m->col_ind[i];
  memcpy(values, m->values, (size_t)hdr_out[2] * sizeof(ValueType));
  delete_csr_matrix(m);
  return 0;
}

void ref_free(void *hv) {
  RefHandle *h = (RefHandle *)hv;
  Quiet q;
  if (h->hw_x) delete_csr_hw_x_vector(h->hw_x);
  int blocks = h->hw[0]->blocks;
  delete_csr_hw_matrix(h->hw);
  for (int b = 0; b < blocks; b++) free(h->bitmap[b]);  // the reference leaks these (src/main.cpp:95)
  free(h->bitmap);
  delete_csr_matrix(h->m);
  delete h;
}

}  // extern "C"
