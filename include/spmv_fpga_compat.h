/*
 * spmv_fpga_compat.h - the C++ host API of euroexa/spmv-fpga on top of the B200 engine (include/spmvb.h).
 *
 * A caller written against the reference (its src/main.cpp:46-96) compiles unchanged against this header with the
 * same -DCU -DVF -DDOUBLE macros and links libspmvb.so instead of the sds++-generated hardware function:
 *
 *     create_csr_hw_matrix(matrix, &hw_matrix, &empty_rows_bitmap);             // csr_hw_wrapper.h:9
 *     create_csr_hw_x_vector(&hw_x, x, hw_matrix[0]->blocks, hw_matrix[0]->nr_cols);   // :11
 *     spmv_hw(hw_matrix, hw_x, y_fpga, empty_rows_bitmap);                      // :13
 *     delete_csr_hw_matrix(hw_matrix); free(empty_rows_bitmap); delete_csr_hw_x_vector(hw_x);   // :15-17
 *
 * Same names, argument meaning and ownership as the reference: the callee allocates, the caller frees with delete_*;
 * y_fpga is caller-zeroed and ACCUMULATED into (csr_hw.cpp:1557); functions are void and print diagnostics
 * (a CUDA failure prints the message and aborts: there is no CPU fallback).  Types are plain C++ instead of the Xilinx
 * ap_uint<> (util.h:9-16,69): IndexType = uint32_t, BusDataType = 16-byte POD, so struct layouts are the same bytes.
 *
 * Header-only (the reference configures itself with compile-time macros, so does this shim); all state lives in the
 * handles: hw_matrix[0] carries the engine.  Not thread-safe per handle, like the reference.
 */
#ifndef SPMV_FPGA_COMPAT_H
#define SPMV_FPGA_COMPAT_H

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "spmvb.h"

#ifndef CU
#define CU 1
#endif
#ifndef VF
#define VF 1
#endif
#ifndef DOUBLE
#define DOUBLE 1
#endif

/* ---- util.h:9-76 */
typedef uint32_t IndexType;
typedef uint16_t CompressedIndexType;
#if DOUBLE == 0
typedef float ValueType;
#define VALUE_TYPE_BIT_WIDTH 32
#else
typedef double ValueType;
#define VALUE_TYPE_BIT_WIDTH 64
#endif
#define INDEX_TYPE_BIT_WIDTH 32
#define COMPRESSED_INDEX_TYPE_BIT_WIDTH 16
#define VectFactor VF
#define ComputeUnits CU
#if CU == 10 || CU == 12
#define COLS_DIV_BLOCKS (16384)
#else
#define COLS_DIV_BLOCKS (32768)
#endif
#define BUS_BIT_WIDTH 128
#define RATIO_v (BUS_BIT_WIDTH / VALUE_TYPE_BIT_WIDTH)
#define RATIO_ci (BUS_BIT_WIDTH / COMPRESSED_INDEX_TYPE_BIT_WIDTH)
#define RATIO_col_val (RATIO_ci / RATIO_v + 1)
typedef struct BusDataType { uint64_t limb[2]; } BusDataType; /* ap_uint<128>: two little-endian 64-bit halves */

static inline double getTimestamp() { /* util.cpp:3-7: microseconds */
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_usec + tv.tv_sec * 1e6;
}

/* ---- csr.h:7-29 */
typedef struct csr_header { IndexType nr_rows, nr_cols, nr_nzeros; int blocks; } csr_header;
typedef struct csr_matrix {
  IndexType *row_ptr, *col_ind;
  ValueType *values;
  IndexType nr_nzeros, nr_rows, nr_cols;
  char *Filename;
} csr_matrix;
typedef struct csr_vector { ValueType *values; IndexType nr_values; } csr_vector;

/* ---- csr_hw.h:16-33 */
typedef struct csr_hw_matrix {
  BusDataType **submatrix;
  IndexType *nr_rows, *nr_cols, *nr_nzeros, *nr_ci, *nr_val;
  int blocks;
} csr_hw_matrix;
typedef struct csr_hw_vector { BusDataType **values; IndexType *nr_values; int blocks; } csr_hw_vector;

/* hw_matrix[0] is really one of these; the public part comes first */
typedef struct spmvb_compat_owner {
  csr_hw_matrix pub;
  uint32_t magic;
  spmvb_layout *layout;
  spmvb_engine *engine; /* one GPU ... */
  spmvb_group *group;   /* ... or several (SPMVB_DEVICES): the compute units mapped to GPUs */
  IndexType rows, cols;
  ValueType *x_scratch; /* expanded_nr_cols values */
} spmvb_compat_owner;
#define SPMVB_COMPAT_MAGIC 0x53504d56u

static inline void spmvb_compat_die(const char *where) {
  fprintf(stderr, "%s: %s\n", where, spmvb_last_error());
  abort();
}

/* ---- csr.cpp:10-46 */
static inline int read_csr_header(csr_header *hdr, char *Filename) {
  FILE *fp = fopen(Filename, "r");
  if (!fp) { printf("Could not open file %s\n", Filename); return 1; }
  unsigned r, c, n;
  int matched = fscanf(fp, "%u %u %u\n", &r, &c, &n);
  fclose(fp);
  if (matched == EOF) { printf("unexpected eof found\n"); return 1; }
  if (matched != 3) { printf("parse error\n"); return 3; }
  hdr->nr_rows = r; hdr->nr_cols = c; hdr->nr_nzeros = n;
  hdr->blocks = (int)(c / COLS_DIV_BLOCKS) + ((c % COLS_DIV_BLOCKS) ? 1 : 0);
  return 0;
}
/* ---- csr.cpp:51-80 */
static inline csr_matrix *create_csr_matrix(csr_header hdr) {
  csr_matrix *m = (csr_matrix *)malloc(sizeof(csr_matrix));
  m->nr_rows = hdr.nr_rows; m->nr_cols = hdr.nr_cols; m->nr_nzeros = hdr.nr_nzeros;
  m->row_ptr = (IndexType *)malloc(((size_t)hdr.nr_rows + 1) * sizeof(IndexType));
  m->col_ind = (IndexType *)malloc((size_t)hdr.nr_nzeros * sizeof(IndexType));
  m->values = (ValueType *)malloc((size_t)hdr.nr_nzeros * sizeof(ValueType));
  m->Filename = NULL;
  return m;
}
static inline void delete_csr_matrix(csr_matrix *m) {
  if (!m) return;
  free(m->values); free(m->col_ind); free(m->row_ptr); free(m);
}
/* ---- csr.cpp:87-136 (reads through the library's parser; trailing empty rows get row_ptr = nnz) */
static inline int read_csr_matrix(csr_matrix *m, char *Filename) {
  spmvb_csr *A = NULL;
  /* SPMVB_CSR_CACHE=1: keep the parsed matrix in a binary sidecar next to the file and reuse it while it is current */
  const char *cache = getenv("SPMVB_CSR_CACHE");
  const int rc = (cache && atoi(cache)) ? spmvb_csr_read_cached(Filename, DOUBLE, &A) : spmvb_csr_read(Filename, DOUBLE, &A);
  if (rc != SPMVB_OK) { printf("parse error: %s\n", spmvb_last_error()); return 1; }
  const uint32_t rows = spmvb_csr_rows(A);
  const uint64_t nnz = spmvb_csr_nnz(A);
  if (rows != m->nr_rows || nnz != m->nr_nzeros) { printf("parse error: header mismatch\n"); spmvb_csr_free(A); return 1; }
  const uint64_t *rp = spmvb_csr_row_ptr(A);
  for (uint32_t i = 0; i <= rows; i++) m->row_ptr[i] = (IndexType)rp[i];
  memcpy(m->col_ind, spmvb_csr_col_ind(A), (size_t)nnz * sizeof(IndexType));
  memcpy(m->values, spmvb_csr_values(A), (size_t)nnz * sizeof(ValueType));
  m->nr_cols = spmvb_csr_cols(A);
  m->Filename = Filename;
  spmvb_csr_free(A);
  return 0;
}
/* ---- csr.cpp:141-179 */
static inline csr_vector *create_csr_vector(IndexType nr_values) {
  csr_vector *v = (csr_vector *)malloc(sizeof(csr_vector));
  v->nr_values = nr_values;
  v->values = (ValueType *)calloc(nr_values ? nr_values : 1, sizeof(ValueType));
  return v;
}
static inline void delete_csr_vector(csr_vector *v) {
  if (!v) return;
  free(v->values); free(v);
}
static inline void init_vector_rand(csr_vector *v, ValueType max) {
  if (!v) return;
  for (IndexType i = 0; i < v->nr_values; i++) v->values[i] = max * (rand() / (ValueType)RAND_MAX);
}

/* ---- csr.cpp:184-194 and csr_hw.cpp:1571-1590: the CALLER's CPU self-check (main.cpp:61, :76).  They are here so that
 * a main() written against the reference compiles unchanged; spmv_hw never calls them and nothing in libspmvb.so does. */
static inline void spmv_gold(csr_matrix *matrix, ValueType *x, ValueType *y) {
  for (IndexType row = 0; row < matrix->nr_rows; row++) {
    ValueType acc = 0.0;
    for (IndexType j = matrix->row_ptr[row]; j < matrix->row_ptr[row + 1]; j++) acc += matrix->values[j] * x[matrix->col_ind[j]];
    y[row] = acc;
  }
}
/* verbose: 0 = nothing, 1 = only errors, 2 = every value; returns 1 if any |sw - hw| >= 1e-5 (or NaN) */
static inline int verification(IndexType nr_values, ValueType *sw_values, ValueType *hw_values, int verbose) {
  const ValueType limit = (ValueType)1e-5;
  IndexType bad = 0;
  for (IndexType i = 0; i < nr_values; i++) {
    const ValueType d = (ValueType)fabs((double)(sw_values[i] - hw_values[i]));
    if (verbose == 2) printf("%u : y_gold = %.14g\ty_hw = %.14g\n", (unsigned)i, (double)sw_values[i], (double)hw_values[i]);
    if (d >= limit || d != d) {
      bad++;
      if (verbose == 1 || verbose == 2)
        printf("\tError occurs at %u : y_gold = %.14g, y_hw = %.14g. Relative difference is %.14g\n", (unsigned)i,
               (double)sw_values[i], (double)hw_values[i], fabs((double)(d / sw_values[i])));
    }
  }
  if (bad) printf("Total errors : %u\n", (unsigned)bad);
  return bad != 0;
}

/* ---- csr_hw_wrapper.cpp:3-80: builds the layout (bit-exact pieces) and uploads it to the GPU */
/* Run-time configuration of a drop-in executable (the reference has compile-time macros only, Makefile:71):
 *   SPMVB_DEVICE=i        the GPU to use (default 0)
 *   SPMVB_DEVICES=a,b,..  several GPUs of this box: the CU dimension mapped to GPUs (spmvb_group_create) - rows are
 *                         split over them by non-zero count, x goes to every GPU, all kernels run concurrently
 *   SPMVB_GPU_BUILD=1     build the hw_matrix layout on the GPU; SPMVB_CSR_CACHE=1 binary sidecar of the matrix file
 *   SPMVB_<OPTION>=n      any spmvb_set_option name (spmvb_options_from_env) */
static inline int spmvb_compat_devices(int *devs, int max) {
  const char *list = getenv("SPMVB_DEVICES");
  int n = 0;
  if (!list) return 0;
  while (*list && n < max) {
    char *end;
    long v = strtol(list, &end, 10);
    if (end == list) break;
    devs[n++] = (int)v;
    list = (*end == ',') ? end + 1 : end;
  }
  return n;
}

static inline void create_csr_hw_matrix(csr_matrix *matrix, csr_hw_matrix ***hw_matrix, bool ***empty_rows_bitmap) {
  spmvb_layout *L = NULL;
  spmvb_engine *E = NULL;
  spmvb_group *G = NULL;
  spmvb_options_from_env();
  const char *dev = getenv("SPMVB_DEVICE");
  const char *gpu_build = getenv("SPMVB_GPU_BUILD");
  int devs[64];
  const int n_devs = spmvb_compat_devices(devs, 64);
  if (n_devs > 1) {
    /* hw_matrix[k] keeps the reference's per-block CU pieces on the host (below); the GPUs get contiguous row ranges of
     * their own, balanced by non-zero count with the same split rule applied to whole rows (SURVEY 8e mapping A) */
    uint64_t *rp64 = (uint64_t *)malloc(((size_t)matrix->nr_rows + 1) * sizeof(uint64_t));
    for (IndexType i = 0; i <= matrix->nr_rows; i++) rp64[i] = matrix->row_ptr[i];
    if (spmvb_group_create(matrix->nr_rows, matrix->nr_cols, rp64, matrix->col_ind, matrix->values, DOUBLE, n_devs, devs, 0,
                           &G) != SPMVB_OK)
      spmvb_compat_die("create_csr_hw_matrix (multi-GPU)");
    free(rp64);
  }
  if (!G && gpu_build && atoi(gpu_build)) {
    /* SPMVB_GPU_BUILD=1: the same layout built by CUDA kernels (needs sorted rows; errors are fatal like any other),
     * then copied back because this API exposes submatrix[b] and the bitmap on the host */
    uint64_t *rp64 = (uint64_t *)malloc(((size_t)matrix->nr_rows + 1) * sizeof(uint64_t));
    for (IndexType i = 0; i <= matrix->nr_rows; i++) rp64[i] = matrix->row_ptr[i];
    if (spmvb_engine_create_from_csr(matrix->nr_rows, matrix->nr_cols, rp64, matrix->col_ind, matrix->values, CU, VF, DOUBLE,
                                     COLS_DIV_BLOCKS, dev ? atoi(dev) : 0, 0, 0, &L, &E) != SPMVB_OK ||
        spmvb_engine_fetch_layout(E, L) != SPMVB_OK)
      spmvb_compat_die("create_csr_hw_matrix (GPU build)");
    free(rp64);
  } else if (spmvb_layout_build_u32(matrix->nr_rows, matrix->nr_cols, matrix->row_ptr, matrix->col_ind, matrix->values, CU,
                                    VF, DOUBLE, COLS_DIV_BLOCKS, &L) != SPMVB_OK) {
    spmvb_compat_die("create_csr_hw_matrix");
  }
  const int blocks = spmvb_layout_blocks(L);
  *hw_matrix = (csr_hw_matrix **)malloc(ComputeUnits * sizeof(csr_hw_matrix *));
  for (int k = 0; k < ComputeUnits; k++) {
    csr_hw_matrix *p;
    if (k == 0) {
      spmvb_compat_owner *o = (spmvb_compat_owner *)calloc(1, sizeof(spmvb_compat_owner));
      o->magic = SPMVB_COMPAT_MAGIC; o->layout = L; o->rows = matrix->nr_rows; o->cols = matrix->nr_cols;
      o->x_scratch = (ValueType *)calloc(spmvb_layout_expanded_cols(L), sizeof(ValueType));
      p = &o->pub;
    } else {
      p = (csr_hw_matrix *)calloc(1, sizeof(csr_hw_matrix));
    }
    p->blocks = blocks;
    p->submatrix = (BusDataType **)malloc(blocks * sizeof(BusDataType *));
    p->nr_rows = (IndexType *)malloc(blocks * sizeof(IndexType));
    p->nr_cols = (IndexType *)malloc(blocks * sizeof(IndexType));
    p->nr_nzeros = (IndexType *)malloc(blocks * sizeof(IndexType));
    p->nr_ci = (IndexType *)malloc(blocks * sizeof(IndexType));
    p->nr_val = (IndexType *)malloc(blocks * sizeof(IndexType));
    for (int b = 0; b < blocks; b++) {
      uint32_t info[5];
      spmvb_layout_piece_info(L, k, b, info);
      p->nr_rows[b] = info[0]; p->nr_cols[b] = info[1]; p->nr_nzeros[b] = info[2]; p->nr_ci[b] = info[3]; p->nr_val[b] = info[4];
      p->submatrix[b] = (BusDataType *)spmvb_layout_piece_words(L, k, b); /* owned by the layout */
      if (info[2] == 0) printf("WARNING !!!!! block %d is empty!\n", b);   /* csr_hw.cpp:171 */
    }
    (*hw_matrix)[k] = p;
  }
  /* empty_rows_bitmap[block][row], one bool per pair like the reference (csr_hw.cpp:391-393) */
  *empty_rows_bitmap = (bool **)malloc(blocks * sizeof(bool *));
  for (int b = 0; b < blocks; b++) {
    (*empty_rows_bitmap)[b] = (bool *)malloc(matrix->nr_rows ? matrix->nr_rows : 1);
    spmvb_layout_bitmap_row(L, b, (uint8_t *)(*empty_rows_bitmap)[b]);
  }
  { /* csr_hw.cpp:420-421 */
    double tot = (double)spmvb_layout_padded_nnz(L);
    double in = tot * (VALUE_TYPE_BIT_WIDTH + COMPRESSED_INDEX_TYPE_BIT_WIDTH) / (8.0 * 1024 * 1024);
    double out = (double)blocks * matrix->nr_rows * VALUE_TYPE_BIT_WIDTH / (8.0 * 1024 * 1024);
    printf("Total non-zeros : %.0f. Total %g MB transferred ( in : %g, out : %g)\n", tot, in + out, in, out);
  }
  spmvb_compat_owner *o = (spmvb_compat_owner *)(*hw_matrix)[0];
  if (!E && !G && spmvb_engine_create(L, dev ? atoi(dev) : 0, 0, &E) != SPMVB_OK) spmvb_compat_die("create_csr_hw_matrix (GPU upload)");
  o->engine = E;
  o->group = G;
}

/* ---- csr_hw.cpp:1436-1488: per-block packed x slices, zero padded */
static inline void create_csr_hw_x_vector(csr_hw_vector **hw_x, csr_vector *x, int blocks, IndexType *nr_cols) {
  csr_hw_vector *v = (csr_hw_vector *)malloc(sizeof(csr_hw_vector));
  v->blocks = blocks;
  v->nr_values = (IndexType *)malloc(blocks * sizeof(IndexType));
  v->values = (BusDataType **)malloc(blocks * sizeof(BusDataType *));
  IndexType cnt = 0;
  for (int b = 0; b < blocks; b++) {
    v->nr_values[b] = nr_cols[b];
    v->values[b] = (BusDataType *)calloc((size_t)(nr_cols[b] / RATIO_v) ? (size_t)(nr_cols[b] / RATIO_v) : 1, sizeof(BusDataType));
    ValueType *dst = (ValueType *)v->values[b];
    for (IndexType j = 0; j < nr_cols[b] / RATIO_v * RATIO_v; j++, cnt++) dst[j] = cnt < x->nr_values ? x->values[cnt] : 0;
  }
  *hw_x = v;
}

/* ---- csr_hw_wrapper.cpp:82-185 (kept for API parity; spmv_hw no longer needs partial-y buffers) */
static inline void create_csr_hw_y_vector(csr_hw_matrix **hw_matrix, csr_hw_vector ***hw_vector) {
  *hw_vector = (csr_hw_vector **)malloc(ComputeUnits * sizeof(csr_hw_vector *));
  for (int k = 0; k < ComputeUnits; k++) {
    csr_hw_vector *v = (csr_hw_vector *)malloc(sizeof(csr_hw_vector));
    v->blocks = hw_matrix[k]->blocks;
    v->nr_values = (IndexType *)malloc(v->blocks * sizeof(IndexType));
    v->values = (BusDataType **)malloc(v->blocks * sizeof(BusDataType *));
    for (int b = 0; b < v->blocks; b++) {
      v->nr_values[b] = hw_matrix[k]->nr_rows[b];
      size_t words = hw_matrix[k]->nr_rows[b] / RATIO_v;
      v->values[b] = (BusDataType *)calloc(words ? words : 1, sizeof(BusDataType));
    }
    (*hw_vector)[k] = v;
  }
}
static inline void spmvb_compat_delete_vector(csr_hw_vector *v) {
  if (!v) return;
  for (int b = 0; b < v->blocks; b++) free(v->values[b]);
  free(v->values); free(v->nr_values); free(v);
}
static inline void delete_csr_hw_y_vector(csr_hw_vector **hw_vector) {
  for (int k = 0; k < ComputeUnits; k++) spmvb_compat_delete_vector(hw_vector[k]);
  free(hw_vector);
}
static inline void delete_csr_hw_x_vector(csr_hw_vector *hw_vector) { spmvb_compat_delete_vector(hw_vector); }

/* ---- csr_hw_wrapper.cpp:193-288: one fused kernel launch over all (CU, block) pieces instead of the per-block spmv()
 * calls and the host accum_results loop; y_fpga is accumulated into */
static inline void spmv_hw(csr_hw_matrix **hw_matrix, csr_hw_vector *hw_x, csr_vector *y_fpga, bool **empty_rows_bitmap) {
  (void)empty_rows_bitmap; /* the device keeps the bitmap in its compact row-map form */
  spmvb_compat_owner *o = (spmvb_compat_owner *)hw_matrix[0];
  if (o->magic != SPMVB_COMPAT_MAGIC) { fprintf(stderr, "spmv_hw: hw_matrix was not made by create_csr_hw_matrix\n"); abort(); }
  /* hw_x is x in block order: concatenate the slices again.  It must be the vector create_csr_hw_x_vector made for
   * THIS matrix (same blocks, nr_values[b] = nr_cols[b], a whole number of bus words each): anything else would overrun
   * x_scratch, which holds expanded_nr_cols values */
  const uint32_t cap = spmvb_layout_expanded_cols(o->layout);
  uint32_t n = 0;
  if (hw_x->blocks != hw_matrix[0]->blocks) { fprintf(stderr, "spmv_hw: hw_x has %d blocks, the matrix %d\n", hw_x->blocks, hw_matrix[0]->blocks); abort(); }
  for (int b = 0; b < hw_x->blocks; b++) {
    const uint32_t nv = hw_x->nr_values[b];
    if (nv % RATIO_v != 0 || nv > cap - n) { fprintf(stderr, "spmv_hw: hw_x block %d does not belong to this matrix\n", b); abort(); }
    memcpy(o->x_scratch + n, hw_x->values[b], (size_t)nv * sizeof(ValueType));
    n += nv;
  }
  double t0 = getTimestamp();
  const int rc = o->group ? spmvb_group_spmv_host(o->group, o->x_scratch, n, y_fpga->values, 1)
                          : spmvb_engine_spmv_host(o->engine, o->x_scratch, n, y_fpga->values, 1);
  if (rc != SPMVB_OK) spmvb_compat_die("spmv_hw");
  double t1 = getTimestamp();
  printf("Hardware execution time : %.6f ms elapsed\n", (t1 - t0) / 1000);
  printf("Result accumulation time : %.6f ms elapsed\n", 0.0); /* fused on the device */
  printf("Total time  : %.6f ms elapsed\n", (t1 - t0) / 1000);
}

/* ---- csr_hw_wrapper.cpp:291-296 */
static inline void delete_csr_hw_matrix(csr_hw_matrix **hw_matrix) {
  spmvb_compat_owner *o = (spmvb_compat_owner *)hw_matrix[0];
  if (o->engine) spmvb_engine_free(o->engine);
  if (o->group) spmvb_group_free(o->group);
  spmvb_layout *L = o->layout;
  for (int k = 0; k < ComputeUnits; k++) {
    csr_hw_matrix *p = hw_matrix[k];
    free(p->submatrix); free(p->nr_rows); free(p->nr_cols); free(p->nr_nzeros); free(p->nr_ci); free(p->nr_val);
    if (k == 0) free(o->x_scratch);
    free(p);
  }
  spmvb_layout_free(L);
  free(hw_matrix);
}
/* not in the reference (its main() leaks the rows, main.cpp:95): frees the rows AND the outer array */
static inline void delete_empty_rows_bitmap(bool **bitmap, int blocks) {
  for (int b = 0; b < blocks; b++) free(bitmap[b]);
  free(bitmap);
}
/* ---- csr_hw.cpp:1493-1521: debug print of bus words, flag 0 = as RATIO_v values, otherwise as 8 x (15-bit column
 * index <end-of-row bit>); same text as the reference's std::cout << std::setprecision(14) */
static inline void print_wide(BusDataType *values, IndexType nr_values, int flag) {
  for (IndexType j = 0; j < nr_values; j++) {
    const unsigned char *w = (const unsigned char *)&values[j];
    if (flag == 0) {
      for (int k = 0; k < RATIO_v; k++) {
        ValueType v;
        memcpy(&v, w + k * sizeof(ValueType), sizeof(ValueType));
        printf("| (%d : %d) = %.14g ", VALUE_TYPE_BIT_WIDTH * (k + 1) - 1, VALUE_TYPE_BIT_WIDTH * k, (double)v);
      }
    } else {
      for (int k = 0; k < RATIO_ci; k++) {
        uint16_t ci;
        memcpy(&ci, w + 2 * k, 2);
        printf("| (%d : %d) = %u <%u>\t", COMPRESSED_INDEX_TYPE_BIT_WIDTH * (k + 1) - 1, COMPRESSED_INDEX_TYPE_BIT_WIDTH * k,
               (unsigned)(ci & 0x7FFFu), (unsigned)(ci >> 15));
      }
    }
    printf("|\n");
  }
}

/* ---- csr_hw.cpp:1401-1409 */
static inline ValueType storage_overhead(csr_hw_matrix *matrix) {
  double bits = (double)matrix->blocks * 5 * INDEX_TYPE_BIT_WIDTH;
  for (int i = 0; i < matrix->blocks; i++) bits += ((double)matrix->nr_ci[i] + matrix->nr_val[i]) * BUS_BIT_WIDTH;
  return (ValueType)(bits / (8.0 * 1024 * 1024));
}

#endif /* SPMV_FPGA_COMPAT_H */
