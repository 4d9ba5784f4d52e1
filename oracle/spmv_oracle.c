/* TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path; see spmv_oracle.h.
 *
 * CPU restatement of the reference algorithm, one plain-C function per reference
 * step, each citing the reference file:line it follows (/root/reference/src/...).
 * Run-time (cu, vf, is_double) replace the reference's -DCU -DVF -DDOUBLE macros.
 * It deliberately keeps the reference's O(blocks x rows) tables: it is only used
 * at sizes where that is fine.  Differences from the reference, all confined to
 * inputs where the reference itself is undefined:
 *   Q1  buffers are sized with ceil(nnz/RATIO_v) value words (nr_val still floors)
 *   Q2  compute units whose split never fired get 0 rows / 0 non-zeros
 *   Q4  the bitmap walk in the accumulation is bounded by the row count
 *   Q5  unused 16-bit index slots are zero
 * PARITY PINNED against oracle/_ref (the compiled reference) by tests/test_oracle.py
 * and against tests/golden/ fixtures generated from it.
 */
#include "spmv_oracle.h"

#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdlib.h>
#include <string.h>

#define BUS_BYTES 16
#define RATIO_CI 8 /* src/util.h:65 */

struct orc_layout {
  int cu, vf, is_double, blocks;
  uint32_t rows, cols, expanded_cols, cdb;
  int ratio_v;       /* src/util.h:64 */
  int ratio_col_val; /* src/util.h:67 */
  uint32_t *thres_l, *thres_h;
  uint8_t **bitmap;                                     /* [block][row], src/csr_hw.cpp:391-393 */
  uint32_t **nr_rows, **nr_nzeros, **nr_ci, **nr_val;   /* [cu][block] */
  uint32_t *nr_cols;                                    /* [block] */
  uint8_t ***words;                                     /* [cu][block] -> bytes */
};

static uint32_t default_cdb(int cu) { return (cu == 10 || cu == 12) ? 16384u : 32768u; } /* src/util.h:41-59 */

static int vbytes(const orc_layout *l) { return l->is_double ? 8 : 4; }

/* src/csr_hw.cpp:270-318 generate_balanced_hw_submatrix: pack `n` entries into 128-bit words.
 * entry e -> index slot (e%8) of word (e/8)*ratio_col_val; value lane (s%ratio_v) of word
 * g*ratio_col_val + 1 + s/ratio_v.  Little-endian ap_uint<128> => byte addressing below. */
static void pack_piece(const orc_layout *l, uint8_t *w, const uint32_t *col, const uint8_t *eor, const void *val,
                       uint32_t first, uint32_t n) {
  const int vb = vbytes(l);
  for (uint32_t e = 0; e < n; e++) {
    uint32_t g = e / RATIO_CI, s = e % RATIO_CI;
    uint8_t *grp = w + (size_t)g * l->ratio_col_val * BUS_BYTES;
    uint16_t ci = (uint16_t)((col[first + e] & 0x7FFFu) | (eor[first + e] ? 0x8000u : 0));
    memcpy(grp + 2 * s, &ci, 2);
    memcpy(grp + BUS_BYTES + (size_t)s * vb, (const uint8_t *)val + (size_t)(first + e) * vb, vb);
  }
}

orc_layout *orc_layout_build(uint32_t rows, uint32_t cols, const uint64_t *row_ptr, const uint32_t *col_ind,
                             const void *values, int cu, int vf, int is_double, uint32_t cols_div_blocks) {
  orc_layout *l = (orc_layout *)calloc(1, sizeof(*l));
  l->cu = cu; l->vf = vf; l->is_double = is_double;
  l->rows = rows; l->cols = cols;
  l->cdb = cols_div_blocks ? cols_div_blocks : default_cdb(cu);
  l->ratio_v = is_double ? 2 : 4;
  l->ratio_col_val = RATIO_CI / l->ratio_v + 1;
  const int vb = vbytes(l);

  /* --- scan_matrix, src/csr_hw.cpp:25-76: blocks, expanded_nr_cols, thresholds --- */
  int blocks = (int)(cols / l->cdb) + 1;
  if (cols % l->cdb == 0) blocks--;
  l->blocks = blocks;
  uint32_t N = (uint32_t)l->ratio_v * (uint32_t)blocks;
  l->expanded_cols = cols;
  if (cols % N != 0) l->expanded_cols += N - cols % N;
  l->thres_l = (uint32_t *)calloc(blocks, 4);
  l->thres_h = (uint32_t *)calloc(blocks, 4);
  l->thres_l[0] = 0;
  l->thres_h[0] = (blocks == 1) ? l->expanded_cols - 1 : l->cdb - 1;
  for (int b = 1; b < blocks; b++) {
    l->thres_l[b] = l->thres_h[b - 1] + 1;
    l->thres_h[b] = (b == blocks - 1) ? l->expanded_cols - 1 : l->thres_l[b] + l->cdb - 1;
  }

  /* --- scan_matrix, src/csr_hw.cpp:87-119: padded per-block row lengths (prefix sums) --- */
  uint32_t **brp = (uint32_t **)calloc(blocks, sizeof(uint32_t *));
  for (int b = 0; b < blocks; b++) brp[b] = (uint32_t *)calloc((size_t)rows + 1, 4);
  uint32_t *cur = (uint32_t *)calloc(blocks, 4);
  for (uint32_t r = 0; r < rows; r++) {
    for (uint64_t j = row_ptr[r]; j < row_ptr[r + 1]; j++) {
      uint32_t c = col_ind[j];
      for (int k = 0; k < blocks; k++) /* linear search, src/csr_hw.cpp:97-103 */
        if (c <= l->thres_h[k] && c >= l->thres_l[k]) { cur[k]++; break; }
    }
    for (int k = 0; k < blocks; k++) {
      uint32_t len = cur[k], pad = 0;
      if (len != 0 && len % (uint32_t)vf != 0) pad = (uint32_t)vf - len % (uint32_t)vf; /* :108-112 */
      brp[k][r + 1] = brp[k][r] + len + pad;
      cur[k] = 0;
    }
  }
  free(cur);

  /* --- prepare_balanced_hw_matrix, src/csr_hw.cpp:432-484 (CU=2; same rule for every CU) --- */
  l->bitmap = (uint8_t **)calloc(blocks, sizeof(uint8_t *));
  l->nr_rows = (uint32_t **)calloc(cu, sizeof(uint32_t *));
  l->nr_nzeros = (uint32_t **)calloc(cu, sizeof(uint32_t *));
  l->nr_ci = (uint32_t **)calloc(cu, sizeof(uint32_t *));
  l->nr_val = (uint32_t **)calloc(cu, sizeof(uint32_t *));
  l->words = (uint8_t ***)calloc(cu, sizeof(uint8_t **));
  for (int k = 0; k < cu; k++) {
    l->nr_rows[k] = (uint32_t *)calloc(blocks, 4);
    l->nr_nzeros[k] = (uint32_t *)calloc(blocks, 4);
    l->nr_ci[k] = (uint32_t *)calloc(blocks, 4);
    l->nr_val[k] = (uint32_t *)calloc(blocks, 4);
    l->words[k] = (uint8_t **)calloc(blocks, sizeof(uint8_t *));
  }
  l->nr_cols = (uint32_t *)calloc(blocks, 4);
  for (int b = 0; b < blocks; b++) {
    l->bitmap[b] = (uint8_t *)calloc(rows ? rows : 1, 1);
    uint32_t block_nnz = brp[b][rows];
    uint32_t nz = 0, rc = 0;
    int examined = 1;
    for (uint32_t r = 0; r < rows; r++) {
      uint32_t len = brp[b][r + 1] - brp[b][r];
      if (len == 0) { l->bitmap[b][r] = 1; continue; }
      l->bitmap[b][r] = 0;
      nz += len; rc++;
      if (cu > 1) { /* S1 && S2 && S3, src/csr_hw.cpp:459-468 */
        int S1 = nz > block_nnz / (uint32_t)cu, S2 = nz % (uint32_t)l->ratio_v == 0, S3 = rc % (uint32_t)l->ratio_v == 0;
        if (S1 && S2 && S3) {
          if (examined <= cu - 1) { l->nr_rows[examined - 1][b] = rc; l->nr_nzeros[examined - 1][b] = nz; }
          nz = 0; rc = 0; examined++;
        }
      }
    }
    uint32_t mod = rc % (uint32_t)l->ratio_v; /* last CU: pad rows to RATIO_v, src/csr_hw.cpp:474-482 */
    if (mod != 0) { rc += (uint32_t)l->ratio_v - mod; nz += ((uint32_t)l->ratio_v - mod) * (uint32_t)vf; }
    l->nr_rows[cu - 1][b] = rc;
    l->nr_nzeros[cu - 1][b] = nz;
    l->nr_cols[b] = l->thres_h[b] - l->thres_l[b] + 1; /* hw_matrix_alloc, src/csr_hw.cpp:167 */
  }

  /* --- per block: create_block_matrix (src/csr_hw.cpp:190-265) then the packer --- */
  for (int b = 0; b < blocks; b++) {
    uint32_t tot_nnz = 0, tot_rows = 0;
    for (int k = 0; k < cu; k++) { tot_nnz += l->nr_nzeros[k][b]; tot_rows += l->nr_rows[k][b]; }
    uint32_t *bcol = (uint32_t *)calloc(tot_nnz ? tot_nnz : 1, 4);
    uint8_t *beor = (uint8_t *)calloc(tot_nnz ? tot_nnz : 1, 1);
    uint8_t *bval = (uint8_t *)calloc(tot_nnz ? tot_nnz : 1, vb);
    uint32_t *brow_ptr = (uint32_t *)calloc((size_t)tot_rows + 1, 4);
    uint32_t c_cnt = 0, r_cnt = 0;
    for (uint32_t r = 0; r < rows; r++) {
      if (brp[b][r + 1] - brp[b][r] == 0) continue;
      uint32_t elems = 0;
      for (uint64_t j = row_ptr[r]; j < row_ptr[r + 1]; j++) {
        uint32_t c = col_ind[j];
        if (c <= l->thres_h[b] && c >= l->thres_l[b]) {
          bcol[c_cnt] = c - l->thres_l[b]; /* :220 */
          memcpy(bval + (size_t)c_cnt * vb, (const uint8_t *)values + (size_t)j * vb, vb);
          c_cnt++; elems++;
        }
      }
      while (elems % (uint32_t)vf != 0) { bcol[c_cnt] = 0; c_cnt++; elems++; } /* VF padding, :229-238 (val stays 0) */
      beor[c_cnt - 1] = 1;
      brow_ptr[++r_cnt] = c_cnt;
    }
    while (r_cnt < tot_rows) { /* padding rows, :246-255 */
      c_cnt += (uint32_t)vf;
      if (c_cnt <= tot_nnz) beor[c_cnt - 1] = 1;
      brow_ptr[++r_cnt] = c_cnt;
    }
    /* generate_balanced_hw_matrix: CU pieces in order, src/csr_hw.cpp:486-494; alloc :174-180 */
    uint32_t first_row = 0;
    for (int k = 0; k < cu; k++) {
      uint32_t nnz_k = l->nr_nzeros[k][b];
      l->nr_ci[k][b] = (nnz_k + RATIO_CI - 1) / RATIO_CI;
      l->nr_val[k][b] = nnz_k / (uint32_t)l->ratio_v; /* floors, Q1 */
      uint32_t alloc_words = l->nr_ci[k][b] + (nnz_k + (uint32_t)l->ratio_v - 1) / (uint32_t)l->ratio_v;
      /* a partial last group still owns a full index word placed before its values */
      l->words[k][b] = (uint8_t *)calloc((size_t)(alloc_words ? alloc_words : 1), BUS_BYTES);
      uint32_t first = brow_ptr[first_row];
      pack_piece(l, l->words[k][b], bcol, beor, bval, first, nnz_k);
      first_row += l->nr_rows[k][b];
    }
    free(bcol); free(beor); free(bval); free(brow_ptr);
  }
  for (int b = 0; b < blocks; b++) free(brp[b]);
  free(brp);
  return l;
}

void orc_layout_free(orc_layout *l) {
  if (!l) return;
  for (int k = 0; k < l->cu; k++) {
    for (int b = 0; b < l->blocks; b++) free(l->words[k][b]);
    free(l->words[k]); free(l->nr_rows[k]); free(l->nr_nzeros[k]); free(l->nr_ci[k]); free(l->nr_val[k]);
  }
  for (int b = 0; b < l->blocks; b++) free(l->bitmap[b]);
  free(l->words); free(l->nr_rows); free(l->nr_nzeros); free(l->nr_ci); free(l->nr_val);
  free(l->bitmap); free(l->nr_cols); free(l->thres_l); free(l->thres_h);
  free(l);
}

int orc_blocks(const orc_layout *l) { return l->blocks; }
uint32_t orc_expanded_cols(const orc_layout *l) { return l->expanded_cols; }
void orc_piece_info(const orc_layout *l, int cu, int block, uint32_t *out) {
  out[0] = l->nr_rows[cu][block]; out[1] = l->nr_cols[block]; out[2] = l->nr_nzeros[cu][block];
  out[3] = l->nr_ci[cu][block];   out[4] = l->nr_val[cu][block];
}
const void *orc_piece_words(const orc_layout *l, int cu, int block) { return l->words[cu][block]; }
uint32_t orc_piece_alloc_words(const orc_layout *l, int cu, int block) {
  uint32_t n = l->nr_nzeros[cu][block];
  return l->nr_ci[cu][block] + (n + (uint32_t)l->ratio_v - 1) / (uint32_t)l->ratio_v;
}
const uint8_t *orc_bitmap_row(const orc_layout *l, int block) { return l->bitmap[block]; }

/* write_csr_hw_vector, src/csr_hw.cpp:1470-1488: consecutive x values, zero beyond n */
void orc_hw_x(const orc_layout *l, const void *x, uint32_t n, void *out) {
  const int vb = vbytes(l);
  memset(out, 0, (size_t)l->expanded_cols * vb);
  uint32_t m = n < l->expanded_cols ? n : l->expanded_cols;
  memcpy(out, x, (size_t)m * vb);
}

/* compute_results (src/spmv.cpp:66-104) + write_back + accum_results (src/csr_hw.cpp:1531-1565),
 * instantiated for double and float so the arithmetic happens in ValueType like the reference. */
#define DEFINE_EMU(NAME, VT)                                                                                   \
  static int NAME(const orc_layout *l, const VT *hwx, VT *y) {                                                 \
    for (int b = 0; b < l->blocks; b++) {                                                                      \
      const VT *x_local = hwx + l->thres_l[b]; /* L0 copy, src/spmv.cpp:182-192 */                             \
      uint32_t offset = 0;                                                                                     \
      for (int k = 0; k < l->cu; k++) {                                                                        \
        uint32_t nnz = l->nr_nzeros[k][b], nrows = l->nr_rows[k][b];                                           \
        const uint8_t *w = l->words[k][b];                                                                     \
        VT *hw_y = (VT *)calloc(nrows ? nrows : 1, sizeof(VT));                                                \
        uint32_t emitted = 0;                                                                                  \
        VT sum = 0;                                                                                            \
        for (uint32_t i = 0; i < nnz; i += (uint32_t)l->vf) {                                                  \
          VT sum_tmp = 0;                                                                                      \
          uint16_t last_ci = 0;                                                                                \
          for (uint32_t j = 0; j < (uint32_t)l->vf; j++) {                                                     \
            uint32_t e = i + j, g = e / RATIO_CI, s = e % RATIO_CI;                                            \
            const uint8_t *grp = w + (size_t)g * l->ratio_col_val * BUS_BYTES;                                 \
            uint16_t ci; VT v;                                                                                 \
            memcpy(&ci, grp + 2 * s, 2);                                                                       \
            memcpy(&v, grp + BUS_BYTES + (size_t)s * sizeof(VT), sizeof(VT));                                  \
            VT term = v * x_local[ci & 0x7FFF]; /* src/spmv.cpp:84-88 */                                       \
            sum_tmp += term;                    /* :91-96 */                                                   \
            last_ci = ci;                                                                                      \
          }                                                                                                    \
          sum += sum_tmp; /* :97 */                                                                            \
          if (last_ci & 0x8000) { /* :99-102 */                                                                \
            if (emitted < nrows) hw_y[emitted] = sum;                                                          \
            emitted++;                                                                                         \
            sum = 0;                                                                                           \
          }                                                                                                    \
        }                                                                                                      \
        if (emitted != nrows) { free(hw_y); return 1; }                                                        \
        /* accum_results */                                                                                    \
        uint32_t i = 0;                                                                                        \
        for (uint32_t j = 0; j < nrows; j++) {                                                                 \
          if (i + offset < l->rows) {                                                                          \
            while (i + offset < l->rows && l->bitmap[b][i + offset] != 0) i++; /* bounded: Q4 */               \
            if (i + offset >= l->rows) break;                                                                  \
            y[i + offset] += hw_y[j];                                                                          \
            i++;                                                                                               \
          } else break;                                                                                        \
        }                                                                                                      \
        offset += i;                                                                                           \
        free(hw_y);                                                                                            \
      }                                                                                                        \
    }                                                                                                          \
    return 0;                                                                                                  \
  }
DEFINE_EMU(emu_f64, double)
DEFINE_EMU(emu_f32, float)

int orc_spmv_emu(const orc_layout *l, const void *x, uint32_t n, void *y) {
  void *hwx = malloc((size_t)l->expanded_cols * vbytes(l) + 16);
  orc_hw_x(l, x, n, hwx);
  int rc = l->is_double ? emu_f64(l, (const double *)hwx, (double *)y) : emu_f32(l, (const float *)hwx, (float *)y);
  free(hwx);
  return rc;
}

/* spmv_gold, src/csr.cpp:184-194, with 64-bit row offsets */
void orc_spmv_gold(uint32_t rows, const uint64_t *row_ptr, const uint32_t *col_ind, const void *values,
                   const void *x, void *y, int is_double) {
  if (is_double) {
    const double *v = (const double *)values, *xx = (const double *)x;
    double *yy = (double *)y;
    for (uint32_t i = 0; i < rows; i++) {
      double acc = 0.0;
      for (uint64_t j = row_ptr[i]; j < row_ptr[i + 1]; j++) acc += v[j] * xx[col_ind[j]];
      yy[i] = acc;
    }
  } else {
    const float *v = (const float *)values, *xx = (const float *)x;
    float *yy = (float *)y;
    for (uint32_t i = 0; i < rows; i++) {
      float acc = 0.0f;
      for (uint64_t j = row_ptr[i]; j < row_ptr[i + 1]; j++) acc += v[j] * xx[col_ind[j]];
      yy[i] = acc;
    }
  }
}

/* The same loop with the rows split over all host cores (baseline ii): every row is still
 * summed left to right by one thread, so the result is bit-identical to orc_spmv_gold.  Returns the thread count. */
int orc_spmv_gold_omp(uint32_t rows, const uint64_t *row_ptr, const uint32_t *col_ind, const void *values,
                      const void *x, void *y, int is_double) {
  int threads = 1;
#ifdef _OPENMP
  threads = omp_get_max_threads();
#endif
  if (is_double) {
    const double *v = (const double *)values, *xx = (const double *)x;
    double *yy = (double *)y;
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < (int64_t)rows; i++) {
      double acc = 0.0;
      for (uint64_t j = row_ptr[i]; j < row_ptr[i + 1]; j++) acc += v[j] * xx[col_ind[j]];
      yy[i] = acc;
    }
  } else {
    const float *v = (const float *)values, *xx = (const float *)x;
    float *yy = (float *)y;
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < (int64_t)rows; i++) {
      float acc = 0.0f;
      for (uint64_t j = row_ptr[i]; j < row_ptr[i + 1]; j++) acc += v[j] * xx[col_ind[j]];
      yy[i] = acc;
    }
  }
  return threads;
}

/* per-row sum |a||x| in double: the normaliser of the north-star tolerance */
void orc_abs_ax(uint32_t rows, const uint64_t *row_ptr, const uint32_t *col_ind, const void *values,
                const void *x, double *out, int is_double) {
#pragma omp parallel for schedule(static, 4096)
  for (int64_t i = 0; i < (int64_t)rows; i++) {
    double acc = 0.0;
    for (uint64_t j = row_ptr[i]; j < row_ptr[i + 1]; j++) {
      double a = is_double ? ((const double *)values)[j] : (double)((const float *)values)[j];
      double b = is_double ? ((const double *)x)[col_ind[j]] : (double)((const float *)x)[col_ind[j]];
      acc += fabs(a) * fabs(b);
    }
    out[i] = acc;
  }
}

/* verification, src/csr_hw.cpp:1571-1590: |a-b| >= 1e-5 or NaN => fail */
int orc_verification(uint32_t n, const void *sw, const void *hw, int is_double) {
  int status = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (is_double) {
      double d = fabs(((const double *)sw)[i] - ((const double *)hw)[i]);
      if (d >= 1e-5 || d != d) status = 1;
    } else {
      float d = fabsf(((const float *)sw)[i] - ((const float *)hw)[i]);
      if (d >= (float)1e-5 || d != d) status = 1;
    }
  }
  return status;
}
