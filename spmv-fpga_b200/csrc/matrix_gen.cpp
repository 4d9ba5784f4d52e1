// Synthetic inputs of BASELINE.json's configs as library-owned CSR matrices (band, 2-D 5-point Laplacian, uniform
// random, R-MAT), seeded and generated in parallel; [row_begin, row_end) selects a row slice of the same global matrix.
// This file and errors.cpp depend on nothing else: they are also built alone into oracle/_ref/libmatgen.so, which is how
// the reference arm of bench.py gets the very same matrices without loading the engine.
#include <omp.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/spmvb.h"
#include "csr.h"
#include "errors.h"

using namespace spmvb;

extern "C" {

void spmvb_csr_free(spmvb_csr *m) { delete (Csr *)m; }
uint32_t spmvb_csr_rows(const spmvb_csr *m) { return ((const Csr *)m)->rows; }
uint32_t spmvb_csr_cols(const spmvb_csr *m) { return ((const Csr *)m)->cols; }
uint64_t spmvb_csr_nnz(const spmvb_csr *m) { return ((const Csr *)m)->nnz(); }
int spmvb_csr_is_double(const spmvb_csr *m) { return ((const Csr *)m)->is_double; }
const uint64_t *spmvb_csr_row_ptr(const spmvb_csr *m) { return ((const Csr *)m)->row_ptr.data(); }
const uint32_t *spmvb_csr_col_ind(const spmvb_csr *m) { return ((const Csr *)m)->col_ind.data(); }
const void *spmvb_csr_values(const spmvb_csr *m) { return ((const Csr *)m)->values.data(); }

int spmvb_csr_gen_band(uint32_t n, int hb, uint64_t seed, int is_double, spmvb_csr **out) {
  if (!out || n == 0 || hb < 0) return fail(SPMVB_E_ARG, "gen_band");
  Csr *A = new Csr();
  A->rows = A->cols = n; A->is_double = is_double ? 1 : 0;
  A->row_ptr.assign((size_t)n + 1, 0);
  for (uint32_t r = 0; r < n; r++) {
    uint32_t lo = r >= (uint32_t)hb ? r - hb : 0, hi = std::min<uint64_t>((uint64_t)r + hb, n - 1);
    A->row_ptr[r + 1] = A->row_ptr[r] + (hi - lo + 1);
  }
  const uint64_t nnz = A->nnz();
  A->col_ind.resize(nnz);
  A->values.resize(nnz * (is_double ? 8 : 4));
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < (int64_t)n; r++) {
    uint32_t lo = r >= hb ? (uint32_t)r - hb : 0;
    uint64_t j = A->row_ptr[r];
    for (uint64_t k = 0; k < A->row_ptr[r + 1] - A->row_ptr[r]; k++, j++) {
      A->col_ind[j] = lo + (uint32_t)k;
      A->set_value(j, value_at(seed, (uint32_t)r, lo + (uint32_t)k));
    }
  }
  *out = (spmvb_csr *)A;
  return SPMVB_OK;
}

int spmvb_csr_gen_laplacian2d(uint32_t nx, uint32_t ny, uint32_t row_begin, uint32_t row_end, int is_double,
                              spmvb_csr **out) {
  if (!out || nx == 0 || ny == 0 || (uint64_t)nx * ny > 0xFFFFFFFFull) return fail(SPMVB_E_ARG, "gen_laplacian2d");
  const uint32_t n = nx * ny;
  if (row_end == 0) row_end = n;
  if (row_begin >= row_end || row_end > n) return fail(SPMVB_E_ARG, "gen_laplacian2d: row range");
  Csr *A = new Csr();
  A->rows = row_end - row_begin; A->cols = n; A->is_double = is_double ? 1 : 0;
  A->row_ptr.assign((size_t)A->rows + 1, 0);
  for (uint32_t i = 0; i < A->rows; i++) {
    uint32_t r = row_begin + i, ix = r % nx, iy = r / nx;
    uint32_t d = 1 + (iy > 0) + (ix > 0) + (ix + 1 < nx) + (iy + 1 < ny);
    A->row_ptr[i + 1] = A->row_ptr[i] + d;
  }
  const uint64_t nnz = A->nnz();
  A->col_ind.resize(nnz);
  A->values.resize(nnz * (is_double ? 8 : 4));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)A->rows; i++) {
    uint32_t r = row_begin + (uint32_t)i, ix = r % nx, iy = r / nx;
    uint64_t j = A->row_ptr[i];
    if (iy > 0) { A->col_ind[j] = r - nx; A->set_value(j++, -1.0); }
    if (ix > 0) { A->col_ind[j] = r - 1; A->set_value(j++, -1.0); }
    A->col_ind[j] = r; A->set_value(j++, 4.0);
    if (ix + 1 < nx) { A->col_ind[j] = r + 1; A->set_value(j++, -1.0); }
    if (iy + 1 < ny) { A->col_ind[j] = r + nx; A->set_value(j++, -1.0); }
  }
  *out = (spmvb_csr *)A;
  return SPMVB_OK;
}

int spmvb_csr_gen_uniform(uint32_t rows, uint32_t cols, int k, uint64_t seed, uint32_t row_begin, uint32_t row_end,
                          int is_double, spmvb_csr **out) {
  if (!out || rows == 0 || cols == 0 || k < 1 || (uint32_t)k > cols || k > 1024) return fail(SPMVB_E_ARG, "gen_uniform");
  if (row_end == 0) row_end = rows;
  if (row_begin >= row_end || row_end > rows) return fail(SPMVB_E_ARG, "gen_uniform: row range");
  Csr *A = new Csr();
  A->rows = row_end - row_begin; A->cols = cols; A->is_double = is_double ? 1 : 0;
  A->row_ptr.resize((size_t)A->rows + 1);
  for (uint64_t i = 0; i <= A->rows; i++) A->row_ptr[i] = i * (uint64_t)k;
  const uint64_t nnz = A->nnz();
  A->col_ind.resize(nnz);
  A->values.resize(nnz * (is_double ? 8 : 4));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)A->rows; i++) {
    const uint32_t r = row_begin + (uint32_t)i;
    uint32_t c[1024];
    uint64_t ctr = 0;
    const uint64_t key = mix64(seed * 0x100000001B3ull + r);
    int have = 0;
    while (have < k) {  // draw, sort, drop duplicates, top up
      while (have < k) c[have++] = (uint32_t)(((mix64(key + ctr++) >> 32) * (uint64_t)cols) >> 32);
      std::sort(c, c + k);
      have = (int)(std::unique(c, c + k) - c);
    }
    uint64_t j = (uint64_t)i * k;
    for (int q = 0; q < k; q++, j++) {
      A->col_ind[j] = c[q];
      A->set_value(j, value_at(seed, r, c[q]));
    }
  }
  *out = (spmvb_csr *)A;
  return SPMVB_OK;
}

int spmvb_csr_gen_rmat(int scale, int ef, double a, double b, double c, uint64_t seed, uint32_t row_begin,
                       uint32_t row_end, int is_double, spmvb_csr **out) {
  if (!out || scale < 1 || scale > 31 || ef < 1 || a <= 0 || b < 0 || c < 0 || a + b + c >= 1.0)
    return fail(SPMVB_E_ARG, "gen_rmat");
  const uint32_t n = 1u << scale;
  if (row_end == 0) row_end = n;
  if (row_begin >= row_end || row_end > n) return fail(SPMVB_E_ARG, "gen_rmat: row range");
  const uint64_t edges = (uint64_t)ef << scale;
  // thresholds on a 32-bit uniform
  const uint64_t ta = (uint64_t)(a * 4294967296.0), tab = (uint64_t)((a + b) * 4294967296.0),
                 tabc = (uint64_t)((a + b + c) * 4294967296.0);
  // bucket by the top bits of the row so that buckets can be sorted independently
  const int bucket_bits = std::min(scale, 12);
  const uint32_t nb = 1u << bucket_bits;
  const int shift = scale - bucket_bits;
  const int T = std::max(1, omp_get_max_threads());
  std::vector<uint64_t> cnt((size_t)T * nb, 0);
  auto edge = [&](uint64_t e, uint32_t &r, uint32_t &cc) {
    uint64_t key = mix64(seed ^ (e * 0xD1342543DE82EF95ull));
    r = 0; cc = 0;
    for (int lvl = 0; lvl < scale; lvl += 2) {  // one 64-bit hash feeds two levels
      uint64_t h = mix64(key + (uint64_t)lvl);
      for (int half = 0; half < 2 && lvl + half < scale; half++) {
        uint64_t u = (half ? (h >> 32) : (h & 0xFFFFFFFFull));
        uint32_t rb = u >= tab, cb = (u >= ta && u < tab) || u >= tabc;
        r = (r << 1) | rb; cc = (cc << 1) | cb;
      }
    }
  };
  std::vector<uint64_t> tstart(T + 1);
  for (int t = 0; t <= T; t++) tstart[t] = edges / T * t;
  tstart[T] = edges;
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    uint64_t *ct = &cnt[(size_t)t * nb];
    for (uint64_t e = tstart[t]; e < tstart[t + 1]; e++) {
      uint32_t r, cc;
      edge(e, r, cc);
      if (r >= row_begin && r < row_end) ct[r >> shift]++;
    }
  }
  std::vector<uint64_t> bstart(nb + 1, 0);
  for (uint32_t q = 0; q < nb; q++) {
    uint64_t s = 0;
    for (int t = 0; t < T; t++) { uint64_t v = cnt[(size_t)t * nb + q]; cnt[(size_t)t * nb + q] = bstart[q] + s; s += v; }
    bstart[q + 1] = bstart[q] + s;
  }
  std::vector<uint64_t> keys(bstart[nb] + 1);
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    uint64_t *ct = &cnt[(size_t)t * nb];
    for (uint64_t e = tstart[t]; e < tstart[t + 1]; e++) {
      uint32_t r, cc;
      edge(e, r, cc);
      if (r >= row_begin && r < row_end) keys[ct[r >> shift]++] = ((uint64_t)r << 32) | cc;
    }
  }
  std::vector<uint64_t> uniq(nb + 1, 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t q = 0; q < (int64_t)nb; q++) {
    uint64_t *lo = keys.data() + bstart[q], *hi = keys.data() + bstart[q + 1];
    std::sort(lo, hi);
    uniq[q + 1] = (uint64_t)(std::unique(lo, hi) - lo);
  }
  for (uint32_t q = 0; q < nb; q++) uniq[q + 1] += uniq[q];
  Csr *A = new Csr();
  A->rows = row_end - row_begin; A->cols = n; A->is_double = is_double ? 1 : 0;
  // make the globally last row non-empty (reference reader defect Q3)
  const bool owns_last = row_end == n;
  bool need_last = false;
  if (owns_last) {
    const uint32_t q = (n - 1) >> shift;
    const uint64_t cntq = uniq[q + 1] - uniq[q];
    need_last = cntq == 0 || (keys[bstart[q] + cntq - 1] >> 32) != n - 1;
  }
  const uint64_t nnz = uniq[nb] + (need_last ? 1 : 0);
  A->row_ptr.assign((size_t)A->rows + 1, 0);
  A->col_ind.resize(nnz);
  A->values.resize(nnz * (is_double ? 8 : 4));
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t q = 0; q < (int64_t)nb; q++) {
    const uint64_t *src = keys.data() + bstart[q];
    uint64_t dst = uniq[q];
    for (uint64_t i = 0; i < uniq[q + 1] - uniq[q]; i++, dst++) {
      const uint32_t r = (uint32_t)(src[i] >> 32), cc = (uint32_t)src[i];
      A->col_ind[dst] = cc;
      A->set_value(dst, value_at(seed, r, cc));
      A->row_ptr[(size_t)(r - row_begin) + 1]++;  // rows of a bucket belong to this thread only
    }
  }
  if (need_last) {
    A->col_ind[nnz - 1] = n - 1;
    A->set_value(nnz - 1, value_at(seed, n - 1, n - 1));
    A->row_ptr[A->rows]++;
  }
  for (uint32_t i = 0; i < A->rows; i++) A->row_ptr[i + 1] += A->row_ptr[i];
  *out = (spmvb_csr *)A;
  return SPMVB_OK;
}

}  // extern "C"
