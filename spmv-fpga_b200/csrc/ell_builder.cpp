// Host builder of the sliced-ELLPACK device image (ell.h) + its inspection API for the tests.
#include <omp.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "../../include/spmvb.h"
#include "ell.h"
#include "layout.h"

namespace spmvb {

EllImage::~EllImage() { free(image); }

template <typename RP>
EllImage *build_ell(uint32_t rows, uint32_t cols, const RP *row_ptr, const uint32_t *col_ind, const void *values, int is_double) {
  if (options().ell == 0 || rows == 0) return nullptr;
  const uint64_t nnz = (uint64_t)row_ptr[rows];
  if (nnz == 0) return nullptr;
  const uint32_t n_slices = (rows + kEllSliceRows - 1) / kEllSliceRows;
  // pass 1: longest row, column range of every slice
  std::vector<uint32_t> lo(n_slices, 0xFFFFFFFFu), hi(n_slices, 0);
  uint32_t width = 0;
  int wide_slice = 0;
#pragma omp parallel for schedule(static) reduction(max : width) reduction(| : wide_slice)
  for (int64_t s = 0; s < (int64_t)n_slices; s++) {
    const uint32_t r0 = (uint32_t)s * kEllSliceRows, r1 = std::min<uint32_t>(rows, r0 + kEllSliceRows);
    uint32_t l = 0xFFFFFFFFu, h = 0;
    for (uint64_t j = row_ptr[r0]; j < (uint64_t)row_ptr[r1]; j++) { l = std::min(l, col_ind[j]); h = std::max(h, col_ind[j]); }
    for (uint32_t r = r0; r < r1; r++) width = std::max<uint32_t>(width, (uint32_t)std::min<uint64_t>(row_ptr[r + 1] - row_ptr[r], 0xFFFFFFFFull));
    lo[s] = l; hi[s] = h;
    if (l <= h && h - l > 0xFFFFu) wide_slice = 1;
  }
  if (width == 0 || width > (uint32_t)kEllMaxWidth || wide_slice) return nullptr;
  const uint64_t slots = (uint64_t)n_slices * kEllSliceRows * width;
  if (options().ell < 1 && (double)slots > 1.04 * (double)nnz) return nullptr;  // padding would cost more than the gathers save
  EllImage *E = new EllImage();
  E->is_double = is_double ? 1 : 0; E->vb = is_double ? 8 : 4;
  E->rows = rows; E->cols = cols; E->n_slices = n_slices; E->width = width;
  E->slice_bytes = ell_slice_bytes(width, E->vb);
  E->real_nnz = nnz; E->slots = slots; E->bytes = (uint64_t)n_slices * E->slice_bytes;
  E->image = (uint8_t *)calloc((size_t)std::max<uint64_t>(E->bytes, 16), 1);
  if (!E->image) { delete E; return nullptr; }
  E->col_lo.swap(lo); E->col_hi.swap(hi);
  const uint8_t *vals = (const uint8_t *)values;
  const int vb = E->vb;
#pragma omp parallel for schedule(static)
  for (int64_t s = 0; s < (int64_t)n_slices; s++) {
    uint8_t *rec = E->image + (uint64_t)s * E->slice_bytes;
    const uint32_t r0 = (uint32_t)s * kEllSliceRows;
    const uint32_t base = E->col_lo[s] <= E->col_hi[s] ? E->col_lo[s] : 0u;
    const uint32_t head[4] = {base, width, r0, 0u};
    memcpy(rec, head, 16);
    uint8_t *idx = rec + 16, *val = rec + 16 + (size_t)width * 64;
    for (uint32_t l = 0; l < (uint32_t)kEllSliceRows; l++) {
      const uint32_t r = r0 + l;
      const uint64_t j0 = r < rows ? (uint64_t)row_ptr[r] : 0, j1 = r < rows ? (uint64_t)row_ptr[r + 1] : 0;
      const uint16_t pad = j1 > j0 ? (uint16_t)(col_ind[j0] - base) : (uint16_t)0;  // a column the row (or slice) reads anyway
      for (uint32_t k = 0; k < width; k++) {
        const bool real = j0 + k < j1;
        const uint16_t off = real ? (uint16_t)(col_ind[j0 + k] - base) : pad;
        memcpy(idx + (size_t)k * 64 + 2 * l, &off, 2);
        if (real) memcpy(val + ((size_t)k * kEllSliceRows + l) * vb, vals + (j0 + k) * vb, vb);
      }
    }
  }
  return E;
}

template EllImage *build_ell<uint64_t>(uint32_t, uint32_t, const uint64_t *, const uint32_t *, const void *, int);
template EllImage *build_ell<uint32_t>(uint32_t, uint32_t, const uint32_t *, const uint32_t *, const void *, int);

}  // namespace spmvb

using namespace spmvb;

extern "C" {

int spmvb_layout_ell_params(const spmvb_layout *l, uint64_t *out) {
  const Layout *L = (const Layout *)l;
  if (!L || !out) return fail(SPMVB_E_ARG, "ell_params");
  for (int i = 0; i < 8; i++) out[i] = 0;
  const EllImage *E = L->ell;
  if (!E) return SPMVB_OK;
  out[0] = 1; out[1] = E->width; out[2] = E->n_slices; out[3] = E->slice_bytes; out[4] = E->slots; out[5] = E->bytes;
  out[6] = E->real_nnz;
  return SPMVB_OK;
}

int64_t spmvb_layout_ell_image(const spmvb_layout *l, void *out, uint64_t max_bytes) {
  const Layout *L = (const Layout *)l;
  if (!L) return fail(SPMVB_E_ARG, "ell_image");
  if (!L->ell) return 0;
  if (out && max_bytes >= L->ell->bytes) memcpy(out, L->ell->image, (size_t)L->ell->bytes);
  return (int64_t)L->ell->bytes;
}

int64_t spmvb_layout_ell_decode(const spmvb_layout *l, uint32_t *cols_out, void *vals_out, uint64_t max_slots) {
  const Layout *L = (const Layout *)l;
  if (!L || !L->ell) return fail(SPMVB_E_ARG, "ell_decode: no ELL image");
  const EllImage *E = L->ell;
  uint64_t n = 0;
  for (uint32_t s = 0; s < E->n_slices; s++) {
    const uint8_t *rec = E->image + (uint64_t)s * E->slice_bytes;
    uint32_t head[4];
    memcpy(head, rec, 16);
    if (head[1] != E->width || head[2] != s * (uint32_t)kEllSliceRows || head[3] != 0) return fail(SPMVB_E_ARG, "ell_decode: slice header");
    const uint8_t *idx = rec + 16, *val = rec + 16 + (size_t)E->width * 64;
    for (uint32_t lane = 0; lane < (uint32_t)kEllSliceRows; lane++)  // row-major output: slot k of row s * 32 + lane
      for (uint32_t k = 0; k < E->width; k++, n++) {
        if (n >= max_slots) continue;
        uint16_t off;
        memcpy(&off, idx + (size_t)k * 64 + 2 * lane, 2);
        if (cols_out) cols_out[n] = head[0] + off;
        if (vals_out) memcpy((uint8_t *)vals_out + n * E->vb, val + ((size_t)k * kEllSliceRows + lane) * E->vb, E->vb);
      }
  }
  return (int64_t)n;
}

}  // extern "C"
