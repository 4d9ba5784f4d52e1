// The sliced-ELLPACK image (ell.h) built on the GPU from a device-resident CSR: the same bytes as build_ell() on the
// host (tests compare them).  Two kernels, one warp per slice of 32 rows, lane l = row l:
//   ell_scan_kernel  longest row, first / last column of every slice, "does not qualify" flags
//   ell_fill_kernel  header, 16-bit slice-relative columns slot-major, values slot-major, padding = (first column, 0)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ell.h"

namespace spmvb {

struct EllScan { uint32_t width, bad; };

__global__ void ell_scan_kernel(uint32_t rows, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col_ind,
                                uint32_t n_slices, uint32_t *__restrict__ lo, uint32_t *__restrict__ hi, EllScan *out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t slice = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slice >= n_slices) return;
  const uint32_t r = slice * kEllSliceRows + lane;
  uint64_t j0 = 0, j1 = 0;
  if (r < rows) { j0 = row_ptr[r]; j1 = row_ptr[r + 1]; }
  const uint64_t len64 = j1 - j0;
  uint32_t len = len64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)len64;
  uint32_t l = 0xFFFFFFFFu, h = 0;
  if (len <= (uint32_t)kEllMaxWidth)  // a longer row disqualifies the matrix anyway: do not walk it
    for (uint64_t j = j0; j < j1; j++) { const uint32_t c = col_ind[j]; l = min(l, c); h = max(h, c); }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    len = max(len, __shfl_xor_sync(0xFFFFFFFFu, len, d));
    l = min(l, __shfl_xor_sync(0xFFFFFFFFu, l, d));
    h = max(h, __shfl_xor_sync(0xFFFFFFFFu, h, d));
  }
  if (lane == 0) {
    lo[slice] = l; hi[slice] = h;
    atomicMax(&out->width, len);
    if (l <= h && h - l > 0xFFFFu) atomicOr(&out->bad, 1u);
  }
}

template <typename VT>
__global__ void ell_fill_kernel(uint32_t rows, const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col_ind,
                                const VT *__restrict__ values, uint32_t n_slices, uint32_t width, uint32_t slice_bytes,
                                const uint32_t *__restrict__ lo, const uint32_t *__restrict__ hi, uint8_t *__restrict__ image) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t slice = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slice >= n_slices) return;
  uint8_t *rec = image + (uint64_t)slice * slice_bytes;
  const uint32_t base = lo[slice] <= hi[slice] ? lo[slice] : 0u;
  if (lane == 0) *reinterpret_cast<uint4 *>(rec) = make_uint4(base, width, slice * kEllSliceRows, 0u);
  const uint32_t r = slice * kEllSliceRows + lane;
  uint64_t j0 = 0, j1 = 0;
  if (r < rows) { j0 = row_ptr[r]; j1 = row_ptr[r + 1]; }
  const uint16_t pad = j1 > j0 ? (uint16_t)(col_ind[j0] - base) : (uint16_t)0;
  uint16_t *idx = reinterpret_cast<uint16_t *>(rec + 16);
  VT *val = reinterpret_cast<VT *>(rec + 16 + (size_t)width * 64);
  for (uint32_t k = 0; k < width; k++) {
    const bool real = j0 + k < j1;
    idx[k * 32 + lane] = real ? (uint16_t)(col_ind[j0 + k] - base) : pad;
    val[k * 32 + lane] = real ? values[j0 + k] : VT(0);
  }
}

}  // namespace spmvb
