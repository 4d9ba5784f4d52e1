// Engine-private device image for REGULAR matrices (rows of almost equal length whose columns stay close together: a
// stencil, a band): sliced ELLPACK with 16-bit slice-relative columns.  The reference format (8-entry groups of one
// row after the other, end-of-row bits) makes a warp's gather request touch as many lines of x as its 32 lanes are
// rows apart times the bands of the matrix - 12 distinct 128-byte lines per request on the 5-point Laplacian, which is
// what bounds the global-gather kernel there (the LSU takes a cycle per distinct line).  Here a slice is
// 32 consecutive rows, lane l owns row l, and slot s of all 32 rows is one request: 32 neighbouring columns, 2-3 lines.
// No end-of-row bits, no row map, no atomics, no rows to clear: every row is written exactly once.
//
// Image: n_slices records of slice_bytes = 16 + w * 64 + w * 32 * vb each (w = the longest row of the matrix):
//   [0, 16)                    base column of the slice (uint32), w, first row, 0
//   [16, 16 + w * 64)          column offsets: slot s, lane l -> uint16 (column - base) at 16 + s * 64 + 2 l
//   [16 + w * 64, ...)         values: slot s, lane l at s * 32 * vb + l * vb
// Rows shorter than w are padded with (the row's first column, 0).  Same 10 (fp64) / 6 (fp32) bytes per entry as the
// hw_matrix stream.  Built only when it costs next to nothing: w <= kEllMaxWidth, every slice's columns within 65 536
// of each other, and rows * w <= 1.04 x nnz.
#pragma once
#include <cstdint>
#include <vector>

namespace spmvb {

constexpr int kEllMaxWidth = 16;
constexpr int kEllSliceRows = 32;

struct EllImage {
  int is_double = 1, vb = 8;
  uint32_t rows = 0, cols = 0, n_slices = 0, width = 0, slice_bytes = 0;
  uint64_t real_nnz = 0, slots = 0, bytes = 0;
  uint8_t *image = nullptr;
  std::vector<uint32_t> col_lo, col_hi;  // per slice: first / last column it reads (col_lo > col_hi: none)
  ~EllImage();
};

inline uint32_t ell_slice_bytes(uint32_t w, int vb) { return 16u + w * 64u + w * 32u * (uint32_t)vb; }

// nullptr when the matrix does not qualify (or option ell = 0)
template <typename RP>
EllImage *build_ell(uint32_t rows, uint32_t cols, const RP *row_ptr, const uint32_t *col_ind, const void *values, int is_double);

}  // namespace spmvb
