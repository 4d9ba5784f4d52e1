// Host-side work plan of the XS kernel (x window in shared memory).  No CUDA here: testable on the CPU.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/spmvb.h"
#include "layout.h"

namespace spmvb {

// Work items of the XS kernel.  Block-major layouts: every CTA (one per SM) owns ONE contiguous range of chunks of
// (almost) equal length - static and balanced, no quantisation loss - and the range is cut into items at piece
// boundaries and wherever the x window the chunks touch would outgrow the shared-memory budget.  CU-major layouts
// (row tiles): see below.  Cut points are piece-relative multiples of
// U = run length x warps per CTA, so that every warp of the CTA gets the same number of whole runs per item.
void build_xs_items(const Layout *L, int n_cta, uint32_t run_log2, std::vector<XsItem> &items,
                    std::vector<uint32_t> &cta_first, XsTilePlan *tiles) {
  const XsConfig cfg = xs_config(L->is_double, L->xs_cfg);
  const uint64_t U = ((uint64_t)1 << run_log2) * (uint64_t)cfg.warps;
  const uint32_t align = 16u / (uint32_t)L->vb;  // window start in elements: 16-byte aligned for the bulk copy
  // candidate cut points in global chunk indices: block starts and block-relative multiples of U
  std::vector<uint64_t> cuts;
  // pieces in device order; an item never spans two pieces (a piece belongs to one column block)
  std::vector<uint64_t> seg_start, seg_end;
  std::vector<uint32_t> seg_block;
  for (uint32_t bk : L->dev_order) {
    if (L->piece_chunk1[bk] == L->piece_chunk0[bk]) continue;
    seg_start.push_back(L->piece_chunk0[bk]); seg_end.push_back(L->piece_chunk1[bk]); seg_block.push_back(bk / (uint32_t)L->cu);
    for (uint64_t c = L->piece_chunk0[bk]; c < L->piece_chunk1[bk]; c += U) cuts.push_back(c);
  }
  cuts.push_back(L->n_chunks);
  cta_first.assign(n_cta + 1, 0);
  // items of the cut range [ci, ce): each extends unit by unit inside its piece while the window fits (and, with a
  // cap, while it holds fewer than `cap` chunks)
  auto emit = [&](size_t ci, size_t ce, uint64_t cap) {
    while (ci < ce) {
      const uint64_t c0 = cuts[ci];
      const size_t sg = (size_t)(std::upper_bound(seg_start.begin(), seg_start.end(), c0) - seg_start.begin()) - 1;
      const int b = (int)seg_block[sg];
      const uint64_t b_end = seg_end[sg];
      uint32_t lo = 0xFFFF, hi = 0;
      size_t e = ci;
      bool fits = true;
      while (e < ce && cuts[e] < b_end && (cap == 0 || cuts[e] - c0 < cap)) {  // extend by one unit while the window fits
        uint32_t nlo = lo, nhi = hi;
        for (uint64_t q = cuts[e]; q < cuts[e + 1]; q++)
          if (L->chunk_col_lo[q] <= L->chunk_col_hi[q]) {
            nlo = std::min<uint32_t>(nlo, L->chunk_col_lo[q]);
            nhi = std::max<uint32_t>(nhi, L->chunk_col_hi[q]);
          }
        const uint32_t wlo = nlo == 0xFFFF ? 0 : nlo / align * align;
        const uint64_t bytes = nlo == 0xFFFF ? 16 : ((uint64_t)(nhi - wlo + 1) * L->vb + 15) / 16 * 16;
        if (bytes > cfg.cap) {
          if (e == ci) { fits = false; e++; }  // even one unit does not fit: gather from global memory
          break;
        }
        lo = nlo; hi = nhi; e++;
      }
      XsItem it{};
      it.chunk_begin = (uint32_t)c0; it.chunk_count = (uint32_t)(cuts[e] - c0); it.block = (uint32_t)b;
      if (fits && lo != 0xFFFF) {
        const uint32_t wlo = lo / align * align;
        it.col_base = wlo;
        it.x_off = (uint32_t)((uint64_t)b * L->cdb + wlo);
        it.x_bytes = (uint32_t)(((uint64_t)(hi - wlo + 1) * L->vb + 15) / 16 * 16);
      } else if (fits) {
        it.x_bytes = 16; it.x_off = (uint32_t)((uint64_t)b * L->cdb);  // only padding in there: any window will do
      }
      items.push_back(it);
      ci = e;
    }
  };
  if (!L->cu_major) {
    size_t ci = 0;  // index into cuts of the current position
    for (int j = 0; j < n_cta; j++) {
      cta_first[j] = (uint32_t)items.size();
      const uint64_t want_end = L->n_chunks * (uint64_t)(j + 1) / (uint64_t)n_cta;
      size_t ce = ci;  // first cut >= want_end (the last CTA takes everything)
      while (ce + 1 < cuts.size() && (cuts[ce] < want_end || j == n_cta - 1)) ce++;
      if (j == n_cta - 1) ce = cuts.size() - 1;
      emit(ci, ce, 0);
      ci = ce;
    }
    cta_first[n_cta] = (uint32_t)items.size();
  } else {
    // Tall matrices (pieces in CU-major order so that the y range in flight stays L2-resident).  A piece here is one
    // column block of one row tile: short (the tile's share of the block) but with a full-width x window, so a window
    // must serve as many chunks as it can - an item is a WHOLE piece (up to ~256 chunks) - and the items are dealt
    // round-robin, so that all CTAs work on neighbouring column blocks of the same row tile at any time.
    const uint64_t cap = std::max<uint64_t>(U, 256 / U * U);
    emit(0, cuts.size() - 1, cap);
    if (tiles) {
      // The same items dealt tile by tile (one launch per row tile: spmv_host sends a tile's rows of y to the host
      // while the next tile is computed).  In CU-major order tile k's pieces are contiguous: chunks
      // [piece_chunk0 of its first piece, ... of tile k+1's first piece).
      const int T = L->cu, B = L->blocks;
      tiles->n_tiles = T;
      tiles->items.clear();
      tiles->cta_first.assign((size_t)T * (n_cta + 1), 0);
      tiles->rows_end.assign(T, L->rows);
      std::vector<uint64_t> tile_chunk0(T + 1, L->n_chunks);
      for (int k = T - 1; k >= 0; k--) tile_chunk0[k] = L->piece_chunk0[(size_t)0 * T + k];  // piece (block 0, CU k) opens tile k
      size_t i0 = 0;
      for (int k = 0; k < T; k++) {
        size_t i1 = i0;
        while (i1 < items.size() && items[i1].chunk_begin < tile_chunk0[k + 1]) i1++;
        uint32_t *first = &tiles->cta_first[(size_t)k * (n_cta + 1)];
        for (int j = 0; j < n_cta; j++) {
          first[j] = (uint32_t)tiles->items.size();
          for (size_t i = i0 + (size_t)j; i < i1; i += (size_t)n_cta) tiles->items.push_back(items[i]);
        }
        first[n_cta] = (uint32_t)tiles->items.size();
        i0 = i1;
      }
      // rows that are final once tile k is done: everything below the first row of any later tile's piece (the CU
      // split is made per column block, csr_hw.cpp:459-468, so the tiles' row ranges differ a little from block to block)
      std::vector<uint32_t> first_row(T, L->rows);
      for (int b = 0; b < B; b++)
        for (int k = 0; k < T; k++) {
          const size_t bk = (size_t)b * T + k;
          if (L->piece_chunk1[bk] == L->piece_chunk0[bk]) continue;
          const ChunkMeta &m = L->chunks[L->piece_chunk0[bk]];
          if (m.valid & 0x3FFu) first_row[k] = std::min(first_row[k], m.row_first);
        }
      uint32_t later = L->rows;
      for (int k = T - 1; k >= 0; k--) {
        tiles->rows_end[k] = later;          // min first row over the tiles after k
        later = std::min(later, first_row[k]);
      }
      for (int k = 1; k < T; k++) tiles->rows_end[k] = std::max(tiles->rows_end[k], tiles->rows_end[k - 1]);
    }
    std::vector<XsItem> rr;
    rr.reserve(items.size());
    for (int j = 0; j < n_cta; j++) {
      cta_first[j] = (uint32_t)rr.size();
      for (size_t i = (size_t)j; i < items.size(); i += (size_t)n_cta) rr.push_back(items[i]);
    }
    cta_first[n_cta] = (uint32_t)rr.size();
    items.swap(rr);
  }
}

}  // namespace spmvb

using namespace spmvb;

extern "C" int64_t spmvb_layout_xs_plan(const spmvb_layout *l, int n_cta, int run_log2, uint32_t *items_out,
                                        uint64_t max_items, uint32_t *cta_first_out) {
  const Layout *L = (const Layout *)l;
  if (!L || n_cta < 1 || run_log2 < 1 || run_log2 < L->run_log2 || run_log2 > 8) return fail(SPMVB_E_ARG, "xs_plan");
  if (L->dev) L = L->dev;  // the plan is made for what the GPU streams
  std::vector<XsItem> items;
  std::vector<uint32_t> cta_first;
  build_xs_items(L, n_cta, (uint32_t)run_log2, items, cta_first, nullptr);
  if (items_out) memcpy(items_out, items.data(), std::min<uint64_t>(items.size(), max_items) * sizeof(XsItem));
  if (cta_first_out) memcpy(cta_first_out, cta_first.data(), cta_first.size() * 4);
  return (int64_t)items.size();
}
