// GPU builder of the hw_matrix layout: the steps of create_csr_hw_matrix as data-parallel
// passes over a device-resident CSR, producing the engine's device image directly (no host round trip).
//
// Same result, bit for bit, as the host builder (layout_builder.cpp) and therefore as the reference
// (scan_matrix csr_hw.cpp:7-146, prepare_balanced_hw_matrix :327-361/:432-484, hw_matrix_alloc :151-183,
// create_block_matrix :190-265, generate_balanced_hw_submatrix :270-318), but formulated for a GPU:
//
//   1. (row, block) pairs = maximal runs of one row's entries inside one column block.  Heads are flagged per entry,
//      an inclusive scan numbers the pairs in CSR order.
//   2. A stable radix sort of the pairs by block gives every pair its rank in the row map (the compact
//      empty_rows_bitmap); an exclusive scan of the VF-padded run lengths in rank order gives its position in the block.
//   3. One thread per block replays the reference's sequential S1 && S2 && S3 split rule over the block's padded
//      lengths (CU > 1), the host turns the per-piece sizes into offsets (layout_finish_pieces, shared code).
//   4. Every entry is written straight to its final byte of the device image; per-chunk metadata, the rows to clear
//      and the column ranges follow from one pass over the chunks.
//
// Rows whose column blocks do not ascend (unsorted columns) take one extra pass: a stable radix sort of all entries by
// (row, block), which keeps the given order inside every pair exactly like the reference's per-block walk of the row.
//
// The step bodies are plain functions of (index, context) so that tests can run the very same code on the CPU
// (tests/emu/layout_emu.cpp: serial loops in reverse order, std::stable_sort, std::partial_sum) and compare it with
// the host builder without a GPU.  The product only ever runs them as CUDA kernels (layout_gpu.cuh).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "layout.h"

#if defined(__CUDACC__)
#define SPMVB_HD __host__ __device__ __forceinline__
#else
#define SPMVB_HD inline
#endif

namespace spmvb {

enum LbError : uint32_t {
  kLbErrColRange = 1,       // column index >= cols
  kLbErrBlockOrder = 2,     // column blocks do not ascend inside a row
  kLbErrBlockOverflow = 4,  // a block's padded nnz overflows IndexType
  kLbErrRowPtr = 8,         // row_ptr not monotone
};

struct LbCtx {
  // input (device)
  const uint64_t *row_ptr;
  const uint32_t *col_ind;
  const uint8_t *values;
  uint32_t rows, cols, cdb;
  int cdb_shift;  // >= 0: cols_div_blocks is 1 << cdb_shift
  int blocks, cu, vf, vb, ratio_v, gb, chunk_bytes, slot, run_log2;
  uint64_t nnz, n_pairs, n_chunks, n_pieces;
  // reorder pass (rows whose column blocks do not ascend)
  uint64_t *okey;    // [nnz] row << block_bits | block
  uint32_t *oidx;    // [nnz] identity, the sort's payload
  uint32_t *perm;    // [nnz] sorted position -> original entry
  uint32_t *col2;    // [nnz] reordered copies of col_ind / values
  uint8_t *val2;
  const uint32_t *col_in;
  const uint8_t *val_in;
  int block_bits;
  // per entry
  uint8_t *head;    // [nnz] bit 0 = first entry of a (row, block) pair, bit 7 = ... but not of its row
  uint32_t *pincl;  // [nnz] inclusive scan of head: pair number + 1 (CSR order)
  // per pair, CSR order
  uint64_t *pj;     // [n_pairs + 1] first entry
  uint32_t *pkey;   // column block
  uint32_t *pval;   // identity, the radix sort's payload
  uint32_t *prow;   // row
  uint8_t *pmid;    // the pair is not the first of its row
  uint32_t *rank_of;
  // per pair, rank order (= row map order: by block, rows ascending)
  uint32_t *pkey_sorted, *order;
  uint32_t *plen;    // [n_pairs + 1] VF-padded length
  uint64_t *gpos;    // [n_pairs + 1] exclusive scan of plen
  uint8_t *nonsole;  // [n_pairs + 1] the pair's row has entries in other blocks too
  uint32_t *nsp;     // [n_pairs + 1] exclusive scan of nonsole
  uint32_t *rowmap;  // OUTPUT rank -> row
  // per block / piece
  uint64_t *rank_base;  // [blocks + 1]
  uint64_t *fp;         // [blocks * (cu + 1)] first block position of every piece
  uint32_t *nr_rows, *nr_nzeros;  // [k * blocks + b]
  uint32_t *pad_rows;             // [blocks]
  uint64_t *piece_chunk0, *piece_last_rank;  // [b * cu + k]
  uint32_t *piece_real;                      // [b * cu + k]
  uint64_t *ord_chunk0;                      // [n_pieces + 1] chunk0 of the pieces in device order, then n_chunks
  uint32_t *dev_order;                       // [n_pieces]
  // per chunk
  uint32_t *crank0;  // rank of the segment holding the chunk's first entry
  uint8_t *cmid;     // the chunk's first entry continues a row
  uint8_t *image;    // OUTPUT n_chunks slots of chunk_bytes + 16
  ChunkMeta *metas;  // OUTPUT compact copy of the slot metadata
  uint16_t *col_lo, *col_hi;  // OUTPUT
  uint16_t *x_lines;          // OUTPUT distinct 128-byte lines of x per chunk
  uint8_t *needz;    // [rows] OUTPUT row must be cleared before y = A x
  uint32_t *err;     // OR of LbError
};

SPMVB_HD void lb_flag(uint32_t *err, uint32_t bit) {
#if defined(__CUDA_ARCH__)
  atomicOr(err, bit);
#else
  *err |= bit;
#endif
}
SPMVB_HD uint32_t lb_block_of(const LbCtx &c, uint32_t col) { return c.cdb_shift >= 0 ? col >> c.cdb_shift : col / c.cdb; }
SPMVB_HD uint32_t lb_round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }
// CU piece holding block position pos: the number of split points f[1..cu-1] that are <= pos
SPMVB_HD int lb_piece_of(const uint64_t *f, int cu, uint64_t pos) {
  int lo = 0, hi = cu - 1;  // answer in [0, cu-1]
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (f[mid] <= pos) lo = mid; else hi = mid - 1;
  }
  return lo;
}
SPMVB_HD uint8_t *lb_entry_addr(const LbCtx &c, uint64_t bk, uint64_t e) {
  const uint64_t chunk = c.piece_chunk0[bk] + e / kChunkEntries;
  const uint32_t w = (uint32_t)(e % kChunkEntries);
  return c.image + chunk * (uint64_t)c.slot + (uint64_t)(w / kRatioCi) * c.gb;
}
SPMVB_HD void lb_write_ci(const LbCtx &c, uint64_t bk, uint64_t e, uint16_t ci) {
  *reinterpret_cast<uint16_t *>(lb_entry_addr(c, bk, e) + 2 * (e % kRatioCi)) = ci;
}

// head[row_ptr[r]] = 1 for every non-empty row
struct LbRowHeads {
  static SPMVB_HD void run(uint64_t r, const LbCtx &c) {
    const uint64_t a = c.row_ptr[r], b = c.row_ptr[r + 1];
    if (b < a || b > c.nnz || (r == 0 && a != 0)) { lb_flag(c.err, kLbErrRowPtr); return; }
    if (a < b) c.head[a] = 1;
  }
};

// head[j] = 1 where the column block changes inside a row; input checks
struct LbEntryHeads {
  static SPMVB_HD void run(uint64_t j, const LbCtx &c) {
    const uint32_t col = c.col_ind[j];
    if (col >= c.cols) { lb_flag(c.err, kLbErrColRange); return; }
    if (c.head[j] || j == 0) return;
    const uint32_t prev = c.col_ind[j - 1];
    if (prev >= c.cols) return;
    const uint32_t b = lb_block_of(c, col), bp = lb_block_of(c, prev);
    if (b != bp) c.head[j] = 0x81;
    if (b < bp) lb_flag(c.err, kLbErrBlockOrder);
  }
};

// reorder pass: sort key of every entry = (row, column block); a stable sort keeps the given order inside a pair
struct LbEntryKeys {
  static SPMVB_HD void run(uint64_t j, const LbCtx &c) {
    uint64_t lo = 0, hi = c.rows;  // first index with row_ptr[i] > j, in (0, rows]
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (c.row_ptr[mid] > j) hi = mid; else lo = mid + 1;
    }
    c.okey[j] = ((lo - 1) << c.block_bits) | lb_block_of(c, c.col_in[j]);
    c.oidx[j] = (uint32_t)j;
  }
};
struct LbPermute {
  static SPMVB_HD void run(uint64_t j, const LbCtx &c) {
    const uint64_t src = c.perm[j];
    c.col2[j] = c.col_in[src];
    if (c.vb == 8) reinterpret_cast<uint64_t *>(c.val2)[j] = reinterpret_cast<const uint64_t *>(c.val_in)[src];
    else reinterpret_cast<uint32_t *>(c.val2)[j] = reinterpret_cast<const uint32_t *>(c.val_in)[src];
  }
};

// per pair (CSR order): first entry, block
struct LbPairHeads {
  static SPMVB_HD void run(uint64_t j, const LbCtx &c) {
    if (j == 0) c.pj[c.n_pairs] = c.nnz;
    const uint8_t h = c.head[j];
    if (!h) return;
    const uint32_t p = c.pincl[j] - 1;
    c.pj[p] = j;
    c.pkey[p] = lb_block_of(c, c.col_ind[j]);
    c.pval[p] = p;
    c.pmid[p] = h >> 7;
  }
};

// row of the first pair of every non-empty row ...
struct LbRowPairs {
  static SPMVB_HD void run(uint64_t r, const LbCtx &c) {
    const uint64_t a = c.row_ptr[r];
    if (a < c.row_ptr[r + 1]) c.prow[c.pincl[a] - 1] = (uint32_t)r;
  }
};
// ... and of the pairs that follow it in the same row (at most one per column block)
struct LbMidPairs {
  static SPMVB_HD void run(uint64_t p, const LbCtx &c) {
    if (!c.pmid[p]) return;
    uint64_t q = p - 1;
    while (c.pmid[q]) q--;
    c.prow[p] = c.prow[q];
  }
};

// per pair (rank order): row map, padded length, block boundaries of the row map, sole-block rows
struct LbRanks {
  static SPMVB_HD void run(uint64_t i, const LbCtx &c) {
    const uint32_t p = c.order[i];
    const uint32_t row = c.prow[p];
    c.rowmap[i] = row;
    c.needz[row] = 0;  // the row has entries: no clearing unless a later step asks for it
    c.plen[i] = lb_round_up((uint32_t)(c.pj[p + 1] - c.pj[p]), (uint32_t)c.vf);
    c.rank_of[p] = (uint32_t)i;
    c.nonsole[i] = (p > 0 && c.prow[p - 1] == row) || ((uint64_t)p + 1 < c.n_pairs && c.prow[p + 1] == row);
    const int64_t key = c.pkey_sorted[i];
    const int64_t prev = i ? (int64_t)c.pkey_sorted[i - 1] : -1;
    for (int64_t b = prev + 1; b <= key; b++) c.rank_base[b] = i;
    if (i == c.n_pairs - 1) {
      for (int64_t b = key + 1; b <= c.blocks; b++) c.rank_base[b] = c.n_pairs;
      c.plen[c.n_pairs] = 0;
      c.nonsole[c.n_pairs] = 0;
    }
  }
};

// prepare_balanced_hw_matrix: the S1 && S2 && S3 split of one block (csr_hw.cpp:459-468), leftovers and row padding
// to the last CU (:474-482).  Sequential by definition (each split restarts the counters), one thread per block.
struct LbSplit {
  static SPMVB_HD void run(uint64_t b, const LbCtx &c) {
    uint64_t *f = c.fp + b * (uint64_t)(c.cu + 1);
    const uint64_t base = c.rank_base[b], P = c.rank_base[b + 1] - base;
    const uint64_t Z = c.gpos[base + P] - c.gpos[base];
    if (Z + (uint64_t)c.ratio_v * c.vf > 0xFFFFFFFFull) lb_flag(c.err, kLbErrBlockOverflow);
    uint64_t nz = 0, rc = 0, done_rows = 0;
    int fired = 0;
    f[0] = 0;
    if (c.cu > 1) {
      // The reference walks the rows once, firing a split at the first row where the running padded count exceeds
      // thr and both counters are multiples of RATIO_v, then restarts the counters.  With the prefix sums at hand the
      // first row past thr is a binary search; only the wait for the two divisibility conditions is a walk.
      const uint64_t thr = Z / (uint64_t)c.cu;
      const uint64_t *g = c.gpos;
      const uint64_t end = base + P;
      uint64_t s = base;  // first rank of the piece being filled
      while (fired < c.cu - 1 && s < end) {
        uint64_t lo = s, hi = end;  // first i in [s, end) with g[i+1] - g[s] > thr, else end
        while (lo < hi) {
          const uint64_t mid = (lo + hi) >> 1;
          if (g[mid + 1] - g[s] > thr) hi = mid; else lo = mid + 1;
        }
        uint64_t i = lo;
        for (; i < end; i++)
          if ((g[i + 1] - g[s]) % c.ratio_v == 0 && (i + 1 - s) % c.ratio_v == 0) break;
        if (i >= end) break;
        c.nr_rows[(uint64_t)fired * c.blocks + b] = (uint32_t)(i + 1 - s);
        c.nr_nzeros[(uint64_t)fired * c.blocks + b] = (uint32_t)(g[i + 1] - g[s]);
        fired++;
        f[fired] = g[i + 1] - g[base];
        done_rows += i + 1 - s;
        s = i + 1;
      }
      // what is left after the last split (or everything) belongs to the last CU
      nz = Z - f[fired];
      rc = P - done_rows;
    } else {
      nz = Z; rc = P;
    }
    for (int k = fired + 1; k < c.cu; k++) f[k] = f[fired];  // un-fired middle CUs stay empty (reference: garbage, Q2)
    f[c.cu] = Z;
    const uint32_t mod = (uint32_t)(rc % c.ratio_v);
    uint32_t pad = 0;
    if (mod) { pad = c.ratio_v - mod; rc += pad; nz += (uint64_t)pad * c.vf; }
    c.pad_rows[b] = pad;
    c.nr_rows[(uint64_t)(c.cu - 1) * c.blocks + b] = (uint32_t)rc;
    c.nr_nzeros[(uint64_t)(c.cu - 1) * c.blocks + b] = (uint32_t)nz;
  }
};

// create_block_matrix + generate_balanced_hw_submatrix: every entry to its final byte
struct LbScatter {
  static SPMVB_HD void run(uint64_t j, const LbCtx &c) {
    const uint32_t p = c.pincl[j] - 1;
    const uint64_t i = c.rank_of[p];
    const uint32_t b = c.pkey[p];
    const uint64_t posb = c.gpos[i] - c.gpos[c.rank_base[b]];
    const uint64_t *f = c.fp + (uint64_t)b * (c.cu + 1);
    const int k = lb_piece_of(f, c.cu, posb);
    const uint64_t bk = (uint64_t)b * c.cu + k;
    const uint64_t j0 = c.pj[p], len = c.pj[p + 1] - j0;
    const uint64_t s_e = posb - f[k];
    const uint64_t e = s_e + (j - j0);
    const bool last = j == j0 + len - 1;
    const uint32_t col = c.col_ind[j];
    const uint16_t ci = (uint16_t)((col - b * c.cdb) | ((last && len % c.vf == 0) ? 0x8000u : 0u));  // csr_hw.cpp:220, :288-292
    uint8_t *grp = lb_entry_addr(c, bk, e);
    const uint32_t s = (uint32_t)(e % kRatioCi);
    *reinterpret_cast<uint16_t *>(grp + 2 * s) = ci;
    if (c.vb == 8)  // csr_hw.cpp:300-310
      *reinterpret_cast<uint64_t *>(grp + kBusBytes + 8 * s) = reinterpret_cast<const uint64_t *>(c.values)[j];
    else
      *reinterpret_cast<uint32_t *>(grp + kBusBytes + 4 * s) = reinterpret_cast<const uint32_t *>(c.values)[j];
    if (last && len % c.vf != 0)  // VF padding: (col 0, val 0), end-of-row bit on the last slot, csr_hw.cpp:229-238
      lb_write_ci(c, bk, s_e + c.plen[i] - 1, 0x8000u);
  }
};

// chunks whose first entry lies inside this pair's segment start at its rank
struct LbPairChunks {
  static SPMVB_HD void run(uint64_t i, const LbCtx &c) {
    const uint32_t b = c.pkey_sorted[i];
    const uint64_t posb = c.gpos[i] - c.gpos[c.rank_base[b]];
    const uint64_t *f = c.fp + (uint64_t)b * (c.cu + 1);
    const int k = lb_piece_of(f, c.cu, posb);
    const uint64_t bk = (uint64_t)b * c.cu + k;
    const uint64_t s_e = posb - f[k], t_e = s_e + c.plen[i];
    for (uint64_t q = (s_e + kChunkEntries - 1) / kChunkEntries; q * kChunkEntries < t_e; q++) {
      const uint64_t ch = c.piece_chunk0[bk] + q;
      c.crank0[ch] = (uint32_t)i;
      c.cmid[ch] = q * kChunkEntries > s_e;
    }
  }
};

// padding rows of the last CU: VF x (col 0, val 0) with the end-of-row bit on the last (csr_hw.cpp:246-255)
struct LbPadRows {
  static SPMVB_HD void run(uint64_t b, const LbCtx &c) {
    const uint64_t bk = b * (uint64_t)c.cu + (c.cu - 1);
    const uint64_t real = c.piece_real[bk];
    for (uint32_t i = 0; i < c.pad_rows[b]; i++) lb_write_ci(c, bk, real + (uint64_t)i * c.vf + (c.vf - 1), 0x8000u);
  }
};

// per-chunk metadata, column range, and the rows that need clearing before a SpMV
struct LbChunks {
  static SPMVB_HD void run(uint64_t ch, const LbCtx &c) {
    uint64_t lo = 0, hi = c.n_pieces;  // last piece (device order) with chunk0 <= ch
    while (lo < hi) {
      const uint64_t mid = (lo + hi + 1) >> 1;
      if (c.ord_chunk0[mid] <= ch) lo = mid; else hi = mid - 1;
    }
    const uint64_t bk = c.dev_order[lo];
    const uint32_t b = (uint32_t)(bk / c.cu);
    const uint64_t c0 = c.piece_chunk0[bk], c1 = c.ord_chunk0[lo + 1];
    const uint64_t real = c.piece_real[bk];
    const uint64_t first = (ch - c0) * kChunkEntries;
    const uint32_t valid = first >= real ? 0u : (uint32_t)(real - first < kChunkEntries ? real - first : kChunkEntries);
    ChunkMeta m;
    m.rank0 = valid ? c.crank0[ch] : 0u;
    m.block = b;
    m.valid = valid;
    m.row_first = 0;
    uint16_t clo = 0xFFFF, chi = 0;
    uint32_t lines = 0;
    if (valid) {
      if (c.cmid[ch]) m.valid |= kChunkStartsMid;
      m.row_first = c.rowmap[m.rank0];
      const bool next_valid = ch + 1 < c1 && (ch + 1 - c0) * kChunkEntries < real;
      const uint64_t last_rank = next_valid ? c.crank0[ch + 1] : c.piece_last_rank[bk];
      if ((uint64_t)c.rowmap[last_rank] - m.row_first == last_rank - m.rank0) m.valid |= kChunkRowsConsecutive;
      const bool sole = c.nsp[last_rank + 1] - c.nsp[m.rank0] == 0;
      if (sole) m.valid |= kChunkSole;
      // rows to clear: a row split across two runs (runs are aligned globally or to the piece start), and every row
      // with a segment (or part of one) in an atomics-only chunk
      const uint64_t R = 1ull << c.run_log2;
      if (c.cmid[ch] && ((ch % R) == 0 || ((ch - c0) % R) == 0)) c.needz[m.row_first] = 1;
      if (!sole)
        for (uint64_t rk = m.rank0; rk <= last_rank; rk++) c.needz[c.rowmap[rk]] = 1;
      const uint8_t *base = c.image + ch * (uint64_t)c.slot;
      uint32_t row_ends = 0;
      uint32_t seen[64];
      for (int i = 0; i < 64; i++) seen[i] = 0;
      const int line_shift = c.vb == 8 ? 4 : 5;
      for (uint32_t g = 0; g * kRatioCi < valid; g++) {
        const uint16_t *w = reinterpret_cast<const uint16_t *>(base + (uint64_t)g * c.gb);
        const uint32_t n = valid - g * kRatioCi < (uint32_t)kRatioCi ? valid - g * kRatioCi : (uint32_t)kRatioCi;
        for (uint32_t s = 0; s < n; s++) {
          uint16_t ci = w[s];
          row_ends += ci >> 15;
          ci &= 0x7FFF;
          if (ci < clo) clo = ci;
          if (ci > chi) chi = ci;
          const uint32_t ln = (uint32_t)ci >> line_shift, bit = 1u << (ln & 31);
          lines += !(seen[ln >> 5] & bit);
          seen[ln >> 5] |= bit;
        }
      }
      m.block |= row_ends << kMetaRowsShift;
    }
    c.col_lo[ch] = clo; c.col_hi[ch] = chi; c.x_lines[ch] = (uint16_t)lines;
    c.metas[ch] = m;
    *reinterpret_cast<ChunkMeta *>(c.image + ch * (uint64_t)c.slot + c.chunk_bytes) = m;
  }
};

// What the builder hands to the engine: device buffers (owned by the caller from then on).
struct LbImage {
  uint8_t *image = nullptr;      // n_chunks slots
  uint32_t *rowmap = nullptr;    // n_pairs + 1
  uint32_t *zero_rows = nullptr; // n_zero
  uint64_t n_zero = 0;
};

inline std::string lb_error_text(uint32_t err) {
  std::string s;
  if (err & kLbErrColRange) s += "column index out of range; ";
  if (err & kLbErrBlockOrder) s += "column blocks do not ascend inside a row; ";
  if (err & kLbErrBlockOverflow) s += "block nnz overflows IndexType; ";
  if (err & kLbErrRowPtr) s += "row_ptr must start at 0 and be non-decreasing; ";
  return s;
}

// The builder proper.  BE supplies memory (alloc/release: temporaries, alloc_output/release_output: what the engine
// keeps; reserve: a hint of how much alloc() will hand out next), launches, scans, the sort and the compaction
// (CUDA: layout_gpu.cuh).
// On success *out is a Layout with every host-side table filled except the two big arrays, stream and rowmap
// (lb_fetch_host brings them over on demand), and img holds the device buffers.
template <class BE>
int lb_build(BE &be, uint32_t rows, uint32_t cols, uint64_t nnz, const uint64_t *d_row_ptr, const uint32_t *d_col_ind,
             const void *d_values, int cu, int vf, int is_double, uint32_t cdb_in, Layout **out, LbImage *img) {
  *out = nullptr;
  if (nnz >= 0x7FFFFFFFull || rows >= 0x7FFFFFFFu)
    return fail(SPMVB_E_RANGE, "the GPU builder takes fewer than 2^31 rows and non-zeros per engine");
  Layout *L = new Layout();
  int rc = layout_init_header(L, rows, cols, nnz, cu, vf, is_double, cdb_in);
  if (rc) { delete L; return rc; }
  const int blocks = L->blocks;
  const uint64_t KB = (uint64_t)cu * blocks;

  LbCtx c;
  memset(&c, 0, sizeof(c));
  c.row_ptr = d_row_ptr; c.col_ind = d_col_ind; c.values = (const uint8_t *)d_values;
  c.rows = rows; c.cols = cols; c.cdb = L->cdb;
  c.cdb_shift = -1;
  if ((L->cdb & (L->cdb - 1)) == 0) { c.cdb_shift = 0; while ((1u << c.cdb_shift) != L->cdb) c.cdb_shift++; }
  c.blocks = blocks; c.cu = cu; c.vf = vf; c.vb = L->vb; c.ratio_v = L->ratio_v; c.gb = L->group_bytes;
  c.chunk_bytes = L->chunk_bytes; c.slot = L->chunk_bytes + (int)sizeof(ChunkMeta); c.run_log2 = L->run_log2;
  c.nnz = nnz; c.n_pieces = KB;

  std::vector<void *> temps;  // freed on every exit path
  auto T = [&](void *p) { temps.push_back(p); return p; };
  auto cleanup = [&](bool keep_outputs) {
    for (void *p : temps) be.release(p);
    if (!keep_outputs) {
      be.release_output(img->image); be.release_output(img->rowmap); be.release_output(img->zero_rows);
      *img = LbImage();
    }
  };
  auto bail = [&](int code, const std::string &msg) { cleanup(false); delete L; return fail(code, msg); };
#define LB_CHECK() do { if (!be.ok()) return bail(be.code(), be.error()); } while (0)

  // temporaries come out of three arenas (per entry, per pair, per chunk): one device allocation each
  const uint64_t kSlack = (uint64_t)4 << 20;
  be.reserve((uint64_t)rows + (uint64_t)(blocks + 1) * 8 + nnz * 5 + kSlack);
  c.err = (uint32_t *)T(be.alloc(4)); be.fill(c.err, 0, 4);
  c.needz = (uint8_t *)T(be.alloc(rows)); be.fill(c.needz, 1, rows);
  c.rank_base = (uint64_t *)T(be.alloc((size_t)(blocks + 1) * 8)); be.fill(c.rank_base, 0, (size_t)(blocks + 1) * 8);
  LB_CHECK();
  be.trace("arena 1");

  // ---- 1. pairs
  uint32_t n_pairs32 = 0;
  if (nnz) {
    c.head = (uint8_t *)T(be.alloc(nnz)); be.fill(c.head, 0, nnz);
    c.pincl = (uint32_t *)T(be.alloc(nnz * 4));
    LB_CHECK();
    be.template launch<LbRowHeads>(rows, c);
    be.template launch<LbEntryHeads>(nnz, c);
    be.inclusive_sum_bit0_u32(c.head, c.pincl, nnz);
    be.to_host(&n_pairs32, c.pincl + (nnz - 1), 4);
    LB_CHECK();
  } else {
    be.template launch<LbRowHeads>(rows, c);  // row_ptr sanity only
  }
  uint32_t err = 0;
  be.to_host(&err, c.err, 4);
  LB_CHECK();
  if (err == kLbErrBlockOrder) {
    // Some row visits its column blocks out of order (unsorted columns).  The layout keeps, inside every (row, block)
    // pair, the order the entries were given in (create_block_matrix walks the row once per block,
    // csr_hw.cpp:212-226), so: stable sort of all entries by (row, block), then the passes above on the reordered copy.
    int row_bits = 1;
    while ((1ull << row_bits) < (uint64_t)rows) row_bits++;
    c.block_bits = 1;
    while ((1ull << c.block_bits) < (uint64_t)blocks) c.block_bits++;
    be.reserve(nnz * (uint64_t)(44 + c.vb) + kSlack);
    c.okey = (uint64_t *)T(be.alloc(nnz * 8));
    uint64_t *okey_sorted = (uint64_t *)T(be.alloc(nnz * 8));
    c.oidx = (uint32_t *)T(be.alloc(nnz * 4));
    c.perm = (uint32_t *)T(be.alloc(nnz * 4));
    c.col2 = (uint32_t *)T(be.alloc(nnz * 4));
    c.val2 = (uint8_t *)T(be.alloc(nnz * (uint64_t)c.vb));
    LB_CHECK();
    c.col_in = c.col_ind; c.val_in = c.values;
    be.template launch<LbEntryKeys>(nnz, c);
    be.sort_pairs_u64(c.okey, okey_sorted, c.oidx, c.perm, nnz, c.block_bits + row_bits);
    be.template launch<LbPermute>(nnz, c);
    c.col_ind = c.col2; c.values = c.val2;
    be.fill(c.head, 0, nnz);
    be.fill(c.err, 0, 4);
    be.template launch<LbRowHeads>(rows, c);
    be.template launch<LbEntryHeads>(nnz, c);
    be.inclusive_sum_bit0_u32(c.head, c.pincl, nnz);
    be.to_host(&n_pairs32, c.pincl + (nnz - 1), 4);
    be.to_host(&err, c.err, 4);
    LB_CHECK();
    be.trace("reorder rows by column block");
  }
  if (err) return bail(err & kLbErrBlockOverflow ? SPMVB_E_RANGE : SPMVB_E_ARG, "GPU layout build: " + lb_error_text(err));
  be.trace("pair heads + scan");
  const uint64_t n_pairs = n_pairs32;
  c.n_pairs = n_pairs;
  L->n_pairs = n_pairs;
  if (n_pairs >= 0x7FFFFFF0ull) return bail(SPMVB_E_RANGE, "too many (row, block) pairs for the GPU builder");

  be.reserve((n_pairs + 1) * 64 + (uint64_t)blocks * (cu + 1) * 8 + KB * 48 + kSlack);
  c.rowmap = (uint32_t *)be.alloc_output((n_pairs + 8) * 4);  // slack: the kernels copy 16-byte aligned slices
  img->rowmap = c.rowmap;
  c.plen = (uint32_t *)T(be.alloc((n_pairs + 1) * 4));
  c.gpos = (uint64_t *)T(be.alloc((n_pairs + 1) * 8));
  c.nonsole = (uint8_t *)T(be.alloc(n_pairs + 1));
  c.nsp = (uint32_t *)T(be.alloc((n_pairs + 1) * 4));
  LB_CHECK();
  be.fill(c.plen, 0, (n_pairs + 1) * 4);
  be.fill(c.nonsole, 0, n_pairs + 1);
  if (n_pairs) {
    c.pj = (uint64_t *)T(be.alloc((n_pairs + 1) * 8));
    c.pkey = (uint32_t *)T(be.alloc(n_pairs * 4));
    c.pval = (uint32_t *)T(be.alloc(n_pairs * 4));
    c.prow = (uint32_t *)T(be.alloc(n_pairs * 4));
    c.pmid = (uint8_t *)T(be.alloc(n_pairs));
    c.rank_of = (uint32_t *)T(be.alloc(n_pairs * 4));
    c.pkey_sorted = (uint32_t *)T(be.alloc(n_pairs * 4));
    c.order = (uint32_t *)T(be.alloc(n_pairs * 4));
    LB_CHECK();
    be.template launch<LbPairHeads>(nnz, c);
    be.template launch<LbRowPairs>(rows, c);
    be.template launch<LbMidPairs>(n_pairs, c);
    // ---- 2. rank order: stable sort by block
    int bits = 1;
    while ((1ull << bits) < (uint64_t)blocks) bits++;
    be.sort_pairs(c.pkey, c.pkey_sorted, c.pval, c.order, n_pairs, bits);
    be.template launch<LbRanks>(n_pairs, c);
  }
  be.trace("pairs, sort, ranks");
  be.exclusive_sum_u32_u64(c.plen, c.gpos, n_pairs + 1);
  be.exclusive_sum_u8_u32(c.nonsole, c.nsp, n_pairs + 1);
  LB_CHECK();

  // ---- 3. split per block, piece tables on the host
  c.fp = (uint64_t *)T(be.alloc((size_t)blocks * (cu + 1) * 8));
  c.nr_rows = (uint32_t *)T(be.alloc(KB * 4)); c.nr_nzeros = (uint32_t *)T(be.alloc(KB * 4));
  c.pad_rows = (uint32_t *)T(be.alloc((size_t)blocks * 4));
  LB_CHECK();
  be.fill(c.nr_rows, 0, KB * 4); be.fill(c.nr_nzeros, 0, KB * 4);
  be.template launch<LbSplit>((uint64_t)blocks, c);
  std::vector<uint64_t> fp((size_t)blocks * (cu + 1));
  std::vector<uint32_t> pad_rows(blocks);
  L->rank_base.assign(blocks + 1, 0);
  L->nr_rows.assign(KB, 0); L->nr_nzeros.assign(KB, 0);
  be.to_host(L->rank_base.data(), c.rank_base, (size_t)(blocks + 1) * 8);
  be.to_host(fp.data(), c.fp, fp.size() * 8);
  be.to_host(pad_rows.data(), c.pad_rows, pad_rows.size() * 4);
  be.to_host(L->nr_rows.data(), c.nr_rows, KB * 4);
  be.to_host(L->nr_nzeros.data(), c.nr_nzeros, KB * 4);
  be.to_host(&err, c.err, 4);
  LB_CHECK();
  if (err) return bail(err & kLbErrBlockOverflow ? SPMVB_E_RANGE : SPMVB_E_ARG, "GPU layout build: " + lb_error_text(err));
  be.trace("scans + split + tables to host");
  std::vector<uint64_t> piece_last_rank;
  layout_finish_pieces(L, fp.data(), pad_rows.data(), piece_last_rank);
  if (L->n_chunks >= 0x7FFFFFFFull) return bail(SPMVB_E_RANGE, "too many chunks for one engine");
  c.n_chunks = L->n_chunks;
  std::vector<uint64_t> ord_chunk0(KB + 1);
  for (uint64_t i = 0; i < KB; i++) ord_chunk0[i] = L->piece_chunk0[L->dev_order[i]];
  ord_chunk0[KB] = L->n_chunks;
  c.piece_chunk0 = (uint64_t *)T(be.alloc(KB * 8)); c.piece_last_rank = (uint64_t *)T(be.alloc(KB * 8));
  c.piece_real = (uint32_t *)T(be.alloc(KB * 4)); c.dev_order = (uint32_t *)T(be.alloc(KB * 4));
  c.ord_chunk0 = (uint64_t *)T(be.alloc((KB + 1) * 8));
  LB_CHECK();
  be.to_device(c.piece_chunk0, L->piece_chunk0.data(), KB * 8);
  be.to_device(c.piece_last_rank, piece_last_rank.data(), KB * 8);
  be.to_device(c.piece_real, L->piece_real_nnz.data(), KB * 4);
  be.to_device(c.dev_order, L->dev_order.data(), KB * 4);
  be.to_device(c.ord_chunk0, ord_chunk0.data(), (KB + 1) * 8);

  // ---- 4. the image
  const uint64_t n_chunks = L->n_chunks;
  const size_t image_bytes = (size_t)(n_chunks * (uint64_t)c.slot > 16 ? n_chunks * (uint64_t)c.slot : 16);
  be.reserve((n_chunks + 1) * 36 + (uint64_t)rows / 2 + kSlack);
  c.image = (uint8_t *)be.alloc_output(image_bytes);
  img->image = c.image;
  c.crank0 = (uint32_t *)T(be.alloc((n_chunks + 1) * 4)); c.cmid = (uint8_t *)T(be.alloc(n_chunks + 1));
  c.metas = (ChunkMeta *)T(be.alloc((n_chunks + 1) * sizeof(ChunkMeta)));
  c.col_lo = (uint16_t *)T(be.alloc((n_chunks + 1) * 2)); c.col_hi = (uint16_t *)T(be.alloc((n_chunks + 1) * 2));
  c.x_lines = (uint16_t *)T(be.alloc((n_chunks + 1) * 2));
  LB_CHECK();
  be.trace("piece tables + image alloc");
  be.fill(c.image, 0, image_bytes);
  be.fill(c.crank0, 0, (n_chunks + 1) * 4); be.fill(c.cmid, 0, n_chunks + 1);
  be.template launch<LbScatter>(nnz, c);
  be.template launch<LbPairChunks>(n_pairs, c);
  be.template launch<LbPadRows>((uint64_t)blocks, c);
  be.template launch<LbChunks>(n_chunks, c);
  LB_CHECK();

  be.trace("scatter + chunk metadata");
  // ---- host-side tables of the layout
  L->chunks = (ChunkMeta *)calloc((size_t)(n_chunks ? n_chunks : 1), sizeof(ChunkMeta));
  if (!L->chunks) return bail(SPMVB_E_NOMEM, "chunks");
  L->chunk_col_lo.assign((size_t)n_chunks, 0xFFFF); L->chunk_col_hi.assign((size_t)n_chunks, 0);
  be.to_host(L->chunks, c.metas, n_chunks * sizeof(ChunkMeta));
  be.to_host(L->chunk_col_lo.data(), c.col_lo, n_chunks * 2);
  be.to_host(L->chunk_col_hi.data(), c.col_hi, n_chunks * 2);
  L->chunk_x_lines.assign((size_t)n_chunks, 0);
  be.to_host(L->chunk_x_lines.data(), c.x_lines, n_chunks * 2);
  const uint64_t nz = be.count_nonzero_u8(c.needz, rows);
  LB_CHECK();
  L->zero_all = nz > (uint64_t)rows / 3;
  if (options().zero_all > 0) L->zero_all = true;
  if (!L->zero_all && nz) {
    img->zero_rows = (uint32_t *)be.alloc_output(nz * 4);
    LB_CHECK();
    const uint64_t got = be.select_flagged_iota(c.needz, img->zero_rows, rows);
    LB_CHECK();
    if (got != nz) return bail(SPMVB_E_CUDA, "GPU layout build: zero-row compaction mismatch");
    img->n_zero = nz;
    L->zero_rows.resize(nz);
    be.to_host(L->zero_rows.data(), img->zero_rows, nz * 4);
  }
  LB_CHECK();
#undef LB_CHECK
  be.trace("metadata to host + zero rows");
  cleanup(true);
  *out = L;
  return SPMVB_OK;
}

// The API layout and - when plan_device_params asks for one - the engine-private device layout (Layout::dev), each
// built by lb_build with a backend of its own (their temporaries never live at the same time).  *dev_img stays empty
// when the API image is what the device streams.
template <class BE>
int lb_build_pair(BE &be_api, BE &be_dev, uint32_t rows, uint32_t cols, uint64_t nnz, const uint64_t *d_row_ptr,
                  const uint32_t *d_col_ind, const void *d_values, int cu, int vf, int is_double, uint32_t cdb_in,
                  Layout **out, LbImage *api_img, LbImage *dev_img) {
  int rc = lb_build(be_api, rows, cols, nnz, d_row_ptr, d_col_ind, d_values, cu, vf, is_double, cdb_in, out, api_img);
  if (rc) return rc;
  Layout *L = *out;
  int cu_dev = cu, vf_dev = vf;
  uint32_t cdb_dev = L->cdb;
  if (!plan_device_params(L, &cu_dev, &vf_dev, &cdb_dev)) return SPMVB_OK;
  building_device_layout() = true;
  rc = lb_build(be_dev, rows, cols, nnz, d_row_ptr, d_col_ind, d_values, cu_dev, vf_dev, is_double, cdb_dev, &L->dev, dev_img);
  building_device_layout() = false;
  if (rc) {
    be_api.release_output(api_img->image); be_api.release_output(api_img->rowmap); be_api.release_output(api_img->zero_rows);
    *api_img = LbImage();
    delete L;
    *out = nullptr;
  }
  return rc;
}

// Brings the two big arrays of a GPU-built layout to the host: the pieces (slots stripped of their metadata, i.e.
// hw_matrix[k]->submatrix[b] bit for bit) and the row map.
template <class BE>
int lb_fetch_host(BE &be, Layout *L, const LbImage &img) {
  if (L->stream && L->rowmap) return SPMVB_OK;
  const size_t sb = (size_t)(L->stream_bytes > 16 ? L->stream_bytes : 16);
  uint8_t *stream = (uint8_t *)calloc(sb, 1);
  uint32_t *rowmap = (uint32_t *)malloc((size_t)(L->n_pairs ? L->n_pairs : 1) * 4);
  if (!stream || !rowmap) { free(stream); free(rowmap); return fail(SPMVB_E_NOMEM, "fetch_host"); }
  if (L->n_chunks)
    be.to_host_2d(stream, (size_t)L->chunk_bytes, img.image, (size_t)L->chunk_bytes + sizeof(ChunkMeta),
                  (size_t)L->chunk_bytes, (size_t)L->n_chunks);
  be.to_host(rowmap, img.rowmap, L->n_pairs * 4);
  if (!be.ok()) { free(stream); free(rowmap); return fail(be.code(), be.error()); }
  free(L->stream); free(L->rowmap);
  L->stream = stream; L->rowmap = rowmap;
  return SPMVB_OK;
}

}  // namespace spmvb
