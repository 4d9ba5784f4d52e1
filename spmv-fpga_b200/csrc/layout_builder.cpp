// Host builder of the hw_matrix layout: O(nnz + rows) work, OpenMP-parallel over row ranges,
// bit-exact with the reference's create_csr_hw_matrix wherever the reference is well defined.
//
// What it reproduces (reference = /root/reference/src):
//   scan_matrix                      csr_hw.cpp:7-146    column blocks, expanded_nr_cols, padded row lengths
//   prepare_balanced_hw_matrix       csr_hw.cpp:327-361 (CU=1), 432-484 (CU=2), same rule for 4/8/10/12
//   hw_matrix_alloc                  csr_hw.cpp:151-183  nr_ci / nr_val
//   create_block_matrix              csr_hw.cpp:190-265  per-block rows, rebased columns, VF / row padding
//   generate_balanced_hw_submatrix   csr_hw.cpp:270-318  the 128-bit word packing
// How it differs in method: the reference keeps blocks x (rows+1) prefix tables and re-walks the whole matrix once
// per block; here three passes over the CSR (count, [segment lengths], scatter) with per-thread per-block cursors
// write every entry straight to its final byte, and the empty_rows_bitmap is kept as a rank -> row map.
#include <omp.h>
#include <sys/mman.h>

#include <algorithm>
#include <cctype>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/spmvb.h"
#include "ell.h"
#include "layout.h"

namespace spmvb {

// Large host arrays of the layout (stream, row map): 2 MB-aligned, huge pages requested (the scatter of pass 3 writes
// through thousands of streams at once - one per column block - and would otherwise miss the TLB on every entry and
// fault in 4 KB pages one by one), pages touched by all cores.  Released with free().
void *layout_big_alloc(uint64_t bytes, bool zero) {
  const uint64_t huge = 2ull << 20;
  if (bytes < 8 * huge) return zero ? calloc((size_t)std::max<uint64_t>(bytes, 16), 1) : malloc((size_t)std::max<uint64_t>(bytes, 16));
  const uint64_t rounded = (bytes + huge - 1) / huge * huge;
  uint8_t *p = (uint8_t *)aligned_alloc((size_t)huge, (size_t)rounded);
  if (!p) return nullptr;
  madvise(p, (size_t)rounded, MADV_HUGEPAGE);  // advisory: plain pages if the kernel declines
  const int64_t n = (int64_t)(rounded / huge);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    if (zero) memset(p + (uint64_t)i * huge, 0, (size_t)huge);
    else for (uint64_t o = 0; o < huge; o += 4096) p[(uint64_t)i * huge + o] = 0;
  }
  return p;
}

Layout::~Layout() {
  delete ell;
  free(stream);
  free(rowmap);
  free(chunks);
  delete dev;
  delete wide;
}

bool &building_device_layout() {
  static thread_local bool flag = false;
  return flag;
}

bool &building_wide_layout() {
  static thread_local bool flag = false;
  return flag;
}

Options &options() {
  static Options o;
  return o;
}

namespace {
struct OptionName { const char *name; int64_t Options::*field; };
const OptionName kOptionNames[] = {
    {"run_log2", &Options::run_log2},         {"cu_major", &Options::cu_major},       {"zero_all", &Options::zero_all},
    {"tall", &Options::tall},                 {"occ_run_log2", &Options::occ_run_log2}, {"xs_run_log2", &Options::xs_run_log2},
    {"autotune", &Options::autotune},         {"build_trace", &Options::build_trace}, {"dev_tiles", &Options::dev_tiles},
    {"dev_cdb", &Options::dev_cdb},           {"xs_pairs", &Options::xs_pairs},       {"tile_mb", &Options::tile_mb},
    {"e2e_tiles", &Options::e2e_tiles},       {"xs_config", &Options::xs_config},     {"l2_persist_mb", &Options::l2_persist_mb},
    {"tile_launch", &Options::tile_launch},   {"diag_flags", &Options::diag_flags},   {"wide", &Options::wide},
    {"wide_range_log2", &Options::wide_range_log2}, {"wide_hints", &Options::wide_hints},   {"ell", &Options::ell},
    {"ell_tiles", &Options::ell_tiles},
};
}  // namespace

// The engine-private device layout (see Layout::dev).  Two questions, both answered from the API layout's own tables:
//  1. Is the matrix irregular (layout_is_irregular), so that the kernel with the x window in shared memory is the one
//     to run?
//     Then the window of a column block must fit the kernel's 128 KB: fp64 blocks wider than 16 384 columns are cut
//     in two (the reference's own CU 10/12 block width, util.h:53-58), which is the split on index bit 14.
//  2. Is y larger than the L2 cache can hold next to the stream?  An irregular matrix updates y at random, one
//     red.global per pair: rows are cut into tiles (compute units of the device layout, walked CU-major) so that the
//     y range being updated stays L2-resident, at the price of streaming x once per tile.
// "Irregular": the chunks' entries are scattered over x - on average more than xs_pairs (default 64) distinct 128-byte
// lines per 256-entry chunk (a 5-point Laplacian touches 12, a band 3-4, R-MAT and uniform matrices 150-256).  Gathers
// from global memory then cost ~2 cycles of L1 tag time per entry, and the x window in shared memory is what pays.
bool layout_is_irregular(const Layout *L) {
  const uint64_t thr = options().xs_pairs >= 0 ? (uint64_t)options().xs_pairs : 64;
  uint64_t lines = 0, used = 0;
  for (uint64_t c = 0; c < L->n_chunks; c++)
    if (L->chunk_x_lines[c]) { lines += L->chunk_x_lines[c]; used++; }
  return used > 0 && lines > thr * used;
}

bool plan_device_params(const Layout *L, int *cu_dev, int *vf_dev, uint32_t *cdb_dev) {
  const Options &o = options();
  int cu = L->cu, vf = L->vf;
  uint32_t cdb = L->cdb;
  const bool irregular = layout_is_irregular(L);
  const uint32_t cap = xs_config(L->is_double, L->xs_cfg).cap;
  if (irregular) {
    if (cdb > cap / (uint32_t)L->vb) {  // do the windows of the chunks fit the kernel's?  else: blocks of window width
      uint64_t wide = 0, used = 0;
      for (uint64_t c = 0; c < L->n_chunks; c++)
        if (L->chunk_col_lo[c] <= L->chunk_col_hi[c]) {
          used++;
          wide += (uint64_t)(L->chunk_col_hi[c] - L->chunk_col_lo[c] + 1) * L->vb > cap;
        }
      if (2 * wide > used) cdb = cap / (uint32_t)L->vb;
    }
    const uint64_t ybytes = (uint64_t)L->rows * L->vb;
    // 24 MB: measured on the 1 B-nnz uniform matrix (16 / 20 / 24 / 32 / 48 / 64 MB tiles = 8.9 / 9.0 / 8.3 / 10.7 / 15.7 /
    // 17.1 ms): the L2 keeps a tile's y range next to the stream, the x windows and the row map only up to about there
    const uint64_t tile = (uint64_t)(o.tile_mb > 0 ? o.tile_mb : 24) << 20;
    if (ybytes > 2 * tile) cu = (int)std::min<uint64_t>(4096, (ybytes + tile - 1) / tile);
  }
  if (o.dev_cdb > 0) cdb = (uint32_t)o.dev_cdb;
  if (o.dev_tiles > 0) cu = (int)o.dev_tiles;
  if (cu == L->cu && cdb == L->cdb) return false;
  vf = 1;  // VF only pads (csr_hw.cpp:108-112); the SIMT kernel has no use for it
  *cu_dev = cu; *vf_dev = vf; *cdb_dev = cdb;
  return true;
}

// The wide image (Layout::wide): with 15-bit blocks an irregular matrix costs one scattered y update per (row, block)
// pair - on a uniform matrix one per non-zero - while a block as wide as the L2 cache can hold of x (2^23 columns = 64 MB
// fp64 / 32 MB fp32) leaves a handful of pairs per row, formed in registers, and moves the scattered access to the x
// gather.  Measured on B200 it loses to the row-tiled device layout on every configuration of
// BASELINE.json - under the stream the L2 does not keep a 16-64 MB range of x resident, and gathers that miss run at a
// third of the rate of the updates they replace - so it is built on request only (option wide = 1).
uint32_t plan_wide_cdb(const Layout *L) {
  const Options &o = options();
  if (o.wide <= 0 || L->is_wide) return 0;
  int p = 23;
  if (o.wide_range_log2 >= 2) p = (int)std::min<int64_t>(23, o.wide_range_log2);
  return 1u << p;
}

static inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

namespace {

struct BlockOf {  // col -> block, shift when cols_div_blocks is a power of two
  uint32_t cdb;
  int shift;
  explicit BlockOf(uint32_t c) : cdb(c), shift(-1) {
    if ((c & (c - 1)) == 0) { shift = 0; while ((1u << shift) != c) shift++; }
  }
  inline uint32_t operator()(uint32_t col) const { return shift >= 0 ? col >> shift : col / cdb; }
};

}  // namespace

// Validation + everything that follows from the sizes alone: scan_matrix's block count and expanded_nr_cols
// (csr_hw.cpp:25-33), thres_h - thres_l + 1 per block (:64-76,167), word geometry (util.h:61-67).
int layout_init_header(Layout *L, uint32_t rows, uint32_t cols, uint64_t nnz, int cu, int vf, int is_double,
                       uint32_t cdb_in) {
  if (cu < 1 || cu > 4096) return fail(SPMVB_E_ARG, "n_cu must be >= 1");
  if (!(vf == 1 || vf == 2 || vf == 4 || vf == 8)) return fail(SPMVB_E_ARG, "vf must be 1, 2, 4 or 8");
  if (rows == 0 || cols == 0) return fail(SPMVB_E_ARG, "empty matrix");
  uint32_t cdb = cdb_in ? cdb_in : ((cu == 10 || cu == 12) ? 16384u : 32768u);  // util.h:41-59
  const bool wide = building_wide_layout();
  if (cdb > (wide ? kWideMaxCdb : 32768u) || cdb % 4 != 0)
    return fail(SPMVB_E_ARG, "cols_div_blocks must be a multiple of 4 and <= 32768");
  if ((uint64_t)cols > (uint64_t)cdb * kMetaBlockMask) return fail(SPMVB_E_RANGE, "too many column blocks");
  L->cu = cu; L->vf = vf; L->is_double = is_double ? 1 : 0;
  L->rows = rows; L->cols = cols; L->cdb = cdb;
  L->ratio_v = is_double ? 2 : 4;
  L->ratio_col_val = kRatioCi / L->ratio_v + 1;
  L->vb = is_double ? 8 : 4;
  L->group_bytes = L->ratio_col_val * kBusBytes;
  L->chunk_bytes = L->group_bytes * kGroupsPerChunk;
  L->is_wide = wide;
  if (wide) L->chunk_bytes = (int)wide_chunk_bytes(L->vb);
  L->real_nnz = nnz;
  int blocks = (int)(cols / cdb) + 1;
  if (cols % cdb == 0) blocks--;
  L->blocks = blocks;
  uint64_t N = (uint64_t)L->ratio_v * (uint64_t)blocks;
  uint64_t ec = cols;
  if (ec % N) ec += N - ec % N;
  if (ec > 0xFFFFFFFFull) return fail(SPMVB_E_RANGE, "expanded_nr_cols overflows IndexType");
  L->expanded_cols = (uint32_t)ec;
  L->nr_cols.resize(blocks);
  for (int b = 0; b < blocks; b++)
    L->nr_cols[b] = (b == blocks - 1) ? L->expanded_cols - (uint32_t)b * cdb : cdb;
  L->xs_cfg = options().xs_config >= 0 ? (int)std::min<int64_t>(options().xs_config, kXsConfigs - 1) : 0;
  // >= 1: the chunk walk prefetches two chunks ahead and a run must cover that distance (spmv_kernels.cuh)
  if (options().run_log2 >= 0) L->run_log2 = (int)std::max<int64_t>(1, std::min<int64_t>(8, options().run_log2));
  return SPMVB_OK;
}

// hw_matrix_alloc (csr_hw.cpp:174-180) + device image offsets (each piece padded to whole chunks), from the per-piece
// nr_rows / nr_nzeros and fp[b*(cu+1)+k] = first block position of piece k (real entries only).
// Device order of the pieces.  Block-major (rows ascend through a block) is the default; when y does not fit the L2
// cache and there are several CUs, CU-major order (all blocks of CU 0's row range, then CU 1, ...) keeps the y
// range being updated L2-resident at the price of reading x once per CU - a 1 B-nnz uniform matrix goes from
// DRAM-bound scattered read-modify-writes on every update to L2 atomics.
void layout_finish_pieces(Layout *L, const uint64_t *fp, const uint32_t *pad_rows, std::vector<uint64_t> &piece_last_rank) {
  const int cu = L->cu, blocks = L->blocks;
  const size_t KB = (size_t)cu * blocks;
  L->cu_major = cu > 1 && ((uint64_t)L->rows * L->vb > ((uint64_t)48 << 20) || building_device_layout());
  if (options().cu_major >= 0) L->cu_major = cu > 1 && options().cu_major != 0;
  L->nr_ci.assign(KB, 0); L->nr_val.assign(KB, 0);
  L->piece_off.assign(KB, 0); L->piece_chunk0.assign(KB, 0); L->piece_chunk1.assign(KB, 0); L->piece_real_nnz.assign(KB, 0);
  L->dev_order.resize(KB);
  uint64_t off = 0, chunk0 = 0, padded = 0;
  for (size_t i = 0; i < KB; i++) {
    const int b = L->cu_major ? (int)(i % blocks) : (int)(i / cu);
    const int k = L->cu_major ? (int)(i / blocks) : (int)(i % cu);
    const size_t kb = (size_t)k * blocks + b, bk = (size_t)b * cu + k;
    L->dev_order[i] = (uint32_t)bk;
    uint32_t n = L->nr_nzeros[kb];
    L->nr_ci[kb] = (n + kRatioCi - 1) / kRatioCi;
    L->nr_val[kb] = n / (uint32_t)L->ratio_v;  // floors like the reference (Q1); ceil(n / ratio_v) words are stored
    uint64_t nchunks = ((uint64_t)L->nr_ci[kb] + kGroupsPerChunk - 1) / kGroupsPerChunk;
    L->piece_off[bk] = off; L->piece_chunk0[bk] = chunk0; L->piece_chunk1[bk] = chunk0 + nchunks;
    L->piece_real_nnz[bk] = (uint32_t)(fp[(size_t)b * (cu + 1) + k + 1] - fp[(size_t)b * (cu + 1) + k]);
    off += nchunks * (uint64_t)L->chunk_bytes;
    chunk0 += nchunks;
    padded += n;
  }
  L->stream_bytes = off; L->n_chunks = chunk0; L->padded_nnz = padded;
  // last row-map rank owned by each piece: ranks are contiguous per block in CU order
  piece_last_rank.assign(KB, 0);
  for (int b = 0; b < blocks; b++) {
    uint64_t rows_before = 0;
    for (int k = 0; k < cu; k++) {
      uint64_t real_rows = L->nr_rows[(size_t)k * blocks + b];
      if (k == cu - 1) real_rows -= pad_rows[b];
      if (real_rows) piece_last_rank[(size_t)b * cu + k] = L->rank_base[b] + rows_before + real_rows - 1;
      rows_before += real_rows;
    }
  }
}

namespace {

template <typename RP>
int build_impl(uint32_t rows, uint32_t cols, const RP *row_ptr, const uint32_t *col_ind, const void *values, int cu,
               int vf, int is_double, uint32_t cdb_in, Layout **out, bool plan_device = true) {
  if (!out) return fail(SPMVB_E_ARG, "out is NULL");
  *out = nullptr;
  if (rows == 0 || cols == 0 || !row_ptr) return fail(SPMVB_E_ARG, "empty matrix");
  const uint64_t nnz = (uint64_t)row_ptr[rows];
  if (nnz && (!col_ind || !values)) return fail(SPMVB_E_ARG, "col_ind/values are NULL");
  {  // row_ptr must start at 0 and never decrease (the GPU builder rejects the same input, LbRowHeads)
    int bad = row_ptr[0] != 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t r = 0; r < (int64_t)rows; r++) bad |= row_ptr[r + 1] < row_ptr[r];
    if (bad) return fail(SPMVB_E_ARG, "row_ptr must start at 0 and be non-decreasing");
  }

  Layout *L = new Layout();
  {
    int rc = layout_init_header(L, rows, cols, nnz, cu, vf, is_double, cdb_in);
    if (rc) { delete L; return rc; }
  }
  double t_phase = omp_get_wtime();
  auto phase = [&](const char *what) {  // option build_trace: time of every phase of the host builder
    if (options().build_trace <= 0) return;
    const double t = omp_get_wtime();
    fprintf(stderr, "[host layout build cu=%d cdb=%u] %-28s %8.1f ms\n", cu, L->cdb, what, (t - t_phase) * 1e3);
    t_phase = t;
  };
  const uint32_t cdb = L->cdb;
  const int blocks = L->blocks;
  const int vb = L->vb;
  const uint32_t ratio_v = (uint32_t)L->ratio_v;
  const BlockOf block_of(cdb);
  // where entry e of a piece lives: the reference's groups (index word + value words, csr_hw.cpp:270-318), or the
  // planes of a wide image's chunk (layout.h: kWide*)
  const bool wide = L->is_wide;
  const uint64_t gbytes = (uint64_t)L->group_bytes, cbytes = (uint64_t)L->chunk_bytes;
  struct Slot { uint8_t *idx, *hi, *val; };
  auto slot_of = [=](uint8_t *piece, uint64_t e) -> Slot {
    const uint32_t s = (uint32_t)(e % kRatioCi);
    if (!wide) {
      uint8_t *grp = piece + (e / kRatioCi) * gbytes;
      return Slot{grp + 2 * s, nullptr, grp + kBusBytes + (size_t)s * vb};
    }
    uint8_t *cb = piece + (e / kChunkEntries) * cbytes;
    const uint32_t lane = (uint32_t)(e % kChunkEntries) / kRatioCi;
    return Slot{cb + 16 * lane + 2 * s, cb + kWideHiOff + 8 * lane + s,
                cb + kWideValOff + (s / ratio_v) * kWidePlane + 16 * lane + (size_t)(s % ratio_v) * vb};
  };

  // row ranges of (almost) equal non-zero count, one per thread
  int T = omp_get_max_threads();
  if (T < 1) T = 1;
  if ((uint64_t)T > (uint64_t)rows) T = (int)rows;
  std::vector<uint32_t> rb(T + 1);
  rb[0] = 0; rb[T] = rows;
  for (int t = 1; t < T; t++) {
    uint64_t target = nnz / T * t;
    const RP *p = std::lower_bound(row_ptr, row_ptr + rows + 1, (RP)target);
    uint32_t r = (uint32_t)(p - row_ptr);
    if (r > rows) r = rows;
    rb[t] = std::max(r, rb[t - 1]);
  }

  // ---- pass 1: per (thread, block) counts of non-empty (row, block) pairs and padded entries (csr_hw.cpp:87-119)
  const size_t TB = (size_t)T * blocks;
  std::vector<uint64_t> pairs_t(TB, 0), zpad_t(TB, 0);
  std::vector<uint8_t> rowkind(rows, 0);  // column blocks touched by the row: 0 = none, 1 = one, 2 = several
  int bad_col = 0;
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    std::vector<uint32_t> cnt(blocks, 0), touched;
    uint64_t *pt = &pairs_t[(size_t)t * blocks], *zt = &zpad_t[(size_t)t * blocks];
    for (uint32_t r = rb[t]; r < rb[t + 1]; r++) {
      for (uint64_t j = row_ptr[r]; j < (uint64_t)row_ptr[r + 1]; j++) {
        uint32_t c = col_ind[j];
        if (c >= cols) { bad_col = 1; continue; }
        uint32_t b = block_of(c);
        if (cnt[b]++ == 0) touched.push_back(b);
      }
      for (uint32_t b : touched) { pt[b]++; zt[b] += round_up(cnt[b], (uint32_t)vf); cnt[b] = 0; }
      rowkind[r] = touched.size() > 1 ? 2 : (uint8_t)touched.size();
      touched.clear();
    }
  }
  if (bad_col) { delete L; return fail(SPMVB_E_ARG, "column index out of range"); }

  phase("pass 1 (count)");
  // exclusive scan over threads per block -> per-thread starting rank / position inside each block
  std::vector<uint64_t> P(blocks), Z(blocks);
  for (int b = 0; b < blocks; b++) {
    uint64_t pr = 0, zp = 0;
    for (int t = 0; t < T; t++) {
      uint64_t a = pairs_t[(size_t)t * blocks + b], z = zpad_t[(size_t)t * blocks + b];
      pairs_t[(size_t)t * blocks + b] = pr; zpad_t[(size_t)t * blocks + b] = zp;
      pr += a; zp += z;
    }
    P[b] = pr; Z[b] = zp;
    if (zp + (uint64_t)ratio_v * vf > 0xFFFFFFFFull) { delete L; return fail(SPMVB_E_RANGE, "block nnz overflows IndexType"); }
  }
  L->rank_base.assign(blocks + 1, 0);
  for (int b = 0; b < blocks; b++) L->rank_base[b + 1] = L->rank_base[b] + P[b];
  L->n_pairs = L->rank_base[blocks];
  if (L->n_pairs >= 0xFFFFFFF0ull) { delete L; return fail(SPMVB_E_RANGE, "too many (row, block) pairs for a 32-bit rank"); }
  L->rowmap = (uint32_t *)layout_big_alloc(std::max<uint64_t>(L->n_pairs, 1) * 4, false);
  if (!L->rowmap) { delete L; return fail(SPMVB_E_NOMEM, "rowmap"); }

  // ---- pass 2 (CU > 1 only): padded segment lengths in rank order, needed by the sequential split rule
  struct FreeOnExit { void *p = nullptr; ~FreeOnExit() { free(p); } } seglen_mem;  // every return path releases it
  uint32_t *seglen = nullptr;
  if (cu > 1) {
    seglen_mem.p = layout_big_alloc(std::max<uint64_t>(L->n_pairs, 1) * 4, false);  // every entry is written below
    seglen = (uint32_t *)seglen_mem.p;
    if (!seglen) { delete L; return fail(SPMVB_E_NOMEM, "segment lengths"); }
#pragma omp parallel num_threads(T)
    {
      const int t = omp_get_thread_num();
      std::vector<uint32_t> cnt(blocks, 0), touched;
      std::vector<uint64_t> rank(blocks);
      for (int b = 0; b < blocks; b++) rank[b] = L->rank_base[b] + pairs_t[(size_t)t * blocks + b];
      for (uint32_t r = rb[t]; r < rb[t + 1]; r++) {
        for (uint64_t j = row_ptr[r]; j < (uint64_t)row_ptr[r + 1]; j++) {
          uint32_t b = block_of(col_ind[j]);
          if (cnt[b]++ == 0) { touched.push_back(b); __builtin_prefetch(&seglen[rank[b]], 1); }
        }
        for (uint32_t b : touched) {
          seglen[rank[b]] = round_up(cnt[b], (uint32_t)vf);  // (the row map itself is written by pass 3)
          rank[b]++; cnt[b] = 0;
        }
        touched.clear();
      }
    }
  }

  phase("scan + pass 2");
  // ---- prepare_balanced_hw_matrix: S1 && S2 && S3 split per block (csr_hw.cpp:459-468), leftovers + row padding to
  //      the last CU (:474-482).  fp[b*(cu+1)+k] = first block position of piece k (real entries only).
  const size_t KB = (size_t)cu * blocks;
  L->nr_rows.assign(KB, 0); L->nr_nzeros.assign(KB, 0); L->nr_ci.assign(KB, 0); L->nr_val.assign(KB, 0);
  std::vector<uint64_t> fp((size_t)blocks * (cu + 1), 0);
  std::vector<uint32_t> pad_rows(blocks, 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < blocks; b++) {
    uint64_t *f = &fp[(size_t)b * (cu + 1)];
    uint64_t nz = 0, rc = 0, pos = 0;
    int fired = 0;
    if (cu > 1) {
      const uint64_t thr = Z[b] / (uint64_t)cu;
      const uint64_t base = L->rank_base[b];
      for (uint64_t i = 0; i < P[b]; i++) {
        uint32_t len = seglen[base + i];
        nz += len; rc++; pos += len;
        if (nz > thr && nz % ratio_v == 0 && rc % ratio_v == 0 && fired < cu - 1) {
          L->nr_rows[(size_t)fired * blocks + b] = (uint32_t)rc;
          L->nr_nzeros[(size_t)fired * blocks + b] = (uint32_t)nz;
          fired++;
          f[fired] = pos;
          nz = 0; rc = 0;
        }
      }
    } else {
      nz = Z[b]; rc = P[b];
    }
    for (int k = fired + 1; k < cu; k++) f[k] = f[fired];  // un-fired middle CUs stay empty (reference: garbage, Q2)
    f[cu] = Z[b];
    uint32_t mod = (uint32_t)(rc % ratio_v);
    if (mod) { pad_rows[b] = ratio_v - mod; rc += pad_rows[b]; nz += (uint64_t)pad_rows[b] * vf; }
    L->nr_rows[(size_t)(cu - 1) * blocks + b] = (uint32_t)rc;
    L->nr_nzeros[(size_t)(cu - 1) * blocks + b] = (uint32_t)nz;
  }
  free(seglen_mem.p); seglen_mem.p = nullptr; seglen = nullptr;

  phase("split");
  // ---- hw_matrix_alloc + device image offsets
  std::vector<uint64_t> piece_last_rank;
  layout_finish_pieces(L, fp.data(), pad_rows.data(), piece_last_rank);
  const uint64_t off = L->stream_bytes, chunk0 = L->n_chunks;
  L->stream = (uint8_t *)layout_big_alloc(off, true);
  L->chunks = (ChunkMeta *)calloc((size_t)std::max<uint64_t>(chunk0, 1), sizeof(ChunkMeta));
  if (!L->stream || !L->chunks) { delete L; return fail(SPMVB_E_NOMEM, "stream"); }
#pragma omp parallel for schedule(static)
  for (int64_t bk = 0; bk < (int64_t)KB; bk++) {
    const int b = (int)(bk / cu);
    uint64_t c0 = L->piece_chunk0[bk];
    uint64_t c1 = L->piece_chunk1[bk];
    uint32_t real = L->piece_real_nnz[bk];
    for (uint64_t c = c0; c < c1; c++) {
      uint64_t first = (c - c0) * kChunkEntries;
      L->chunks[c].block = (uint32_t)b;
      L->chunks[c].valid = first >= real ? 0u : (uint32_t)std::min<uint64_t>(kChunkEntries, real - first);
    }
  }

  phase("alloc + chunk init");
  // ---- pass 3: scatter every entry to its final byte (create_block_matrix + generate_balanced_hw_submatrix)
  const uint8_t *vals = (const uint8_t *)values;
  const int gb = L->group_bytes;
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    // everything the scatter needs to know about a column block in ONE cache line: with thousands of blocks the
    // per-block tables (counts, cursors, piece bounds, piece addresses) otherwise cost five or six misses per entry
    struct alignas(64) Cursor {
      uint32_t cnt, fill;     // entries of the current row in this block: counted / written
      int32_t k;              // CU piece that holds pos (monotone)
      uint32_t unused;
      uint64_t pos, rank;     // block position of the row's first entry; rank of the (row, block) pair
      uint64_t piece_start;   // fp[b][k]
      uint8_t *piece;         // stream + piece_off[b][k]
      uint64_t chunk0;        // piece_chunk0[b][k]
    };
    std::vector<Cursor> cur((size_t)blocks);
    std::vector<uint32_t> touched;
    auto enter_piece = [&](Cursor &c, uint32_t b, int k) {
      c.k = k;
      c.piece_start = fp[(size_t)b * (cu + 1) + k];
      c.piece = L->stream + L->piece_off[(size_t)b * cu + k];
      c.chunk0 = L->piece_chunk0[(size_t)b * cu + k];
    };
    for (int b = 0; b < blocks; b++) {
      Cursor &c = cur[b];
      c.cnt = 0; c.fill = 0; c.unused = 0;
      c.rank = L->rank_base[b] + pairs_t[(size_t)t * blocks + b];
      c.pos = zpad_t[(size_t)t * blocks + b];
      enter_piece(c, (uint32_t)b, 0);
    }
    for (uint32_t r = rb[t]; r < rb[t + 1]; r++) {
      const uint64_t j0 = row_ptr[r], j1 = row_ptr[r + 1];
      for (uint64_t j = j0; j < j1; j++) {
        uint32_t b = block_of(col_ind[j]);
        if (cur[b].cnt++ == 0) touched.push_back(b);
      }
      // a segment never straddles pieces: resolve the piece once per (row, block)
      for (uint32_t b : touched) {
        Cursor &c = cur[b];
        const uint64_t *f = &fp[(size_t)b * (cu + 1)];
        int k = c.k;
        while (k < cu - 1 && c.pos >= f[k + 1]) k++;
        if (k != c.k) enter_piece(c, b, k);
        // the row's first slot in this block and its row-map entry are written a few hundred instructions from now
        const Slot first = slot_of(c.piece, c.pos - c.piece_start);
        __builtin_prefetch(first.idx, 1);
        __builtin_prefetch(first.val, 1);
        __builtin_prefetch(&L->rowmap[c.rank], 1);
      }
      for (uint64_t j = j0; j < j1; j++) {
        const uint32_t col = col_ind[j];
        const uint32_t b = block_of(col);
        Cursor &c = cur[b];
        const uint64_t e = c.pos + c.fill - c.piece_start;
        const uint32_t is_last = (++c.fill == c.cnt) && (c.cnt % (uint32_t)vf == 0);
        const Slot sl = slot_of(c.piece, e);
        const uint32_t cin = col - b * cdb;
        const uint16_t ci = (uint16_t)((cin & 0x7FFFu) | (is_last ? 0x8000u : 0u));  // csr_hw.cpp:220, :288-292
        memcpy(sl.idx, &ci, 2);
        if (sl.hi) *sl.hi = (uint8_t)(cin >> 15);
        memcpy(sl.val, vals + (size_t)j * vb, vb);                                 // csr_hw.cpp:300-310
      }
      for (uint32_t b : touched) {
        Cursor &c = cur[b];
        const uint64_t s_e = c.pos - c.piece_start;
        const uint64_t t_e = s_e + round_up(c.cnt, (uint32_t)vf);  // VF padding: (col 0, val 0), csr_hw.cpp:229-238
        if (c.cnt % (uint32_t)vf != 0) {
          const uint64_t e = t_e - 1;
          const uint16_t ci = 0x8000u;
          memcpy(slot_of(c.piece, e).idx, &ci, 2);
        }
        L->rowmap[c.rank] = r;
        // chunks whose first entry lies inside this segment start at this rank
        for (uint64_t ch = (s_e + kChunkEntries - 1) / kChunkEntries; ch * kChunkEntries < t_e; ch++) {
          ChunkMeta &cm = L->chunks[c.chunk0 + ch];
          cm.rank0 = (uint32_t)c.rank;
          if (ch * kChunkEntries > s_e) cm.valid |= kChunkStartsMid;
        }
        c.pos = c.piece_start + t_e;
        c.rank++;
        c.cnt = 0; c.fill = 0;
      }
      touched.clear();
    }
  }

  phase("pass 3 (scatter)");
  // padding rows of the last CU: VF x (col 0, val 0) with the end-of-row bit on the last (csr_hw.cpp:246-255)
  for (int b = 0; b < blocks; b++) {
    const size_t bk = (size_t)b * cu + (cu - 1);
    const uint64_t real = L->piece_real_nnz[bk];
    for (uint32_t i = 0; i < pad_rows[b]; i++) {
      const uint64_t e = real + (uint64_t)i * vf + (vf - 1);
      const uint16_t ci = 0x8000u;
      memcpy(slot_of(L->stream + L->piece_off[bk], e).idx, &ci, 2);
    }
  }

  // consecutive-rows fast path: rows of a block ascend, so first/last rank spanning equal row distance <=> consecutive
#pragma omp parallel for schedule(static)
  for (int64_t bk = 0; bk < (int64_t)KB; bk++) {
    uint64_t c0 = L->piece_chunk0[bk];
    uint64_t c1 = L->piece_chunk1[bk];
    const uint64_t last_piece_rank = piece_last_rank[bk];
    for (uint64_t c = c0; c < c1; c++) {
      ChunkMeta &m = L->chunks[c];
      if (!(m.valid & 0x3FFu)) continue;
      m.row_first = L->rowmap[m.rank0];
      uint64_t last_rank = (c + 1 < c1 && (L->chunks[c + 1].valid & 0x3FFu)) ? L->chunks[c + 1].rank0 : last_piece_rank;
      if ((uint64_t)L->rowmap[last_rank] - m.row_first == last_rank - m.rank0) m.valid |= kChunkRowsConsecutive;
      bool sole = true;
      for (uint64_t rk = m.rank0; rk <= last_rank && sole; rk++) sole = rowkind[L->rowmap[rk]] == 1;
      if (sole) m.valid |= kChunkSole;
    }
  }

  phase("consecutive / sole flags");
  // Rows to clear before every SpMV (see Layout::zero_rows).  Marks are idempotent byte stores.
  {
    const uint64_t R = 1ull << L->run_log2;
    std::vector<uint8_t> needz(rows, 0);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)rows; r++) needz[r] = rowkind[r] == 0;
#pragma omp parallel for schedule(static)
    for (int64_t bk = 0; bk < (int64_t)KB; bk++) {
      uint64_t c0 = L->piece_chunk0[bk];
      uint64_t c1 = L->piece_chunk1[bk];
      for (uint64_t c = c0; c < c1; c++) {
        const ChunkMeta &m = L->chunks[c];
        if (!(m.valid & 0x3FFu)) continue;
        // row split across two runs; runs are aligned globally (OCC kernel) or to the piece start (XS kernel)
        if ((m.valid & kChunkStartsMid) && ((c % R) == 0 || ((c - c0) % R) == 0))
          needz[m.row_first] = 1;
        if (m.valid & kChunkSole) continue;
        // every row with a segment (or part of one) in an atomics-only chunk
        uint64_t rk = m.rank0;
        uint64_t end_rank;
        if (c + 1 < c1 && (L->chunks[c + 1].valid & 0x3FFu))
          end_rank = L->chunks[c + 1].rank0;  // inclusive: also covers a row that spills into the next chunk
        else
          end_rank = piece_last_rank[bk];
        for (; rk <= end_rank; rk++) needz[L->rowmap[rk]] = 1;
      }
    }
    uint64_t nz = 0;
    for (uint32_t r = 0; r < rows; r++) nz += needz[r];
    L->zero_all = nz > (uint64_t)rows / 3;
    if (options().zero_all > 0) L->zero_all = true;
    if (!L->zero_all) {
      L->zero_rows.reserve(nz);
      for (uint32_t r = 0; r < rows; r++)
        if (needz[r]) L->zero_rows.push_back(r);
    }
  }

  phase("rows to clear");
  // column range of every chunk (real entries only)
  L->chunk_col_lo.assign((size_t)L->n_chunks, 0xFFFF);
  L->chunk_col_hi.assign((size_t)L->n_chunks, 0);
  L->chunk_x_lines.assign((size_t)L->n_chunks, 0);
  const int line_shift = vb == 8 ? 4 : 5;  // columns per 128-byte line of x: 16 (fp64) / 32 (fp32)
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < (int64_t)L->n_chunks; c++) {
    const uint32_t valid = L->chunks[c].valid & 0x3FFu;
    uint8_t *base = L->stream + (uint64_t)c * L->chunk_bytes;
    uint16_t lo = 0xFFFF, hi = 0;
    uint32_t row_ends = 0, lines = 0;
    uint32_t seen[64] = {0};  // 32768 columns >> 4 = 2048 lines at most
    for (uint32_t e = 0; e < valid; e++) {
      uint16_t ci;
      memcpy(&ci, slot_of(base, e).idx, 2);
      row_ends += ci >> 15;
      ci &= 0x7FFF;
      if (wide) continue;  // no shared-memory windows over a wide image: its column statistics are not used
      lo = std::min(lo, ci); hi = std::max(hi, ci);
      const uint32_t ln = (uint32_t)ci >> line_shift, bit = 1u << (ln & 31);
      lines += !(seen[ln >> 5] & bit);
      seen[ln >> 5] |= bit;
    }
    if (wide) { lo = 0; hi = 0; }
    L->chunk_col_lo[c] = lo; L->chunk_col_hi[c] = hi; L->chunk_x_lines[c] = (uint16_t)lines;
    L->chunks[c].block = (L->chunks[c].block & kMetaBlockMask) | (row_ends << kMetaRowsShift);
  }

  phase("chunk column ranges");
  // the engine-private device layout, when the API parameters are not what the GPU should stream
  int cu_dev = cu, vf_dev = vf;
  uint32_t cdb_dev = cdb;
  if (plan_device && plan_device_params(L, &cu_dev, &vf_dev, &cdb_dev)) {
    building_device_layout() = true;
    int rc = build_impl<RP>(rows, cols, row_ptr, col_ind, values, cu_dev, vf_dev, is_double, cdb_dev, &L->dev, false);
    building_device_layout() = false;
    if (rc) { delete L; return rc; }
  }
  // the wide image: the same builder with one column block per L2-sized range of x, one compute unit, no VF padding
  if (const uint32_t cdb_wide = plan_device ? plan_wide_cdb(L) : 0u) {
    phase("device layout");
    building_wide_layout() = true;
    int rc = build_impl<RP>(rows, cols, row_ptr, col_ind, values, 1, 1, is_double, cdb_wide, &L->wide, false);
    building_wide_layout() = false;
    if (rc) { delete L; return rc; }
    phase("wide image");
  }
  // the sliced-ELLPACK image, when the matrix is regular enough for it to cost nothing (ell.h)
  if (plan_device) {
    L->ell = build_ell<RP>(rows, cols, row_ptr, col_ind, values, is_double);
    phase("ELL image");
  }

  *out = L;
  return SPMVB_OK;
}

}  // namespace
}  // namespace spmvb

using namespace spmvb;

extern "C" {

int spmvb_version(void) { return 200; }

int spmvb_set_option(const char *name, int64_t value) {
  if (!name) return fail(SPMVB_E_ARG, "set_option: NULL");
  for (const OptionName &n : kOptionNames)
    if (strcmp(n.name, name) == 0) { options().*(n.field) = value; return SPMVB_OK; }
  return fail(SPMVB_E_ARG, std::string("set_option: unknown option ") + name);
}
int64_t spmvb_get_option(const char *name) {
  if (name)
    for (const OptionName &n : kOptionNames)
      if (strcmp(n.name, name) == 0) return options().*(n.field);
  return INT64_MIN;
}
// SPMVB_<NAME>=<integer> for every option; meant for executables that have no other way to be configured (the
// reference's run.elf takes compile-time macros only).  Returns the number of variables applied.
int spmvb_options_from_env(void) {
  int n = 0;
  for (const OptionName &nm : kOptionNames) {
    std::string var = "SPMVB_";
    for (const char *p = nm.name; *p; p++) var += (char)toupper((unsigned char)*p);
    if (const char *v = getenv(var.c_str())) { options().*(nm.field) = atoll(v); n++; }
  }
  return n;
}

int spmvb_layout_build(uint32_t rows, uint32_t cols, const uint64_t *row_ptr, const uint32_t *col_ind,
                       const void *values, int n_cu, int vf, int is_double, uint32_t cols_div_blocks,
                       spmvb_layout **out) {
  return build_impl<uint64_t>(rows, cols, row_ptr, col_ind, values, n_cu, vf, is_double, cols_div_blocks,
                              (Layout **)out);
}

int spmvb_layout_build_u32(uint32_t rows, uint32_t cols, const uint32_t *row_ptr, const uint32_t *col_ind,
                           const void *values, int n_cu, int vf, int is_double, uint32_t cols_div_blocks,
                           spmvb_layout **out) {
  return build_impl<uint32_t>(rows, cols, row_ptr, col_ind, values, n_cu, vf, is_double, cols_div_blocks,
                              (Layout **)out);
}

void spmvb_layout_free(spmvb_layout *l) { delete (Layout *)l; }

int spmvb_layout_blocks(const spmvb_layout *l) { return ((const Layout *)l)->blocks; }
int spmvb_layout_n_cu(const spmvb_layout *l) { return ((const Layout *)l)->cu; }
uint32_t spmvb_layout_rows(const spmvb_layout *l) { return ((const Layout *)l)->rows; }
uint32_t spmvb_layout_cols(const spmvb_layout *l) { return ((const Layout *)l)->cols; }
uint32_t spmvb_layout_expanded_cols(const spmvb_layout *l) { return ((const Layout *)l)->expanded_cols; }
uint64_t spmvb_layout_real_nnz(const spmvb_layout *l) { return ((const Layout *)l)->real_nnz; }
uint64_t spmvb_layout_padded_nnz(const spmvb_layout *l) { return ((const Layout *)l)->padded_nnz; }
uint64_t spmvb_layout_pairs(const spmvb_layout *l) { return ((const Layout *)l)->n_pairs; }
uint64_t spmvb_layout_stream_bytes(const spmvb_layout *l) { return ((const Layout *)l)->stream_bytes; }
// chunks are units of the device image: these two describe the engine-private device layout when there is one
uint64_t spmvb_layout_chunks(const spmvb_layout *l) {
  const Layout *L = (const Layout *)l;
  return (L->dev ? L->dev : L)->n_chunks;
}
int spmvb_layout_chunk_cols(const spmvb_layout *l, uint64_t c, uint32_t *lo, uint32_t *hi, uint32_t *block) {
  const Layout *L = (const Layout *)l;
  if (L && L->dev) L = L->dev;
  if (!L || c >= L->n_chunks) return fail(SPMVB_E_ARG, "chunk index");
  if (lo) *lo = L->chunk_col_lo[c];
  if (hi) *hi = L->chunk_col_hi[c];
  if (block) *block = L->chunks[c].block & kMetaBlockMask;
  return SPMVB_OK;
}
int64_t spmvb_layout_zero_rows(const spmvb_layout *l) {
  const Layout *L = (const Layout *)l;
  return L->zero_all ? -1 : (int64_t)L->zero_rows.size();
}

int spmvb_layout_piece_info(const spmvb_layout *l, int cu, int block, uint32_t *out) {
  const Layout *L = (const Layout *)l;
  if (!L || !out || cu < 0 || cu >= L->cu || block < 0 || block >= L->blocks) return fail(SPMVB_E_ARG, "piece index");
  const size_t kb = (size_t)cu * L->blocks + block;
  out[0] = L->nr_rows[kb]; out[1] = L->nr_cols[block]; out[2] = L->nr_nzeros[kb];
  out[3] = L->nr_ci[kb];   out[4] = L->nr_val[kb];
  return SPMVB_OK;
}

const void *spmvb_layout_piece_words(const spmvb_layout *l, int cu, int block) {
  const Layout *L = (const Layout *)l;
  if (!L || cu < 0 || cu >= L->cu || block < 0 || block >= L->blocks) return nullptr;
  if (!L->stream) { set_error("piece_words: GPU-built layout, call spmvb_engine_fetch_layout first"); return nullptr; }
  return L->stream + L->piece_off[(size_t)block * L->cu + cu];
}

int spmvb_layout_bitmap_row(const spmvb_layout *l, int block, uint8_t *out) {
  const Layout *L = (const Layout *)l;
  if (!L || !out || block < 0 || block >= L->blocks) return fail(SPMVB_E_ARG, "block index");
  if (!L->rowmap) return fail(SPMVB_E_ARG, "bitmap_row: GPU-built layout, call spmvb_engine_fetch_layout first");
  memset(out, 1, L->rows);  // 1 = row has no non-zero in this block (csr_hw.cpp:340-345)
  for (uint64_t i = L->rank_base[block]; i < L->rank_base[block + 1]; i++) out[L->rowmap[i]] = 0;
  return SPMVB_OK;
}

double spmvb_layout_storage_mb(const spmvb_layout *l, int cu) {
  const Layout *L = (const Layout *)l;
  if (!L || cu < 0 || cu >= L->cu) return -1.0;
  // storage_overhead, csr_hw.cpp:1401-1409 (its IndexType accumulator wraps at 2^32 bits; this one does not)
  double bits = (double)L->blocks * 5 * 32;
  for (int b = 0; b < L->blocks; b++) {
    const size_t kb = (size_t)cu * L->blocks + b;
    bits += ((double)L->nr_ci[kb] + (double)L->nr_val[kb]) * 128.0;
  }
  return bits / (8.0 * 1024 * 1024);
}

// Column ranges [first, end) that a SpMV with this layout can read from x: maximal runs of column blocks that hold at
// least one entry, in units of whole blocks (the last block ends at blocks * cols_div_blocks).  A row shard of a banded
// matrix touches only its band; a whole matrix normally gives the single range [0, blocks * cols_div_blocks).
int64_t spmvb_layout_x_ranges(const spmvb_layout *l, uint64_t *out_pairs, uint64_t max_ranges) {
  const Layout *L = (const Layout *)l;
  if (!L) return fail(SPMVB_E_ARG, "x_ranges");
  uint64_t n = 0, open_first = 0;
  bool open = false;
  for (int b = 0; b <= L->blocks; b++) {
    bool touched = false;
    if (b < L->blocks)
      for (int k = 0; k < L->cu && !touched; k++) touched = L->piece_real_nnz[(size_t)b * L->cu + k] != 0;
    if (touched && !open) { open = true; open_first = (uint64_t)b * L->cdb; }
    if (!touched && open) {
      if (out_pairs && n < max_ranges) { out_pairs[2 * n] = open_first; out_pairs[2 * n + 1] = (uint64_t)b * L->cdb; }
      n++;
      open = false;
    }
  }
  return (int64_t)n;
}

int spmvb_layout_pack_x(const spmvb_layout *l, const void *x, uint32_t n, void *out) {
  const Layout *L = (const Layout *)l;
  if (!L || !x || !out) return fail(SPMVB_E_ARG, "pack_x");
  const uint32_t m = std::min(n, L->expanded_cols);
  memcpy(out, x, (size_t)m * L->vb);
  memset((uint8_t *)out + (size_t)m * L->vb, 0, (size_t)(L->expanded_cols - m) * L->vb);
  return SPMVB_OK;
}

static int layout_equal_impl(const Layout *A, const Layout *B, char *why, size_t why_len, bool need_image) {
  auto say = [&](const char *what) { if (why && why_len) snprintf(why, why_len, "%s", what); return 0; };
  if (!A || !B) return fail(SPMVB_E_ARG, "layout_equal: NULL");
  const bool images = A->stream && B->stream && A->rowmap && B->rowmap;
  if (need_image && !images)
    return fail(SPMVB_E_ARG, "layout_equal: host image not present (spmvb_engine_fetch_layout)");
  if (A->cu != B->cu || A->vf != B->vf || A->is_double != B->is_double || A->blocks != B->blocks || A->rows != B->rows ||
      A->cols != B->cols || A->expanded_cols != B->expanded_cols || A->cdb != B->cdb || A->real_nnz != B->real_nnz ||
      A->padded_nnz != B->padded_nnz || A->n_pairs != B->n_pairs || A->stream_bytes != B->stream_bytes ||
      A->n_chunks != B->n_chunks || A->run_log2 != B->run_log2 || A->cu_major != B->cu_major || A->xs_cfg != B->xs_cfg)
    return say("header");
  if (A->nr_rows != B->nr_rows) return say("nr_rows");
  if (A->nr_nzeros != B->nr_nzeros) return say("nr_nzeros");
  if (A->nr_ci != B->nr_ci || A->nr_val != B->nr_val || A->nr_cols != B->nr_cols) return say("nr_ci/nr_val/nr_cols");
  if (A->piece_off != B->piece_off || A->piece_chunk0 != B->piece_chunk0 || A->piece_chunk1 != B->piece_chunk1 ||
      A->dev_order != B->dev_order || A->piece_real_nnz != B->piece_real_nnz)
    return say("piece tables");
  if (A->rank_base != B->rank_base) return say("rank_base");
  if (images && memcmp(A->rowmap, B->rowmap, (size_t)A->n_pairs * 4) != 0) return say("rowmap");
  if (images && memcmp(A->stream, B->stream, (size_t)A->stream_bytes) != 0) return say("stream");
  for (uint64_t c = 0; c < A->n_chunks; c++)
    if (memcmp(&A->chunks[c], &B->chunks[c], sizeof(ChunkMeta)) != 0) {
      if (why && why_len)
        snprintf(why, why_len, "chunk meta %llu: {%u,%#x,%#x,%u} vs {%u,%#x,%#x,%u}", (unsigned long long)c,
                 A->chunks[c].rank0, A->chunks[c].block, A->chunks[c].valid, A->chunks[c].row_first,
                 B->chunks[c].rank0, B->chunks[c].block, B->chunks[c].valid, B->chunks[c].row_first);
      return 0;
    }
  if (A->zero_all != B->zero_all) return say("zero_all");
  if (A->zero_rows != B->zero_rows) return say("zero_rows");
  if (A->chunk_col_lo != B->chunk_col_lo || A->chunk_col_hi != B->chunk_col_hi) return say("chunk column ranges");
  if (A->chunk_x_lines != B->chunk_x_lines) return say("chunk x lines");
  if ((A->dev != nullptr) != (B->dev != nullptr)) return say("device layout present in one only");
  if (A->dev) {  // the engine-private device layout: its image stays on the GPU when it was built there
    char sub[200] = "";
    const int r = layout_equal_impl(A->dev, B->dev, sub, sizeof sub, false);
    if (r != 1) {
      if (why && why_len) snprintf(why, why_len, "device layout: %s", sub);
      return r;
    }
  }
  if (why && why_len) why[0] = 0;
  return 1;
}

int spmvb_layout_equal(const spmvb_layout *a, const spmvb_layout *b, char *why, size_t why_len) {
  return layout_equal_impl((const Layout *)a, (const Layout *)b, why, why_len, true);
}

/* mean number of distinct 128-byte lines of x per (non-empty) chunk of the API layout: the irregularity measure */
double spmvb_layout_x_lines_per_chunk(const spmvb_layout *l) {
  const Layout *L = (const Layout *)l;
  if (!L) return -1.0;
  uint64_t lines = 0, used = 0;
  for (uint64_t c = 0; c < L->n_chunks; c++)
    if (L->chunk_x_lines[c]) { lines += L->chunk_x_lines[c]; used++; }
  return used ? (double)lines / (double)used : 0.0;
}

int spmvb_layout_device_params(const spmvb_layout *l, uint64_t *out) {
  const Layout *L = (const Layout *)l;
  if (!L || !out) return fail(SPMVB_E_ARG, "device_params");
  const Layout *D = L->dev ? L->dev : L;
  out[0] = (uint64_t)D->cu; out[1] = (uint64_t)D->vf; out[2] = D->cdb; out[3] = D->cu_major ? 1u : 0u;
  out[4] = L->dev ? 1u : 0u; out[5] = D->n_pairs; out[6] = D->n_chunks;
  out[7] = D->zero_all ? UINT64_MAX : (uint64_t)D->zero_rows.size(); out[8] = D->stream_bytes;
  return SPMVB_OK;
}

/* The wide image of this layout (engine-private, like the device layout): out[8] = {present, column-block width,
 * column blocks, (row, block) pairs, chunks, rows cleared per SpMV (UINT64_MAX = all), image bytes, real entries}. */
int spmvb_layout_wide_params(const spmvb_layout *l, uint64_t *out) {
  const Layout *L = (const Layout *)l;
  if (!L || !out) return fail(SPMVB_E_ARG, "wide_params");
  const Layout *W = L->wide;
  for (int i = 0; i < 8; i++) out[i] = 0;
  if (!W) return SPMVB_OK;
  out[0] = 1; out[1] = W->cdb; out[2] = (uint64_t)W->blocks; out[3] = W->n_pairs; out[4] = W->n_chunks;
  out[5] = W->zero_all ? UINT64_MAX : (uint64_t)W->zero_rows.size(); out[6] = W->stream_bytes; out[7] = W->real_nnz;
  return SPMVB_OK;
}

/* Walks the wide image the way the kernel does - planes of every chunk, end-of-row bits, row map - and returns its
 * real entries in image order as (row, column, value bits).  For tests: the image must hold exactly the CSR's entries,
 * ordered by (column block, row, CSR position). */
int64_t spmvb_layout_wide_decode(const spmvb_layout *l, uint32_t *rows_out, uint32_t *cols_out, void *vals_out,
                                 uint64_t max_entries) {
  const Layout *L = (const Layout *)l;
  if (!L || !L->wide) return fail(SPMVB_E_ARG, "wide_decode: no wide image");
  const Layout *W = L->wide;
  if (!W->stream || !W->rowmap) return fail(SPMVB_E_ARG, "wide_decode: image not on the host");
  uint64_t n = 0;
  const uint32_t ratio_v = (uint32_t)W->ratio_v;
  for (int b = 0; b < W->blocks; b++) {
    const uint8_t *piece = W->stream + W->piece_off[b];
    uint64_t rank = W->rank_base[b];
    const uint32_t real = W->piece_real_nnz[b];
    for (uint64_t c = W->piece_chunk0[b]; c < W->piece_chunk1[b]; c++) {
      const ChunkMeta &m = W->chunks[c];
      const uint64_t first = (c - W->piece_chunk0[b]) * kChunkEntries;
      const uint32_t valid = m.valid & 0x3FFu;
      if (valid != (first >= real ? 0u : (uint32_t)std::min<uint64_t>(kChunkEntries, real - first)))
        return fail(SPMVB_E_ARG, "wide_decode: chunk entry count");
      if (valid && m.rank0 != rank) return fail(SPMVB_E_ARG, "wide_decode: chunk rank0");
      if (valid && (m.block & kMetaBlockMask) != (uint32_t)b) return fail(SPMVB_E_ARG, "wide_decode: chunk block");
      if (valid && m.row_first != W->rowmap[rank]) return fail(SPMVB_E_ARG, "wide_decode: chunk row_first");
      const uint8_t *cb = piece + (c - W->piece_chunk0[b]) * (uint64_t)W->chunk_bytes;
      uint32_t row_ends = 0;
      for (uint32_t e = 0; e < valid; e++) {
        const uint32_t lane = e / kRatioCi, s = e % kRatioCi;
        uint16_t ci;
        memcpy(&ci, cb + 16 * lane + 2 * s, 2);
        const uint32_t col = (uint32_t)b * W->cdb + (((uint32_t)cb[kWideHiOff + 8 * lane + s] << 15) | (ci & 0x7FFFu));
        if (rank >= W->rank_base[b + 1]) return fail(SPMVB_E_ARG, "wide_decode: rank runs past the block");
        if (n < max_entries) {
          if (rows_out) rows_out[n] = W->rowmap[rank];
          if (cols_out) cols_out[n] = col;
          if (vals_out)
            memcpy((uint8_t *)vals_out + n * W->vb,
                   cb + kWideValOff + (s / ratio_v) * kWidePlane + 16 * lane + (size_t)(s % ratio_v) * W->vb, W->vb);
        }
        n++;
        if (ci & 0x8000u) { rank++; row_ends++; }
      }
      if (valid && (m.block >> kMetaRowsShift) != row_ends) return fail(SPMVB_E_ARG, "wide_decode: chunk row ends");
    }
    if (rank != W->rank_base[b + 1]) return fail(SPMVB_E_ARG, "wide_decode: pairs of the block");
  }
  return (int64_t)n;
}

int spmvb_partition_rows(uint32_t rows, const uint64_t *row_ptr, int parts, int ratio_v, uint32_t *bounds) {
  if (!row_ptr || !bounds || parts < 1 || ratio_v < 1) return fail(SPMVB_E_ARG, "partition_rows");
  const uint64_t total = row_ptr[rows];
  const uint64_t thr = total / (uint64_t)parts;
  bounds[0] = 0;
  int fired = 0;
  uint64_t nz = 0, rc = 0;
  for (uint32_t r = 0; r < rows && fired < parts - 1; r++) {
    nz += row_ptr[r + 1] - row_ptr[r];
    rc++;
    if (nz > thr && nz % (uint64_t)ratio_v == 0 && rc % (uint64_t)ratio_v == 0) {  // S1 && S2 && S3
      bounds[++fired] = r + 1;
      nz = 0; rc = 0;
    }
  }
  for (int k = fired + 1; k <= parts; k++) bounds[k] = rows;
  return SPMVB_OK;
}

}  // extern "C"
