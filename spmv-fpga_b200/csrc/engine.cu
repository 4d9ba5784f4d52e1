// Device side of the C ABI (include/spmvb.h): upload of the hw_matrix image, kernel launches, host<->device
// vector traffic.  Replaces the body of spmv_hw (reference src/csr_hw_wrapper.cpp:193-288): the per-block
// spmv() round trips and the host accum_results loop become ONE kernel launch over all (CU, block) pieces.
#include <cuda_runtime.h>
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/spmvb.h"
#include "ell.h"
#include "ell_gpu.cuh"
#include "layout.h"
#include "layout_gpu.cuh"
#include "spmv_kernels.cuh"

namespace spmvb {

enum Variant { kVariantDefault = 0, kVariantDirect = 1, kVariantOcc4 = 6, kVariantOcc3 = 7, kVariantXs = 8, kVariantWide = 9, kVariantEll = 10 };


struct Engine {
  int device = 0, is_double = 1, vb = 8, variant = kVariantDefault;
  uint32_t rows = 0, cols = 0, expanded_cols = 0, cdb = 32768;
  int blocks = 0;
  uint64_t real_nnz = 0, n_chunks = 0, n_pairs = 0, stream_bytes = 0, x_len = 0;
  uint64_t x_touched = 0;  // columns of the column blocks that hold at least one entry (what a SpMV must read of x)
  std::vector<uint64_t> x_ranges;  // [first, end) pairs of those blocks, merged: what set_x uploads
  uint8_t *d_stream = nullptr;
  uint32_t *d_rowmap = nullptr;
  uint32_t *d_zero_rows = nullptr;
  XsItem *d_items = nullptr;  // work items of the XS kernel
  uint32_t *d_cta_first = nullptr;  // [sms + 1] first item of every CTA
  // CU-major layouts: the XS work dealt tile by tile as well (spmv_host launches one kernel per row tile and sends the
  // tile's rows of y to the host while the next tile is computed)
  int n_tiles = 0;
  XsItem *d_items_t = nullptr;
  uint32_t *d_cta_first_t = nullptr;   // [n_tiles * (xs_ctas + 1)]
  std::vector<uint32_t> tile_rows_end; // rows [0, tile_rows_end[k]) are final once tiles 0..k are done
  std::vector<cudaEvent_t> ev_tile;    // [2 * n_tiles] kernel done / copy done
  cudaStream_t copy_stream = nullptr;
  // API image of a GPU-built layout whose device layout differs from it (kept for spmvb_engine_fetch_layout)
  uint8_t *d_api_stream = nullptr;
  uint32_t *d_api_rowmap = nullptr;
  int dev_cu = 1, dev_vf = 1;  // parameters of the layout the device streams (may differ from the API layout's)
  uint32_t occ_run_log2 = 3, xs_run_log2 = 1;  // run lengths of the kernels (>= the layout's zero-list granularity)
  uint32_t n_items = 0;
  double xs_windowed_frac = 0.0;  // share of the chunks whose x window fits shared memory
  bool tall = false;               // x and y both exceed the L2 cache: explicit L2 eviction policies
  bool cu_major = false;           // pieces in CU-major device order (row tiles)
  bool irregular = false;          // layout_is_irregular(): x gathers are scattered, the x-window kernel pays
  bool wide = false;               // the image is a wide image (Layout::is_wide): only the WIDE kernel can walk it
  // sliced-ELLPACK image (ell.h): regular matrices, one row per lane.  ell = the engine holds one (see ell_on()); a
  // host-built engine then holds nothing of the hw_matrix stream, a GPU-built one keeps both
  bool ell = false;
  uint8_t *d_ell = nullptr;
  uint32_t ell_slices = 0, ell_width = 0, ell_slice_bytes = 0;
  std::vector<uint32_t> ell_tile_slice;   // [tiles + 1] slice ranges of the end-to-end pipeline
  std::vector<uint64_t> ell_tile_x_end;   // [tiles] x[0, end) must be on the device before the tile runs
  uint64_t ell_x_begin = 0;               // first column any slice reads
  cudaStream_t down_stream = nullptr;     // y tiles travel to the host while x tiles still arrive
  std::vector<uint64_t> block_chunk0;  // wide image: [blocks + 1] first chunk of every column block
  std::vector<cudaEvent_t> ev_block;   // wide image: x range of block b is on the device
  int xs_cfg = 0, xs_ctas = 148;   // configuration of the x-window kernel (xs_config) and its grid = SMs x CTAs per SM
  int auto_variant = kVariantOcc3; // what variant 0 resolves to (chosen from the layout at creation)
  float tune_ms[4] = {0.f, 0.f, 0.f, 0.f};  // measured at creation: API image, device layout, wide image, ELL image
  uint32_t n_zero_rows = 0, run_log2 = 2;
  bool zero_all = true;
  void *d_x = nullptr, *d_y = nullptr;
  double *d_scalar = nullptr;
  uint4 *d_flush = nullptr;
  size_t flush_words = 0;
  void *h_stage = nullptr;  // pinned staging for get_y(accumulate)
  size_t h_stage_bytes = 0;
  static constexpr int kPieces = 8;  // get_y(accumulate) overlaps the host addition of piece i with the copy of i+1
  cudaEvent_t ev_piece[kPieces] = {};
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;  // device time of the last iterated call
  float last_iter_ms = 0.f;                      // ... per iteration
  double *h_scalar = nullptr;                    // pinned: scalar read-backs of the iterated callers
  void *d_cg = nullptr;                          // conjugate gradients: r, x (rows values each) + 4 scalars
  size_t d_cg_bytes = 0;
  cudaStream_t stream = nullptr;
  int sms = 148;
  uint64_t launches = 0;
  int grid_cache[11][2] = {};  // [variant][is_double] -> grid size
  // asynchronous step timing (bench): events of the last enqueue_steps()
  std::vector<cudaEvent_t> ev;
  int ev_steps = 0;
  // GPU-built engines (spmvb_engine_create_from_csr): milliseconds of the CSR upload, of the build kernels (CUDA
  // events) and of the whole call (host clock)
  float build_ms[3] = {0.f, 0.f, 0.f};
};

// the ELL image is what the next SpMV streams (an engine may hold it next to the hw_matrix stream: GPU-built engines
// keep both, and spmvb_engine_set_variant switches between them)
static inline bool ell_on(const Engine *E) {
  return E->ell && (E->variant == kVariantDefault ? E->auto_variant : E->variant) == kVariantEll;
}

// CUDA events that are destroyed on every way out of a function (an early return on an error included)
struct EventList {
  std::vector<cudaEvent_t> ev;
  explicit EventList(size_t n) : ev(n, nullptr) {}
  ~EventList() { for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e); }
  cudaEvent_t &operator[](size_t i) { return ev[i]; }
};

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail(SPMVB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
  } while (0)

// Kernel variants (spmvb_engine_set_variant): 0 = auto (7 or 8, chosen from the layout's structure, see autotune());
// 1 = DIRECT (ld.global per lane, contiguous chunk ranges, atomics only: the simple baseline);
// 6 / 7 = OCC (TMA ring, x gathered from global memory, 4 / 3 CTAs per SM); 8 = XS (x window in shared memory).
// Launch with programmatic dependent launch allowed: the kernel may start while the previous kernel of the stream
// (the row-clearing kernel of the same step) is still draining; it executes griddepcontrol.wait before it touches x/y.
// g_persist: when set, the next launch carries an access policy window - the L2 keeps the lines of [base, base + bytes)
// as persisting (the y range of the row tile being computed), everything else of that launch is streaming.
static thread_local struct { void *base; size_t bytes; } g_persist = {nullptr, 0};
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (g_persist.base && g_persist.bytes) {
    attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[1].val.accessPolicyWindow.base_ptr = g_persist.base;
    attr[1].val.accessPolicyWindow.num_bytes = g_persist.bytes;
    attr[1].val.accessPolicyWindow.hitRatio = 1.0f;
    attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kern, (KArgs)args...);
}

// bounds-checked build: the limits the kernels compare their indices with, stream-ordered before EVERY launch (the
// end-to-end paths launch the kernels per row tile / per column block without going through launch_spmv)
static int upload_check_limits(Engine *E, cudaStream_t st) {
#ifdef SPMVB_CHECK_BOUNDS
  const CheckLimits lim = {E->n_chunks, E->n_pairs, E->rows, E->x_len};
  CUDA_TRY(cudaMemcpyToSymbolAsync(g_limits, &lim, sizeof lim, 0, cudaMemcpyHostToDevice, st));
#else
  (void)E; (void)st;
#endif
  return SPMVB_OK;
}

template <typename VT, int MINB>
static int launch_occ(Engine *E, const VT *x, VT *y, cudaStream_t st, int slot, int accumulate) {
  constexpr int WARPS = 8;
  const uint4 *stream = reinterpret_cast<const uint4 *>(E->d_stream);
  auto kern = spmv_occ_kernel<VT, WARPS, MINB>;
  const size_t smem = (size_t)WARPS * 2 * (VTraits<VT>::kGroupWords * 16 * 32 + 16) + (size_t)WARPS * 16;
  int &grid = E->grid_cache[slot][sizeof(VT) == 8];
  if (grid == 0) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    grid = E->sms * std::max(per_sm, 1);
  }
  CUDA_TRY(launch_pdl(kern, grid, WARPS * 32, smem, st, stream, (const uint32_t *)E->d_rowmap, x, y, (uint32_t)E->n_chunks,
                      E->cdb, E->occ_run_log2, (accumulate ? 4u : 0u) | (E->tall ? 8u : 0u)));
  return SPMVB_OK;
}

// the wide image: chunks [chunk_base, chunk_base + n_chunks) in one launch
template <typename VT>
static int launch_wide(Engine *E, const VT *x, VT *y, cudaStream_t st, int accumulate, uint64_t chunk_base, uint64_t n_chunks) {
  constexpr int WARPS = 8, MINB = 3;
  auto kern = spmv_wide_kernel<VT, WARPS, MINB>;
  const size_t slot = (size_t)wide_chunk_bytes((int)sizeof(VT)) + 16;
  const size_t smem = (size_t)WARPS * (2 * slot + (size_t)kChunkEntries * sizeof(VT)) + (size_t)WARPS * 16;
  int &grid = E->grid_cache[kVariantWide][sizeof(VT) == 8];
  if (grid == 0) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    grid = E->sms * std::max(per_sm, 1);
  }
  if (n_chunks == 0) return SPMVB_OK;
  if (int rcl = upload_check_limits(E, st)) return rcl;
  const int64_t hints = options().wide_hints >= 0 ? options().wide_hints : (E->tall ? 3 : 0);
  CUDA_TRY(launch_pdl(kern, grid, WARPS * 32, smem, st, reinterpret_cast<const uint4 *>(E->d_stream),
                      (const uint32_t *)E->d_rowmap, x, y, (uint32_t)chunk_base, (uint32_t)n_chunks, E->cdb, E->occ_run_log2,
                      (accumulate ? 4u : 0u) | ((hints & 1) ? 8u : 0u) | ((hints & 2) ? 16u : 0u) |
                          (uint32_t)(options().diag_flags > 0 ? options().diag_flags & 96 : 0)));
  return SPMVB_OK;
}

// the ELL image: slices [slice_begin, slice_begin + n_slices) in one launch.  Two instantiations: rows of up to 8 entries
// (4 ring stages, 4 CTAs per SM) and of up to kEllMaxWidth = 16
template <typename VT, int STAGES, int WMAX, int MINB>
static int launch_ell_cfg(Engine *E, const VT *x, VT *y, cudaStream_t st, int accumulate, uint32_t slice_begin, uint32_t n_slices) {
  constexpr int WARPS = 8;
  auto kern = spmv_ell_kernel<VT, WARPS, STAGES, WMAX, MINB>;
  const size_t smem = (size_t)WARPS * STAGES * E->ell_slice_bytes + (size_t)WARPS * STAGES * 8;
  int &grid = E->grid_cache[kVariantEll][sizeof(VT) == 8];
  if (grid == 0) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    grid = E->sms * std::max(per_sm, 1);
  }
  if (n_slices == 0) return SPMVB_OK;
  if (int rcl = upload_check_limits(E, st)) return rcl;
  CUDA_TRY(launch_pdl(kern, grid, WARPS * 32, smem, st, (const uint8_t *)E->d_ell, x, y, E->rows, slice_begin, n_slices,
                      E->ell_width, E->ell_slice_bytes, accumulate ? 4u : 0u));
  return SPMVB_OK;
}
template <typename VT>
static int launch_ell(Engine *E, const VT *x, VT *y, cudaStream_t st, int accumulate, uint32_t slice_begin, uint32_t n_slices) {
  if (E->ell_width <= 8) return launch_ell_cfg<VT, 4, 8, 4>(E, x, y, st, accumulate, slice_begin, n_slices);
  return launch_ell_cfg<VT, 3, kEllMaxWidth, 1>(E, x, y, st, accumulate, slice_begin, n_slices);
}

// tile < 0: the whole matrix in one launch; otherwise row tile `tile` only (per-tile plan)
template <typename VT, int WARPS, uint32_t X_CAP, int MINB>
static int launch_xs_cfg(Engine *E, const VT *x, VT *y, cudaStream_t st, int accumulate, int tile) {
  const uint4 *stream = reinterpret_cast<const uint4 *>(E->d_stream);
  auto kern = spmv_xs_kernel<VT, WARPS, X_CAP, MINB>;
  const size_t stage = (size_t)VTraits<VT>::kGroupWords * 16 * 32 + 16;
  const size_t smem = (size_t)X_CAP + (size_t)WARPS * 2 * stage + WARPS * 16 + 16;
  int &grid = E->grid_cache[kVariantXs][sizeof(VT) == 8];
  if (grid == 0) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < MINB) return fail(SPMVB_E_CUDA, "x-window kernel: fewer resident CTAs per SM than its work plan assumes");
    grid = E->xs_ctas;
  }
  if (E->n_items == 0) return SPMVB_OK;
  if (int rcl = upload_check_limits(E, st)) return rcl;
  const XsItem *items = tile < 0 ? E->d_items : E->d_items_t;
  const uint32_t *first = tile < 0 ? E->d_cta_first : E->d_cta_first_t + (size_t)tile * (E->xs_ctas + 1);
  CUDA_TRY(launch_pdl(kern, grid, WARPS * 32, smem, st, stream, (const uint32_t *)E->d_rowmap, x, y, items, first, E->cdb,
                      E->xs_run_log2, (accumulate ? 4u : 0u) | (E->tall ? 8u : 0u) | (uint32_t)(options().diag_flags > 0 ? options().diag_flags & 48 : 0)));
  return SPMVB_OK;
}
template <typename VT>
static int launch_xs(Engine *E, const VT *x, VT *y, cudaStream_t st, int accumulate, int tile = -1) {
  constexpr bool D = sizeof(VT) == 8;  // the instantiations = xs_config() in layout.h
  switch (E->xs_cfg) {
    case 1: return launch_xs_cfg<VT, D ? 9 : 14, 64u << 10, 2>(E, x, y, st, accumulate, tile);
    case 2: return launch_xs_cfg<VT, D ? 8 : 10, 32u << 10, 3>(E, x, y, st, accumulate, tile);
    default: return launch_xs_cfg<VT, D ? 18 : 24, 128u << 10, 1>(E, x, y, st, accumulate, tile);
  }
}

template <typename VT>
static int launch_spmv(Engine *E, const VT *x, VT *y, cudaStream_t st, int accumulate) {
  if (int rcl = upload_check_limits(E, st)) return rcl;
  const uint4 *stream = reinterpret_cast<const uint4 *>(E->d_stream);
  constexpr int WARPS = 8;
  int variant = E->variant == kVariantDefault ? E->auto_variant : E->variant;
  if (ell_on(E)) {  // an ELL image has one kernel; every row is written, nothing needs clearing
    int rce = launch_ell<VT>(E, x, y, st, accumulate, 0, E->ell_slices);
    if (rce) return rce;
    E->launches++;
    CUDA_TRY(cudaGetLastError());
    return SPMVB_OK;
  }
  if (E->n_chunks == 0) return SPMVB_OK;
  if (E->n_chunks >= 0x7FFFFFFFull) return fail(SPMVB_E_RANGE, "too many chunks for one engine");
  int rc = SPMVB_OK;
  if (E->wide) {  // a wide image has one kernel
    rc = launch_wide<VT>(E, x, y, st, accumulate, 0, E->n_chunks);
  } else if (variant == kVariantDirect) {
    auto kern = spmv_direct_kernel<VT, WARPS, 4>;
    int &grid = E->grid_cache[kVariantDirect][sizeof(VT) == 8];
    if (grid == 0) {
      int per_sm = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, 0));
      grid = E->sms * std::max(per_sm, 1);
    }
    kern<<<grid, WARPS * 32, 0, st>>>(stream, E->d_rowmap, x, y, E->n_chunks, E->cdb);
  } else if (variant == kVariantXs) {
    rc = launch_xs<VT>(E, x, y, st, accumulate);
  } else if (variant == 6) {
    rc = launch_occ<VT, 4>(E, x, y, st, 6, accumulate);
  } else if (variant == 7) {
    rc = launch_occ<VT, 3>(E, x, y, st, 7, accumulate);
  } else {
    rc = launch_occ<VT, 3>(E, x, y, st, 7, accumulate);
  }
  if (rc) return rc;
  E->launches++;
  CUDA_TRY(cudaGetLastError());
  return SPMVB_OK;
}

// y = A x needs y prepared only where the kernel uses atomics or writes nothing: either the listed rows or all of y
static int zero_y(Engine *E, void *y, cudaStream_t st) {
  if (ell_on(E)) return SPMVB_OK;  // the ELL kernel writes every row
  const int variant = E->variant == kVariantDefault ? E->auto_variant : E->variant;
  if (E->zero_all || variant == kVariantDirect) {
    CUDA_TRY(cudaMemsetAsync(y, 0, (size_t)E->rows * E->vb, st));
  } else if (E->n_zero_rows) {
    const int grid = (int)std::min<uint64_t>(((uint64_t)E->n_zero_rows + 255) / 256, (uint64_t)E->sms * 8);
    if (E->is_double) zero_rows_kernel<double><<<grid, 256, 0, st>>>((double *)y, E->d_zero_rows, E->n_zero_rows);
    else zero_rows_kernel<float><<<grid, 256, 0, st>>>((float *)y, E->d_zero_rows, E->n_zero_rows);
    E->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  return SPMVB_OK;
}

static int do_spmv(Engine *E, const void *x_dev, void *y_dev, int accumulate, cudaStream_t st) {
  const void *x = x_dev ? x_dev : E->d_x;
  void *y = y_dev ? y_dev : E->d_y;
  if (!accumulate) {
    int rc = zero_y(E, y, st);
    if (rc) return rc;
  }
  const int variant = E->variant == kVariantDefault ? E->auto_variant : E->variant;
  if (variant == kVariantXs && E->n_tiles > 1 && options().tile_launch > 0) {
    // one launch per row tile, each with the tile's y range as the persisting window of the L2 cache
    uint32_t prev = 0;
    for (int k = 0; k < E->n_tiles; k++) {
      const uint32_t end = k + 1 == E->n_tiles ? E->rows : E->tile_rows_end[k];
      const uint32_t slack = (end - prev) / 16 + 1024;
      const uint32_t lo = prev > slack ? prev - slack : 0, hi = (uint64_t)end + slack < E->rows ? end + slack : E->rows;
      g_persist.base = (uint8_t *)y + (size_t)lo * E->vb;
      g_persist.bytes = (size_t)(hi - lo) * E->vb;
      int rc = E->is_double ? launch_xs<double>(E, (const double *)x, (double *)y, st, accumulate, k)
                            : launch_xs<float>(E, (const float *)x, (float *)y, st, accumulate, k);
      g_persist.base = nullptr; g_persist.bytes = 0;
      if (rc) return rc;
      E->launches++;
      prev = std::max(prev, end);
    }
    CUDA_TRY(cudaGetLastError());
    return SPMVB_OK;
  }
  if (E->is_double) return launch_spmv<double>(E, (const double *)x, (double *)y, st, accumulate);
  return launch_spmv<float>(E, (const float *)x, (float *)y, st, accumulate);
}

// variant 0: pick between the two production kernels from the structure of the layout (deterministic).  The
// shared-memory x window pays when the (row, block) pairs are short and their columns scattered - then the global
// gathers touch one 128-byte line per entry, 2 cycles of L1 tag time each - and it is required for tall matrices
// (CU-major order, x streamed once per row tile).  Banded matrices keep the global gathers with more resident warps.
// The same criterion (layout_is_irregular: distinct x lines per chunk) decides the engine-private device layout (plan_device_params), so that the
// windows of the chosen kernel fit.  Where that rule would take the x-window kernel the two kernels are also TIMED on the
// actual matrix (three SpMVs each; option autotune = 0 keeps the rule's choice, e.g. under a profiler where timings are
// noise): the rule is right for whole matrices, but a row shard of an irregular matrix - one GPU's part of a group - has
// to load every x window for an eighth of the entries, and the global-gather kernel is the faster one there (R-MAT scale
// 24 in fp32, 8 shards: 0.25-0.33 ms per shard with the x-window kernel against 1.07 ms for the whole matrix).
static int autotune(Engine *E) {
  if (E->wide) { E->auto_variant = kVariantWide; return SPMVB_OK; }
  E->auto_variant = kVariantOcc3;
  if (E->n_chunks == 0 || E->xs_windowed_frac < 0.5) return SPMVB_OK;
  if (E->cu_major || E->irregular) E->auto_variant = kVariantXs;
  if (options().autotune == 0 || E->auto_variant != kVariantXs || E->variant != kVariantDefault) return SPMVB_OK;
  const int cand[2] = {kVariantOcc3, kVariantXs};
  cudaEvent_t a, b;
  CUDA_TRY(cudaEventCreate(&a));
  CUDA_TRY(cudaEventCreate(&b));
  const int saved = E->variant;
  float best = 1e30f;
  int rc = SPMVB_OK;
  for (int c = 0; c < 2 && rc == SPMVB_OK; c++) {
    E->variant = cand[c];
    for (int rep = 0; rep < 3 && rc == SPMVB_OK; rep++) {
      cudaEventRecord(a, E->stream);
      rc = do_spmv(E, nullptr, nullptr, 0, E->stream);  // a whole step: clear listed rows + kernel
      cudaEventRecord(b, E->stream);
      if (rc) break;
      if (cudaEventSynchronize(b) != cudaSuccess) { rc = fail(SPMVB_E_CUDA, "autotune"); break; }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, a, b);
      if (rep) E->tune_ms[c] = rep == 1 ? ms : std::min(E->tune_ms[c], ms);
    }
    if (rc == SPMVB_OK && E->tune_ms[c] < best) { best = E->tune_ms[c]; E->auto_variant = cand[c]; }
  }
  E->variant = saved;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  E->launches = 0;
  return rc;
}

// host y (+)= device y for rows [a, b): the addition of spmv_hw (csr_hw.cpp:1557) on all host cores
static void host_accumulate(Engine *E, void *y_host, int64_t a, int64_t b) {
  if (E->is_double) {
    double *dst = (double *)y_host; const double *src = (const double *)E->h_stage;
#pragma omp parallel for schedule(static)
    for (int64_t i = a; i < b; i++) dst[i] += src[i];
  } else {
    float *dst = (float *)y_host; const float *src = (const float *)E->h_stage;
#pragma omp parallel for schedule(static)
    for (int64_t i = a; i < b; i++) dst[i] += src[i];
  }
}

// spmv_hw end to end for a row-tiled (CU-major) layout: x up, then one kernel launch per row tile; as soon as a tile
// is done its final rows of y travel to the host on a second stream while the next tiles are computed, and the host adds
// them into y_host while later rows are still on their way.
template <typename VT>
static int spmv_host_tiled(Engine *E, void *y_host, int accumulate) {
  const size_t bytes = (size_t)E->rows * E->vb;
  if (accumulate && E->h_stage_bytes < bytes) {
    if (E->h_stage) cudaFreeHost(E->h_stage);
    E->h_stage = nullptr; E->h_stage_bytes = 0;
    CUDA_TRY(cudaMallocHost(&E->h_stage, bytes));
    E->h_stage_bytes = bytes;
  }
  int rc = zero_y(E, E->d_y, E->stream);
  if (rc) return rc;
  uint8_t *dst = accumulate ? (uint8_t *)E->h_stage : (uint8_t *)y_host;
  uint32_t prev = 0;
  for (int k = 0; k < E->n_tiles; k++) {
    rc = launch_xs<VT>(E, (const VT *)E->d_x, (VT *)E->d_y, E->stream, 0, k);
    if (rc) return rc;
    E->launches++;
    CUDA_TRY(cudaEventRecord(E->ev_tile[2 * k], E->stream));
    CUDA_TRY(cudaStreamWaitEvent(E->copy_stream, E->ev_tile[2 * k], 0));
    const uint32_t end = k + 1 == E->n_tiles ? E->rows : E->tile_rows_end[k];
    if (end > prev)
      CUDA_TRY(cudaMemcpyAsync(dst + (size_t)prev * E->vb, (const uint8_t *)E->d_y + (size_t)prev * E->vb,
                               (size_t)(end - prev) * E->vb, cudaMemcpyDeviceToHost, E->copy_stream));
    CUDA_TRY(cudaEventRecord(E->ev_tile[2 * k + 1], E->copy_stream));
    prev = std::max(prev, end);
  }
  cudaError_t werr = cudaSuccess;
  prev = 0;
  for (int k = 0; k < E->n_tiles; k++) {
    cudaError_t r = cudaEventSynchronize(E->ev_tile[2 * k + 1]);
    if (r != cudaSuccess) werr = r;
    const uint32_t end = k + 1 == E->n_tiles ? E->rows : E->tile_rows_end[k];
    if (werr == cudaSuccess && accumulate && end > prev) host_accumulate(E, y_host, prev, end);
    prev = std::max(prev, end);
  }
  if (werr != cudaSuccess) return fail(SPMVB_E_CUDA, std::string("spmv_host: ") + cudaGetErrorString(werr));
  return SPMVB_OK;
}

// spmv_hw end to end over an ELL image.  Slices are rows: tile k of the slices can run as soon as the columns ITS rows
// read are on the device (a band or a stencil reads a window of x that moves with the rows), and its rows of y are final
// when it is done.  Three streams: x pieces go up, kernels run, y tiles come down - PCIe carries both directions at
// once - and the host adds the tiles that have arrived (accumulate) while the rest is still on its way.
template <typename VT>
static int spmv_host_ell(Engine *E, const void *x_host, uint32_t n, void *y_host, int accumulate) {
  const int T = (int)E->ell_tile_slice.size() - 1;
  if (!E->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&E->copy_stream, cudaStreamNonBlocking));
  if (!E->down_stream) CUDA_TRY(cudaStreamCreateWithFlags(&E->down_stream, cudaStreamNonBlocking));
  while (E->ev_block.size() < (size_t)3 * T + 1) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    E->ev_block.push_back(e);
  }
  const size_t bytes = (size_t)E->rows * E->vb;
  if (accumulate && E->h_stage_bytes < bytes) {
    if (E->h_stage) cudaFreeHost(E->h_stage);
    E->h_stage = nullptr; E->h_stage_bytes = 0;
    CUDA_TRY(cudaMallocHost(&E->h_stage, bytes));
    E->h_stage_bytes = bytes;
  }
  uint8_t *dst = accumulate ? (uint8_t *)E->h_stage : (uint8_t *)y_host;
  const uint64_t m = std::min<uint32_t>(n, E->expanded_cols);
  // the upload must not overtake a kernel of an earlier call that still reads x
  CUDA_TRY(cudaEventRecord(E->ev_block[3 * T], E->stream));
  CUDA_TRY(cudaStreamWaitEvent(E->copy_stream, E->ev_block[3 * T], 0));
  uint64_t have = E->ell_x_begin;  // x[ell_x_begin, have) is on its way
  for (int k = 0; k < T; k++) {
    const uint64_t want = std::min<uint64_t>(E->ell_tile_x_end[k], E->x_len);
    if (want > have) {
      const uint64_t up = std::min(want, m);
      if (up > have)
        CUDA_TRY(cudaMemcpyAsync((uint8_t *)E->d_x + have * E->vb, (const uint8_t *)x_host + have * E->vb,
                                 (size_t)(up - have) * E->vb, cudaMemcpyHostToDevice, E->copy_stream));
      if (up < want)  // zero padding behind a short x (csr_hw.cpp:1478-1481)
        CUDA_TRY(cudaMemsetAsync((uint8_t *)E->d_x + std::max(up, have) * E->vb, 0, (size_t)(want - std::max(up, have)) * E->vb,
                                 E->copy_stream));
      have = want;
    }
    CUDA_TRY(cudaEventRecord(E->ev_block[3 * k], E->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(E->stream, E->ev_block[3 * k], 0));
    const uint32_t s0 = E->ell_tile_slice[k], s1 = E->ell_tile_slice[k + 1];
    int rc = launch_ell<VT>(E, (const VT *)E->d_x, (VT *)E->d_y, E->stream, 0, s0, s1 - s0);
    if (rc) return rc;
    E->launches++;
    CUDA_TRY(cudaEventRecord(E->ev_block[3 * k + 1], E->stream));
    CUDA_TRY(cudaStreamWaitEvent(E->down_stream, E->ev_block[3 * k + 1], 0));
    const uint64_t r0 = (uint64_t)s0 * kEllSliceRows, r1 = std::min<uint64_t>((uint64_t)s1 * kEllSliceRows, E->rows);
    if (r1 > r0)
      CUDA_TRY(cudaMemcpyAsync(dst + r0 * E->vb, (const uint8_t *)E->d_y + r0 * E->vb, (size_t)(r1 - r0) * E->vb,
                               cudaMemcpyDeviceToHost, E->down_stream));
    CUDA_TRY(cudaEventRecord(E->ev_block[3 * k + 2], E->down_stream));
  }
  CUDA_TRY(cudaGetLastError());
  cudaError_t werr = cudaSuccess;
  for (int k = 0; k < T; k++) {
    cudaError_t r = cudaEventSynchronize(E->ev_block[3 * k + 2]);
    if (r != cudaSuccess) werr = r;
    const int64_t r0 = (int64_t)E->ell_tile_slice[k] * kEllSliceRows;
    const int64_t r1 = std::min<int64_t>((int64_t)E->ell_tile_slice[k + 1] * kEllSliceRows, E->rows);
    if (werr == cudaSuccess && accumulate && r1 > r0) host_accumulate(E, y_host, r0, r1);
  }
  if (werr != cudaSuccess) return fail(SPMVB_E_CUDA, std::string("spmv_host: ") + cudaGetErrorString(werr));
  return SPMVB_OK;
}

// spmv_hw end to end over a wide image: the column blocks are walked one after the other, so block b can be computed
// as soon as ITS range of x has arrived - the upload of the next ranges runs on a second stream under the kernels.
template <typename VT>
static int spmv_host_wide(Engine *E, const void *x_host, uint32_t n) {
  if (!E->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&E->copy_stream, cudaStreamNonBlocking));
  while (E->ev_block.size() < (size_t)E->blocks + 1) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    E->ev_block.push_back(e);
  }
  const uint64_t m = std::min<uint32_t>(n, E->expanded_cols);
  // the copy stream must not overtake a kernel of the previous call that still reads x
  CUDA_TRY(cudaEventRecord(E->ev_block[E->blocks], E->stream));
  CUDA_TRY(cudaStreamWaitEvent(E->copy_stream, E->ev_block[E->blocks], 0));
  int rc = zero_y(E, E->d_y, E->stream);
  if (rc) return rc;
  for (int b = 0; b < E->blocks; b++) {
    const uint64_t first = (uint64_t)b * E->cdb, end = first + E->cdb;
    const uint64_t c0 = E->block_chunk0[b], c1 = E->block_chunk0[b + 1];
    if (c1 == c0) continue;  // nothing reads this range
    const uint64_t up = std::min(end, m);
    if (up > first)
      CUDA_TRY(cudaMemcpyAsync((uint8_t *)E->d_x + first * E->vb, (const uint8_t *)x_host + first * E->vb,
                               (size_t)(up - first) * E->vb, cudaMemcpyHostToDevice, E->copy_stream));
    if (up < end)  // zero padding behind a short x (csr_hw.cpp:1478-1481)
      CUDA_TRY(cudaMemsetAsync((uint8_t *)E->d_x + std::max(up, first) * E->vb, 0, (size_t)(end - std::max(up, first)) * E->vb,
                               E->copy_stream));
    CUDA_TRY(cudaEventRecord(E->ev_block[b], E->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(E->stream, E->ev_block[b], 0));
    rc = launch_wide<VT>(E, (const VT *)E->d_x, (VT *)E->d_y, E->stream, 0, c0, c1 - c0);
    if (rc) return rc;
    E->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return SPMVB_OK;
}

}  // namespace spmvb

using namespace spmvb;

namespace {
// Conjugate gradients (spmvb_engine_cg).  The work vectors r, x and the four scalars live in one engine-owned
// allocation that is made on the first call and reused; the scalar read-backs go through pinned memory.  The device
// time of the iteration loop is taken with two events (spmvb_engine_last_iter_ms).
template <typename VT>
int cg_impl(Engine *E, const VT *b_host, VT *x_host, int max_iters, double rel_tol, int *iters_out, double *relres_out) {
  const uint32_t n = E->rows;
  const size_t bytes = ((size_t)n * sizeof(VT) + 255) & ~(size_t)255;
  if (E->d_cg_bytes < 2 * bytes + 64) {
    cudaFree(E->d_cg);
    E->d_cg = nullptr; E->d_cg_bytes = 0;
    CUDA_TRY(cudaMalloc(&E->d_cg, 2 * bytes + 64));
    E->d_cg_bytes = 2 * bytes + 64;
  }
  VT *d_r = (VT *)E->d_cg, *d_xs = (VT *)((uint8_t *)E->d_cg + bytes);
  double *s = (double *)((uint8_t *)E->d_cg + 2 * bytes);  // s[0], s[2]: r.r (roles swap every iteration); s[1]: p.q
  cudaStream_t st = E->stream;
  const int grid = E->sms * 4;
  int it = 0;
  VT *p = (VT *)E->d_x, *q = (VT *)E->d_y;
  // x0 = 0: r = b, p = b (the padding of p beyond n stays zero), rr = b.b
  CUDA_TRY(cudaMemsetAsync(d_xs, 0, (size_t)n * sizeof(VT), st));
  CUDA_TRY(cudaMemcpyAsync(d_r, b_host, (size_t)n * sizeof(VT), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(E->d_x, 0, E->x_len * sizeof(VT), st));
  CUDA_TRY(cudaMemcpyAsync(p, d_r, (size_t)n * sizeof(VT), cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemsetAsync(s, 0, 4 * sizeof(double), st));
  dot_kernel<VT><<<grid, 256, 0, st>>>(d_r, d_r, n, s + 0);
  E->launches++;
  CUDA_TRY(cudaMemcpyAsync(E->h_scalar, s + 0, sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const double bnorm2 = E->h_scalar[0];
  double rr_host = bnorm2;
  E->last_iter_ms = 0.f;
  if (bnorm2 == 0.0) {  // b = 0: x = 0
    memset(x_host, 0, (size_t)n * sizeof(VT));
    if (iters_out) *iters_out = 0;
    if (relres_out) *relres_out = 0.0;
    return SPMVB_OK;
  }
  const double stop2 = rel_tol * rel_tol * bnorm2;
  int cur = 0;  // index of the current r.r
  constexpr int kCheckEvery = 8;
  CUDA_TRY(cudaEventRecord(E->ev_t0, st));
  while (it < max_iters) {
    const int nxt = 2 - cur;
    int r2 = do_spmv(E, p, q, 0, st);  // q = A p
    if (r2) return r2;
    CUDA_TRY(cudaMemsetAsync(s + 1, 0, sizeof(double), st));
    CUDA_TRY(cudaMemsetAsync(s + nxt, 0, sizeof(double), st));
    dot_kernel<VT><<<grid, 256, 0, st>>>(p, q, n, s + 1);
    cg_update_kernel<VT><<<grid, 256, 0, st>>>(d_xs, d_r, p, q, n, s + cur, s + 1, s + nxt);
    cg_direction_kernel<VT><<<grid, 256, 0, st>>>(p, d_r, n, s + cur, s + nxt);
    E->launches += 3;
    CUDA_TRY(cudaGetLastError());
    cur = nxt;
    it++;
    if (it % kCheckEvery == 0 || it == max_iters) {
      CUDA_TRY(cudaMemcpyAsync(E->h_scalar, s + cur, sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      rr_host = E->h_scalar[0];
      if (!(rr_host > stop2)) break;  // converged (or NaN: give up)
    }
  }
  CUDA_TRY(cudaEventRecord(E->ev_t1, st));
  CUDA_TRY(cudaMemcpyAsync(x_host, d_xs, (size_t)n * sizeof(VT), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (it > 0) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, E->ev_t0, E->ev_t1) == cudaSuccess) E->last_iter_ms = ms / (float)it;
  }
  if (iters_out) *iters_out = it;
  if (relres_out) *relres_out = std::sqrt(rr_host / bnorm2);
  return SPMVB_OK;
}
}  // namespace

extern "C" {

// device checks, the Engine object with everything that follows from the layout's tables, stream, x / y
static int engine_open(int device, int variant, Engine **out) {
  *out = nullptr;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(SPMVB_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(ce));
  if (device < 0 || device >= ndev) return fail(SPMVB_E_ARG, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(SPMVB_E_CUDA, "an sm_100-class GPU (B200) is required; there is no fallback path");
  Engine *E = new Engine();
  E->device = device; E->variant = variant;
  E->sms = prop.multiProcessorCount;
  cudaError_t e = cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc((void **)&E->d_scalar, 64);
  if (e == cudaSuccess) e = cudaMallocHost((void **)&E->h_scalar, 64);
  for (int p = 0; p < Engine::kPieces && e == cudaSuccess; p++) e = cudaEventCreateWithFlags(&E->ev_piece[p], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreate(&E->ev_t0);
  if (e == cudaSuccess) e = cudaEventCreate(&E->ev_t1);
  if (e != cudaSuccess) {
    std::string msg = std::string("engine_create: ") + cudaGetErrorString(e);
    spmvb_engine_free((spmvb_engine *)E);
    return fail(SPMVB_E_CUDA, msg);
  }
  *out = E;
  return SPMVB_OK;
}

// everything the engine derives from the layout's host-side tables + the x / y vectors
static int engine_adopt_layout(Engine *E, const Layout *L) {
  E->is_double = L->is_double; E->vb = L->vb;
  E->rows = L->rows; E->cols = L->cols; E->expanded_cols = L->expanded_cols; E->cdb = L->cdb; E->blocks = L->blocks;
  E->real_nnz = L->real_nnz; E->n_chunks = L->n_chunks; E->n_pairs = L->n_pairs; E->stream_bytes = L->stream_bytes;
  E->cu_major = L->cu_major; E->dev_cu = L->cu; E->dev_vf = L->vf;
  E->irregular = layout_is_irregular(L);
  E->wide = L->is_wide;
  if (L->is_wide) {
    E->block_chunk0.assign((size_t)L->blocks + 1, L->n_chunks);
    for (int b = 0; b < L->blocks; b++) E->block_chunk0[b] = L->piece_chunk0[b];
  }
  E->xs_cfg = L->xs_cfg;
  E->xs_ctas = E->sms * xs_config(L->is_double, L->xs_cfg).ctas_per_sm;
  E->tall = (uint64_t)L->rows * L->vb > ((uint64_t)48 << 20) && (uint64_t)L->cols * L->vb > ((uint64_t)48 << 20);
  if (options().tall >= 0) E->tall = options().tall != 0;
  if (E->tall && options().l2_persist_mb > 0) {
    // evict-last lines only outlive the stream if the L2 has a set-aside portion for them (off by default)
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, E->device));
    const size_t want = std::min<size_t>((size_t)options().l2_persist_mb << 20, (size_t)prop.persistingL2CacheMaxSize);
    CUDA_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
    fprintf(stderr, "[spmvb] persisting L2 set-aside: %zu MB (device maximum %d MB)\n", want >> 20, prop.persistingL2CacheMaxSize >> 20);
  }
  E->x_touched = 0;
  for (int b = 0; b < L->blocks; b++) {
    uint64_t nz = 0;
    for (int k = 0; k < L->cu; k++) nz += L->piece_real_nnz[(size_t)b * L->cu + k];
    if (nz) E->x_touched += std::min<uint64_t>(L->cdb, (uint64_t)L->cols - (uint64_t)b * L->cdb);
  }
  E->x_len = (uint64_t)L->blocks * L->cdb;  // >= expanded_cols: any 15-bit index of any block stays in range
  {
    const int64_t nr = spmvb_layout_x_ranges((const spmvb_layout *)L, nullptr, 0);
    E->x_ranges.assign((size_t)(nr > 0 ? 2 * nr : 0), 0);
    if (nr > 0) spmvb_layout_x_ranges((const spmvb_layout *)L, E->x_ranges.data(), (uint64_t)nr);
  }
  E->zero_all = L->zero_all; E->n_zero_rows = (uint32_t)L->zero_rows.size(); E->run_log2 = (uint32_t)L->run_log2;
  if (options().occ_run_log2 >= 0) E->occ_run_log2 = (uint32_t)options().occ_run_log2;
  if (options().xs_run_log2 >= 0) E->xs_run_log2 = (uint32_t)options().xs_run_log2;
  // rows split across run boundaries are only cleared at the layout's granularity: runs must be multiples of it;
  // and a run is at least two chunks (the walk refills two chunks ahead)
  E->occ_run_log2 = std::min(8u, std::max(1u, std::max(E->occ_run_log2, E->run_log2)));
  E->xs_run_log2 = std::min(8u, std::max(1u, std::max(E->xs_run_log2, E->run_log2)));
  CUDA_TRY(cudaMalloc(&E->d_x, E->x_len * E->vb));
  CUDA_TRY(cudaMalloc(&E->d_y, (size_t)E->rows * E->vb));
  return SPMVB_OK;
}

// work plan of the XS kernel, cleared vectors, kernel choice; the image must be on the device (or on its way, on E->stream)
static int engine_finish(Engine *E, const Layout *L) {
  std::vector<XsItem> items;
  std::vector<uint32_t> cta_first;
  XsTilePlan tiles;
  if (L->is_wide) cta_first.assign((size_t)E->xs_ctas + 1, 0);  // no shared-memory windows over a wide image
  else build_xs_items(L, E->xs_ctas, E->xs_run_log2, items, cta_first, L->cu_major ? &tiles : nullptr);
  E->n_items = (uint32_t)items.size();
  if (tiles.n_tiles > 1 && tiles.n_tiles <= 1024 && !tiles.items.empty()) {
    E->n_tiles = tiles.n_tiles;
    E->tile_rows_end = tiles.rows_end;
    CUDA_TRY(cudaMalloc((void **)&E->d_items_t, tiles.items.size() * sizeof(XsItem)));
    CUDA_TRY(cudaMalloc((void **)&E->d_cta_first_t, tiles.cta_first.size() * 4));
    CUDA_TRY(cudaMemcpy(E->d_items_t, tiles.items.data(), tiles.items.size() * sizeof(XsItem), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(E->d_cta_first_t, tiles.cta_first.data(), tiles.cta_first.size() * 4, cudaMemcpyHostToDevice));
    E->ev_tile.assign(2 * (size_t)E->n_tiles, nullptr);
    for (auto &e : E->ev_tile) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_TRY(cudaStreamCreateWithFlags(&E->copy_stream, cudaStreamNonBlocking));
  }
  CUDA_TRY(cudaMalloc((void **)&E->d_items, std::max<size_t>(items.size(), 1) * sizeof(XsItem)));
  CUDA_TRY(cudaMalloc((void **)&E->d_cta_first, cta_first.size() * 4));
  CUDA_TRY(cudaMemcpy(E->d_items, items.data(), items.size() * sizeof(XsItem), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(E->d_cta_first, cta_first.data(), cta_first.size() * 4, cudaMemcpyHostToDevice));
  uint64_t windowed = 0;
  for (auto &it : items) windowed += it.x_bytes ? it.chunk_count : 0;
  E->xs_windowed_frac = L->n_chunks ? (double)windowed / (double)L->n_chunks : 0.0;
  CUDA_TRY(cudaMemsetAsync(E->d_x, 0, E->x_len * E->vb, E->stream));
  CUDA_TRY(cudaMemsetAsync(E->d_y, 0, (size_t)E->rows * E->vb, E->stream));
  CUDA_TRY(cudaStreamSynchronize(E->stream));
  int rc = autotune(E);
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(E->d_y, 0, (size_t)E->rows * E->vb, E->stream));
  CUDA_TRY(cudaStreamSynchronize(E->stream));
  return SPMVB_OK;
}

// uploads ONE layout (the device image of L itself) and makes it runnable
static int engine_create_single(const Layout *L, int device, int variant, Engine **out) {
  *out = nullptr;
  Engine *E = nullptr;
  int rc = engine_open(device, variant, &E);
  if (rc) return rc;
  auto upload = [&]() -> int {
    int r = engine_adopt_layout(E, L);
    if (r) return r;
    // device image: one slot per chunk = the chunk's words (bit-exact hw_matrix bytes) followed by its 16-byte
    // ChunkMeta, so that a single bulk copy brings both into shared memory
    const size_t slot = (size_t)L->chunk_bytes + sizeof(ChunkMeta);
    CUDA_TRY(cudaMalloc((void **)&E->d_stream, std::max<uint64_t>(L->n_chunks * slot, 16)));
    CUDA_TRY(cudaMalloc((void **)&E->d_rowmap, (L->n_pairs + 8) * 4));
    CUDA_TRY(cudaMalloc((void **)&E->d_zero_rows, std::max<size_t>(L->zero_rows.size(), 1) * 4));
    if (L->n_chunks) {
      CUDA_TRY(cudaMemcpy2DAsync(E->d_stream, slot, L->stream, L->chunk_bytes, L->chunk_bytes, L->n_chunks,
                                 cudaMemcpyHostToDevice, E->stream));
      CUDA_TRY(cudaMemcpy2DAsync(E->d_stream + L->chunk_bytes, slot, L->chunks, sizeof(ChunkMeta), sizeof(ChunkMeta),
                                 L->n_chunks, cudaMemcpyHostToDevice, E->stream));
    }
    CUDA_TRY(cudaMemsetAsync(E->d_rowmap + L->n_pairs, 0, 8 * 4, E->stream));
    CUDA_TRY(cudaMemcpyAsync(E->d_rowmap, L->rowmap, L->n_pairs * 4, cudaMemcpyHostToDevice, E->stream));
    CUDA_TRY(cudaMemcpyAsync(E->d_zero_rows, L->zero_rows.data(), L->zero_rows.size() * 4, cudaMemcpyHostToDevice, E->stream));
    return engine_finish(E, L);
  };
  rc = upload();
  if (rc) { spmvb_engine_free((spmvb_engine *)E); return rc; }
  *out = E;
  return SPMVB_OK;
}

static int engine_time_step(Engine *E, int reps, float *best_ms);

// switches an engine whose vectors and sizes are set up (engine_adopt_layout) over to an ELL image that is on the device
// already: kernel choice, row tiles of the end-to-end pipeline (equal slice counts; x[.., end) each tile needs = running
// maximum of the slices' last columns - whatever lies below the window of a later tile is on the device by then)
static void engine_set_ell(Engine *E, uint8_t *d_image, uint32_t n_slices, uint32_t width, uint32_t slice_bytes,
                           const uint32_t *col_lo, const uint32_t *col_hi) {
  E->ell = true;
  E->d_ell = d_image;
  E->ell_slices = n_slices; E->ell_width = width; E->ell_slice_bytes = slice_bytes;
  E->auto_variant = kVariantEll;
  int T = options().ell_tiles > 0 ? (int)std::min<int64_t>(options().ell_tiles, 256) : 8;
  T = (int)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)T, n_slices));
  E->ell_tile_slice.assign((size_t)T + 1, 0);
  for (int k = 0; k <= T; k++) E->ell_tile_slice[k] = (uint32_t)((uint64_t)n_slices * k / T);
  E->ell_tile_x_end.assign((size_t)T, 0);
  uint64_t lo_all = E->x_len, hi_run = 0;
  for (uint32_t s = 0; s < n_slices; s++)
    if (col_lo[s] <= col_hi[s]) lo_all = std::min<uint64_t>(lo_all, col_lo[s]);
  for (int k = 0; k < T; k++) {
    for (uint32_t s = E->ell_tile_slice[k]; s < E->ell_tile_slice[k + 1]; s++)
      if (col_lo[s] <= col_hi[s]) hi_run = std::max<uint64_t>(hi_run, (uint64_t)col_hi[s] + 1);
    E->ell_tile_x_end[k] = (hi_run + 63) & ~(uint64_t)63;  // whole 512-byte pieces
  }
  E->ell_x_begin = lo_all == E->x_len ? 0 : (lo_all & ~(uint64_t)63);
}

// an engine over the host-built ELL image of layout A (sizes, x ranges and vectors follow the API layout; the hw_matrix
// stream itself is not uploaded)
static int engine_create_ell(const Layout *A, int device, Engine **out) {
  *out = nullptr;
  const EllImage *I = A->ell;
  Engine *E = nullptr;
  int rc = engine_open(device, kVariantDefault, &E);
  if (rc) return rc;
  auto upload = [&]() -> int {
    int r = engine_adopt_layout(E, A);
    if (r) return r;
    uint8_t *d_image = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d_image, std::max<uint64_t>(I->bytes, 16)));
    engine_set_ell(E, d_image, I->n_slices, I->width, I->slice_bytes, I->col_lo.data(), I->col_hi.data());
    CUDA_TRY(cudaMemcpyAsync(E->d_ell, I->image, I->bytes, cudaMemcpyHostToDevice, E->stream));
    CUDA_TRY(cudaMemsetAsync(E->d_x, 0, E->x_len * E->vb, E->stream));
    CUDA_TRY(cudaMemsetAsync(E->d_y, 0, (size_t)E->rows * E->vb, E->stream));
    CUDA_TRY(cudaStreamSynchronize(E->stream));
    return SPMVB_OK;
  };
  rc = upload();
  if (rc) { spmvb_engine_free((spmvb_engine *)E); return rc; }
  *out = E;
  return SPMVB_OK;
}

// GPU-built engines (spmvb_engine_create_from_csr): the ELL image built on the device from the device-resident CSR
// (ell_gpu.cuh), timed against the stream image the engine already runs; the faster stays.  The API image is kept for
// spmvb_engine_fetch_layout either way.
static int engine_try_ell_gpu(Engine *E, const Layout *L, uint32_t rows, uint64_t nnz, const uint64_t *rp, const uint32_t *ci,
                              const void *va) {
  if (options().ell == 0 || nnz == 0) return SPMVB_OK;
  const uint32_t n_slices = (rows + kEllSliceRows - 1) / kEllSliceRows;
  uint32_t *d_lohi = nullptr; EllScan *d_scan = nullptr; uint8_t *d_image = nullptr;
  auto run = [&]() -> int {
    CUDA_TRY(cudaMalloc((void **)&d_lohi, (size_t)n_slices * 8));
    CUDA_TRY(cudaMalloc((void **)&d_scan, sizeof(EllScan)));
    CUDA_TRY(cudaMemsetAsync(d_scan, 0, sizeof(EllScan), E->stream));
    const int grid = (int)(((uint64_t)n_slices * 32 + 255) / 256);
    ell_scan_kernel<<<grid, 256, 0, E->stream>>>(rows, rp, ci, n_slices, d_lohi, d_lohi + n_slices, d_scan);
    EllScan scan;
    CUDA_TRY(cudaMemcpyAsync(&scan, d_scan, sizeof scan, cudaMemcpyDeviceToHost, E->stream));
    CUDA_TRY(cudaStreamSynchronize(E->stream));
    const uint64_t slots = (uint64_t)n_slices * kEllSliceRows * scan.width;
    if (scan.width == 0 || scan.width > (uint32_t)kEllMaxWidth || scan.bad) return SPMVB_OK;
    if (options().ell < 1 && (double)slots > 1.04 * (double)nnz) return SPMVB_OK;
    const uint32_t sb = ell_slice_bytes(scan.width, E->vb);
    CUDA_TRY(cudaMalloc((void **)&d_image, std::max<uint64_t>((uint64_t)n_slices * sb, 16)));
    if (E->is_double)
      ell_fill_kernel<double><<<grid, 256, 0, E->stream>>>(rows, rp, ci, (const double *)va, n_slices, scan.width, sb, d_lohi,
                                                            d_lohi + n_slices, d_image);
    else
      ell_fill_kernel<float><<<grid, 256, 0, E->stream>>>(rows, rp, ci, (const float *)va, n_slices, scan.width, sb, d_lohi,
                                                           d_lohi + n_slices, d_image);
    CUDA_TRY(cudaGetLastError());
    std::vector<uint32_t> lohi((size_t)n_slices * 2);
    CUDA_TRY(cudaMemcpyAsync(lohi.data(), d_lohi, lohi.size() * 4, cudaMemcpyDeviceToHost, E->stream));
    CUDA_TRY(cudaStreamSynchronize(E->stream));
    float t_cur = 0.f, t_ell = 0.f;
    int rc = engine_time_step(E, 5, &t_cur);
    if (rc) return rc;
    const int auto_before = E->auto_variant;
    engine_set_ell(E, d_image, n_slices, scan.width, sb, lohi.data(), lohi.data() + n_slices);
    rc = engine_time_step(E, 5, &t_ell);
    if (rc) return rc;
    E->tune_ms[E->wide ? 2 : (L->dev ? 1 : 0)] = t_cur; E->tune_ms[3] = t_ell;
    if (t_ell < t_cur || options().ell == 1) {
      d_image = nullptr;  // owned by the engine now; the stream image stays (spmvb_engine_set_variant, fetch_layout)
    } else {
      E->ell = false; E->d_ell = nullptr; E->ell_slices = 0; E->ell_width = 0; E->ell_slice_bytes = 0;
      E->auto_variant = auto_before;
      E->ell_tile_slice.clear(); E->ell_tile_x_end.clear();
    }
    return SPMVB_OK;
  };
  const int rc = run();
  cudaFree(d_lohi); cudaFree(d_scan); cudaFree(d_image);
  return rc;
}

// milliseconds of one y = A x (clear rows + kernel) with the engine's own kernel choice, best of `reps` after a warm-up
static int engine_time_step(Engine *E, int reps, float *best_ms) {
  *best_ms = 1e30f;
  for (int rep = 0; rep <= reps; rep++) {
    CUDA_TRY(cudaEventRecord(E->ev_t0, E->stream));
    int rc = do_spmv(E, nullptr, nullptr, 0, E->stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(E->ev_t1, E->stream));
    CUDA_TRY(cudaEventSynchronize(E->ev_t1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, E->ev_t0, E->ev_t1));
    if (rep) *best_ms = std::min(*best_ms, ms);
  }
  E->launches = 0;
  return SPMVB_OK;
}

// What the GPU streams is the engine's choice.  A layout that carries an engine-private device layout next to its API
// pieces (an irregular matrix: plan_device_params) offers two candidates - the API image with the global-gather kernel,
// the device layout with the x-window kernel.  Which one is faster depends on how the matrix's scattered accesses split
// between x (gathers) and y (updates), so the engine measures: both are uploaded, one SpMV of each is timed on the
// actual matrix (not under a profiler: the timings are noise there), the slower one is freed.  An explicit variant or
// option autotune = 0 skips the measurement and takes the device layout.
int spmvb_engine_create(const spmvb_layout *l, int device, int variant, spmvb_engine **out) {
  const Layout *A = (const Layout *)l;
  if (!A || !out) return fail(SPMVB_E_ARG, "engine_create: NULL");
  *out = nullptr;
  const Layout *L = A->dev ? A->dev : A;
  if (!L->stream || !L->rowmap || !A->stream || !A->rowmap)
    return fail(SPMVB_E_ARG, "engine_create: this layout was built on a GPU and lives in its engine; fetch it first");
  Engine *E = nullptr;
  int rc;
  if (variant == kVariantWide) {  // explicitly the wide image
    if (!A->wide) return fail(SPMVB_E_ARG, "engine_create: variant 9 needs a layout with a wide image (option wide = 1)");
    rc = engine_create_single(A->wide, device, kVariantDefault, &E);
    if (rc) return rc;
    *out = (spmvb_engine *)E;
    return SPMVB_OK;
  }
  if (variant == kVariantEll) {  // explicitly the ELL image
    if (!A->ell) return fail(SPMVB_E_ARG, "engine_create: variant 10 needs a layout with an ELL image (a regular matrix)");
    rc = engine_create_ell(A, device, &E);
    if (rc) return rc;
    *out = (spmvb_engine *)E;
    return SPMVB_OK;
  }
  rc = engine_create_single(L, device, variant, &E);
  if (rc) return rc;
  if (A->dev && variant == kVariantDefault && options().autotune != 0) {
    Engine *E0 = nullptr;
    float t_dev = 0.f, t_api = 0.f;
    rc = engine_time_step(E, 3, &t_dev);
    if (rc == SPMVB_OK) rc = engine_create_single(A, device, variant, &E0);
    if (rc == SPMVB_OK) rc = engine_time_step(E0, t_dev < 5.f ? 3 : 1, &t_api);
    if (rc) { spmvb_engine_free((spmvb_engine *)E); spmvb_engine_free((spmvb_engine *)E0); return rc; }
    E->tune_ms[0] = E0->tune_ms[0] = t_api; E->tune_ms[1] = E0->tune_ms[1] = t_dev;
    if (t_api < t_dev) std::swap(E, E0);
    spmvb_engine_free((spmvb_engine *)E0);
  }
  if (A->wide && variant == kVariantDefault && options().autotune != 0) {  // third candidate: the wide image
    Engine *E2 = nullptr;
    float t_cur = std::min(E->tune_ms[0] > 0.f ? E->tune_ms[0] : 1e30f, E->tune_ms[1] > 0.f ? E->tune_ms[1] : 1e30f), t_wide = 0.f;
    rc = SPMVB_OK;
    if (t_cur > 1e29f) {  // no device layout: the engine so far has not been timed yet
      rc = engine_time_step(E, 3, &t_cur);
      E->tune_ms[0] = t_cur;
    }
    if (rc == SPMVB_OK) rc = engine_create_single(A->wide, device, kVariantDefault, &E2);
    if (rc == SPMVB_OK) rc = engine_time_step(E2, 3, &t_wide);
    if (rc) { spmvb_engine_free((spmvb_engine *)E); spmvb_engine_free((spmvb_engine *)E2); return rc; }
    E->tune_ms[2] = t_wide;
    for (int i = 0; i < 4; i++) E2->tune_ms[i] = E->tune_ms[i];
    if (t_wide < t_cur) std::swap(E, E2);
    spmvb_engine_free((spmvb_engine *)E2);
  }
  if (A->ell && variant == kVariantDefault && options().autotune != 0) {  // regular matrix: the ELL image against the best so far
    Engine *E3 = nullptr;
    float t_cur = 1e30f, t_ell = 0.f;
    for (int i = 0; i < 3; i++)
      if (E->tune_ms[i] > 0.f) t_cur = std::min(t_cur, E->tune_ms[i]);
    rc = SPMVB_OK;
    if (t_cur > 1e29f) {
      rc = engine_time_step(E, 5, &t_cur);
      E->tune_ms[E->wide ? 2 : 0] = t_cur;
    }
    if (rc == SPMVB_OK) rc = engine_create_ell(A, device, &E3);
    if (rc == SPMVB_OK) rc = engine_time_step(E3, 5, &t_ell);
    if (rc) { spmvb_engine_free((spmvb_engine *)E); spmvb_engine_free((spmvb_engine *)E3); return rc; }
    E->tune_ms[3] = t_ell;
    for (int i = 0; i < 4; i++) E3->tune_ms[i] = E->tune_ms[i];
    if (t_ell < t_cur || options().ell == 1) std::swap(E, E3);
    spmvb_engine_free((spmvb_engine *)E3);
  }
  *out = (spmvb_engine *)E;
  return SPMVB_OK;
}

// create_csr_hw_matrix on the GPU: the CSR goes to the device (unless it is there already), the layout is built there
// (layout_gpu_steps.h) straight into the engine's image.  *layout_out gets every host-side table (csr_hw_matrix
// fields, chunk metadata, rows to clear); its two big arrays - the pieces and the row map - stay on the device until
// spmvb_engine_fetch_layout asks for them.  When the engine wants to stream the matrix under other parameters than the
// API layout's (plan_device_params), the builder runs a second time with those: the first image then only serves
// spmvb_engine_fetch_layout.
int spmvb_engine_create_from_csr(uint32_t rows, uint32_t cols, const uint64_t *row_ptr, const uint32_t *col_ind,
                                 const void *values, int n_cu, int vf, int is_double, uint32_t cols_div_blocks,
                                 int device, int variant, int csr_on_device, spmvb_layout **layout_out,
                                 spmvb_engine **engine_out) {
  if (!layout_out || !engine_out || !row_ptr) return fail(SPMVB_E_ARG, "engine_create_from_csr: NULL");
  *layout_out = nullptr; *engine_out = nullptr;
  if (rows == 0 || cols == 0) return fail(SPMVB_E_ARG, "empty matrix");
  const double t0 = omp_get_wtime();
  Engine *E = nullptr;
  int rc = engine_open(device, variant, &E);
  if (rc) return rc;
  uint64_t *d_rp = nullptr; uint32_t *d_ci = nullptr; void *d_va = nullptr, *d_csr = nullptr;
  Layout *L = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  auto run = [&]() -> int {
    const int vb = is_double ? 8 : 4;
    uint64_t nnz = 0;
    for (auto &x : ev) CUDA_TRY(cudaEventCreate(&x));
    CUDA_TRY(cudaEventRecord(ev[0], E->stream));
    if (csr_on_device) {
      CUDA_TRY(cudaMemcpyAsync(&nnz, row_ptr + rows, 8, cudaMemcpyDeviceToHost, E->stream));
      CUDA_TRY(cudaStreamSynchronize(E->stream));
      if (nnz && (!col_ind || !values)) return fail(SPMVB_E_ARG, "col_ind/values are NULL");
    } else {
      nnz = row_ptr[rows];
      if (nnz && (!col_ind || !values)) return fail(SPMVB_E_ARG, "col_ind/values are NULL");
      // one allocation for the three arrays (values first: 8-byte aligned)
      const size_t va_bytes = (std::max<size_t>(nnz, 1) * vb + 255) & ~(size_t)255;
      const size_t rp_bytes = (((size_t)rows + 1) * 8 + 255) & ~(size_t)255;
      CUDA_TRY(cudaMalloc(&d_csr, va_bytes + rp_bytes + std::max<size_t>(nnz, 1) * 4));
      d_va = d_csr;
      d_rp = (uint64_t *)((uint8_t *)d_csr + va_bytes);
      d_ci = (uint32_t *)((uint8_t *)d_csr + va_bytes + rp_bytes);
      CUDA_TRY(cudaMemcpyAsync(d_rp, row_ptr, ((size_t)rows + 1) * 8, cudaMemcpyHostToDevice, E->stream));
      CUDA_TRY(cudaMemcpyAsync(d_ci, col_ind, (size_t)nnz * 4, cudaMemcpyHostToDevice, E->stream));
      CUDA_TRY(cudaMemcpyAsync(d_va, values, (size_t)nnz * vb, cudaMemcpyHostToDevice, E->stream));
    }
    CUDA_TRY(cudaEventRecord(ev[1], E->stream));
    const uint64_t *rp = csr_on_device ? row_ptr : d_rp;
    const uint32_t *ci = csr_on_device ? col_ind : d_ci;
    const void *va = csr_on_device ? values : d_va;
    const Layout *D = nullptr;
    {
      CudaBackend be_api, be_dev;
      be_api.st = be_dev.st = E->stream; be_api.sms = be_dev.sms = E->sms;
      LbImage api, dev;
      int r = lb_build_pair(be_api, be_dev, rows, cols, nnz, rp, ci, va, n_cu, vf, is_double, cols_div_blocks, &L, &api, &dev);
      if (r) return r;
      if (L->dev) {
        // the API image keeps the pieces and the row map for fetch_layout; its rows-to-clear list is of no use
        E->d_api_stream = api.image; E->d_api_rowmap = api.rowmap;
        cudaFree(api.zero_rows);
        E->d_stream = dev.image; E->d_rowmap = dev.rowmap; E->d_zero_rows = dev.zero_rows;
        D = L->dev;
      } else {
        E->d_stream = api.image; E->d_rowmap = api.rowmap; E->d_zero_rows = api.zero_rows;
        D = L;
      }
      be_dev.trace("(end of build)");
    }
    CUDA_TRY(cudaEventRecord(ev[2], E->stream));
    int r = engine_adopt_layout(E, D);
    if (r) return r;
    r = engine_finish(E, D);
    if (r) return r;
    if (variant == kVariantDefault && options().autotune != 0) {  // regular matrix: the sliced-ELLPACK image as well
      r = engine_try_ell_gpu(E, L, rows, nnz, rp, ci, va);
      if (r) return r;
    }
    CUDA_TRY(cudaEventElapsedTime(&E->build_ms[0], ev[0], ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&E->build_ms[1], ev[1], ev[2]));
    return SPMVB_OK;
  };
  rc = run();
  if (d_csr) cudaFree(d_csr);
  for (auto &x : ev) if (x) cudaEventDestroy(x);
  if (rc) { delete L; spmvb_engine_free((spmvb_engine *)E); return rc; }
  cudaStreamSynchronize(E->stream);
  E->build_ms[2] = (float)((omp_get_wtime() - t0) * 1e3);
  *layout_out = (spmvb_layout *)L;
  *engine_out = (spmvb_engine *)E;
  return SPMVB_OK;
}

int spmvb_engine_fetch_layout(spmvb_engine *e, spmvb_layout *l) {
  Engine *E = (Engine *)e;
  Layout *L = (Layout *)l;
  if (!E || !L) return fail(SPMVB_E_ARG, "fetch_layout: NULL");
  const Layout *D = L->dev ? L->dev : L;
  if (D->n_chunks != E->n_chunks || D->n_pairs != E->n_pairs || D->stream_bytes != E->stream_bytes || L->rows != E->rows)
    return fail(SPMVB_E_ARG, "fetch_layout: this layout does not belong to this engine");
  if (L->stream && L->rowmap) return SPMVB_OK;
  if (L->dev && !E->d_api_stream) return fail(SPMVB_E_ARG, "fetch_layout: the API image was already released");
  CUDA_TRY(cudaSetDevice(E->device));
  CudaBackend be;
  be.st = E->stream; be.sms = E->sms;
  LbImage img;
  img.image = L->dev ? E->d_api_stream : E->d_stream;
  img.rowmap = L->dev ? E->d_api_rowmap : E->d_rowmap;
  int rc = lb_fetch_host(be, L, img);
  if (rc == SPMVB_OK && L->dev) {  // the pieces are on the host now: the device copy of the API image can go
    cudaFree(E->d_api_stream); cudaFree(E->d_api_rowmap);
    E->d_api_stream = nullptr; E->d_api_rowmap = nullptr;
  }
  return rc;
}

int spmvb_engine_build_ms(const spmvb_engine *e, float *out3) {
  const Engine *E = (const Engine *)e;
  if (!E || !out3) return fail(SPMVB_E_ARG, "build_ms");
  for (int i = 0; i < 3; i++) out3[i] = E->build_ms[i];
  return SPMVB_OK;
}

void spmvb_engine_free(spmvb_engine *e) {
  Engine *E = (Engine *)e;
  if (!E) return;
  cudaSetDevice(E->device);
  if (E->stream) cudaStreamSynchronize(E->stream);
  cudaFree(E->d_stream); cudaFree(E->d_rowmap); cudaFree(E->d_zero_rows); cudaFree(E->d_items); cudaFree(E->d_cta_first);
  cudaFree(E->d_x); cudaFree(E->d_y); cudaFree(E->d_scalar); cudaFree(E->d_flush);
  cudaFree(E->d_items_t); cudaFree(E->d_cta_first_t); cudaFree(E->d_ell);
  if (E->down_stream) cudaStreamDestroy(E->down_stream);
  for (auto &x : E->ev_tile) if (x) cudaEventDestroy(x);
  for (auto &x : E->ev_block) if (x) cudaEventDestroy(x);
  if (E->copy_stream) cudaStreamDestroy(E->copy_stream);
  cudaFree(E->d_api_stream); cudaFree(E->d_api_rowmap); cudaFree(E->d_cg);
  if (E->h_stage) cudaFreeHost(E->h_stage);
  if (E->h_scalar) cudaFreeHost(E->h_scalar);
  for (auto &x : E->ev_piece) if (x) cudaEventDestroy(x);
  if (E->ev_t0) cudaEventDestroy(E->ev_t0);
  if (E->ev_t1) cudaEventDestroy(E->ev_t1);
  for (auto &x : E->ev) cudaEventDestroy(x);
  if (E->stream) cudaStreamDestroy(E->stream);
  delete E;
}

int spmvb_engine_set_variant(spmvb_engine *e, int variant) {
  if (!e || variant < 0 || variant > 10) return fail(SPMVB_E_ARG, "variant");
  const Engine *E0 = (const Engine *)e;
  if (variant == kVariantEll && !E0->ell) return fail(SPMVB_E_ARG, "variant 10: this engine holds no ELL image");
  if (variant != 0 && variant != kVariantEll && (!E0->d_stream || E0->wide != (variant == kVariantWide)))
    return fail(SPMVB_E_ARG, "variant: this engine holds no image that kernel can walk (the wide image has its own kernel, 9; "
                             "a host-built ELL engine holds the ELL image only)");
  ((Engine *)e)->variant = variant;
  return SPMVB_OK;
}
int spmvb_engine_variant(const spmvb_engine *e) {
  const Engine *E = (const Engine *)e;
  return E->variant == kVariantDefault ? E->auto_variant : E->variant;
}
uint64_t spmvb_engine_launches(const spmvb_engine *e) { return ((const Engine *)e)->launches; }
uint64_t spmvb_engine_algorithmic_bytes(const spmvb_engine *e) {
  const Engine *E = (const Engine *)e;
  // x is credited once, and only the column blocks this shard touches (a row shard of a banded matrix reads a
  // band of x, not all of it): nnz*(2+vb) + rows*vb + x_touched*vb
  return E->real_nnz * (2 + (uint64_t)E->vb) + (uint64_t)E->rows * E->vb + E->x_touched * E->vb;
}
uint64_t spmvb_engine_x_upload_bytes(const spmvb_engine *e) {
  const Engine *E = (const Engine *)e;
  uint64_t cols = 0;
  for (size_t i = 0; i + 1 < E->x_ranges.size(); i += 2) {
    const uint64_t end = std::min<uint64_t>(E->x_ranges[i + 1], E->expanded_cols);
    if (end > E->x_ranges[i]) cols += end - E->x_ranges[i];
  }
  return cols * E->vb;
}
void *spmvb_engine_x_dev(spmvb_engine *e) { return ((Engine *)e)->d_x; }
uint64_t spmvb_engine_x_len(const spmvb_engine *e) { return ((const Engine *)e)->x_len; }
void *spmvb_engine_y_dev(spmvb_engine *e) { return ((Engine *)e)->d_y; }
void *spmvb_engine_stream(spmvb_engine *e) { return (void *)((Engine *)e)->stream; }

int spmvb_engine_set_x(spmvb_engine *e, const void *x_host, uint32_t n) {
  Engine *E = (Engine *)e;
  if (!E || !x_host) return fail(SPMVB_E_ARG, "set_x");
  CUDA_TRY(cudaSetDevice(E->device));
  const uint64_t m = std::min<uint32_t>(n, E->expanded_cols);
  for (size_t i = 0; i + 1 < E->x_ranges.size(); i += 2) {  // only what the matrix can read
    const uint64_t first = E->x_ranges[i], end = std::min<uint64_t>(E->x_ranges[i + 1], m);
    if (first >= end) continue;
    CUDA_TRY(cudaMemcpyAsync((uint8_t *)E->d_x + first * E->vb, (const uint8_t *)x_host + first * E->vb,
                             (size_t)(end - first) * E->vb, cudaMemcpyHostToDevice, E->stream));
  }
  if (m < E->x_len)  // zero-pad the remaining columns (csr_hw.cpp:1478-1481)
    CUDA_TRY(cudaMemsetAsync((uint8_t *)E->d_x + (size_t)m * E->vb, 0, (E->x_len - m) * E->vb, E->stream));
  return SPMVB_OK;
}

int spmvb_engine_spmv_dev(spmvb_engine *e, const void *x_dev, void *y_dev, int accumulate, void *stream) {
  Engine *E = (Engine *)e;
  if (!E) return fail(SPMVB_E_ARG, "spmv_dev");
  CUDA_TRY(cudaSetDevice(E->device));
  return do_spmv(E, x_dev, y_dev, accumulate, stream ? (cudaStream_t)stream : E->stream);
}

int spmvb_engine_sync(spmvb_engine *e) {
  Engine *E = (Engine *)e;
  if (!E) return fail(SPMVB_E_ARG, "sync");
  CUDA_TRY(cudaSetDevice(E->device));
  CUDA_TRY(cudaStreamSynchronize(E->stream));
  return SPMVB_OK;
}

int spmvb_engine_get_y(spmvb_engine *e, void *y_host, uint32_t n, int accumulate) {
  Engine *E = (Engine *)e;
  if (!E || !y_host) return fail(SPMVB_E_ARG, "get_y");
  CUDA_TRY(cudaSetDevice(E->device));
  const uint32_t m = std::min<uint32_t>(n, E->rows);
  const size_t bytes = (size_t)m * E->vb;
  if (!accumulate) {
    CUDA_TRY(cudaMemcpyAsync(y_host, E->d_y, bytes, cudaMemcpyDeviceToHost, E->stream));
    CUDA_TRY(cudaStreamSynchronize(E->stream));
    return SPMVB_OK;
  }
  if (E->h_stage_bytes < bytes) {
    if (E->h_stage) cudaFreeHost(E->h_stage);
    E->h_stage = nullptr; E->h_stage_bytes = 0;
    CUDA_TRY(cudaMallocHost(&E->h_stage, bytes));
    E->h_stage_bytes = bytes;
  }
  // y_fpga[row] += partial, csr_hw.cpp:1557 (the per-block partials were already summed on the device).  The copy is
  // cut into pieces so that the host addition of piece i overlaps the transfer of piece i+1.
  constexpr int kPieces = Engine::kPieces;
  uint32_t cutp[kPieces + 1];
  for (int p = 0; p <= kPieces; p++) cutp[p] = (uint32_t)((uint64_t)m * p / kPieces);
  cudaError_t werr = cudaSuccess;
  int queued = 0;  // pieces whose copy + event made it into the stream (on an error the rest is not waited for)
  for (int p = 0; p < kPieces && werr == cudaSuccess; p++) {
    const size_t o = (size_t)cutp[p] * E->vb, len = (size_t)(cutp[p + 1] - cutp[p]) * E->vb;
    if (len) werr = cudaMemcpyAsync((uint8_t *)E->h_stage + o, (const uint8_t *)E->d_y + o, len, cudaMemcpyDeviceToHost, E->stream);
    if (werr == cudaSuccess) werr = cudaEventRecord(E->ev_piece[p], E->stream);
    if (werr == cudaSuccess) queued++;
  }
  for (int p = 0; p < queued; p++) {
    cudaError_t r = cudaEventSynchronize(E->ev_piece[p]);
    if (r != cudaSuccess) werr = r;
    if (werr != cudaSuccess) continue;
    host_accumulate(E, y_host, cutp[p], cutp[p + 1]);
  }
  if (werr != cudaSuccess) return fail(SPMVB_E_CUDA, std::string("get_y: ") + cudaGetErrorString(werr));
  return SPMVB_OK;
}

int spmvb_engine_spmv_host(spmvb_engine *e, const void *x_host, uint32_t n, void *y_host, int accumulate) {
  Engine *E = (Engine *)e;
  if (!E) return fail(SPMVB_E_ARG, "spmv_host");
  if (ell_on(E) && x_host && y_host && options().e2e_tiles != 0 && E->ell_tile_slice.size() > 2) {
    CUDA_TRY(cudaSetDevice(E->device));
    return E->is_double ? spmv_host_ell<double>(E, x_host, n, y_host, accumulate) : spmv_host_ell<float>(E, x_host, n, y_host, accumulate);
  }
  if (E->wide && E->blocks > 1 && x_host && options().e2e_tiles != 0) {
    CUDA_TRY(cudaSetDevice(E->device));
    int rcw = E->is_double ? spmv_host_wide<double>(E, x_host, n) : spmv_host_wide<float>(E, x_host, n);
    if (rcw) return rcw;
    return spmvb_engine_get_y(e, y_host, E->rows, accumulate);
  }
  int rc = spmvb_engine_set_x(e, x_host, n);
  if (rc) return rc;
  return spmvb_engine_spmv_host_x_resident(e, y_host, accumulate);
}

// the second half of spmv_host for callers that brought x to the device themselves (the group replicates it over NVLink):
// kernel(s) + y down, with the row-tile pipeline where the layout has one
int spmvb_engine_spmv_host_x_resident(spmvb_engine *e, void *y_host, int accumulate) {
  Engine *E = (Engine *)e;
  if (!E || !y_host) return fail(SPMVB_E_ARG, "spmv_host_x_resident");
  CUDA_TRY(cudaSetDevice(E->device));
  const int variant = E->variant == kVariantDefault ? E->auto_variant : E->variant;
  if (variant == kVariantXs && E->n_tiles > 1 && options().e2e_tiles != 0 && !ell_on(E))
    return E->is_double ? spmv_host_tiled<double>(E, y_host, accumulate) : spmv_host_tiled<float>(E, y_host, accumulate);
  int rc = do_spmv(E, nullptr, nullptr, 0, E->stream);
  if (rc) return rc;
  return spmvb_engine_get_y(e, y_host, E->rows, accumulate);
}

int spmvb_engine_time_spmv(spmvb_engine *e, int iters, int flush_l2, float *ms_out) {
  Engine *E = (Engine *)e;
  if (!E || iters < 1 || !ms_out) return fail(SPMVB_E_ARG, "time_spmv");
  CUDA_TRY(cudaSetDevice(E->device));
  if (flush_l2 && !E->d_flush) {
    E->flush_words = (size_t)256 * 1024 * 1024 / 16;  // 256 MiB > 126 MB L2
    CUDA_TRY(cudaMalloc((void **)&E->d_flush, E->flush_words * 16));
  }
  EventList ev(2 * (size_t)iters);
  for (auto &x : ev.ev) CUDA_TRY(cudaEventCreate(&x));
  int rc = SPMVB_OK;
  for (int i = 0; i < iters && rc == SPMVB_OK; i++) {
    if (flush_l2) l2_flush_kernel<<<E->sms * 4, 256, 0, E->stream>>>(E->d_flush, E->flush_words);
    cudaEventRecord(ev[2 * i], E->stream);
    rc = do_spmv(E, nullptr, nullptr, 0, E->stream);
    cudaEventRecord(ev[2 * i + 1], E->stream);
  }
  cudaError_t ce = cudaStreamSynchronize(E->stream);
  for (int i = 0; i < iters; i++) {
    ms_out[i] = 0.f;
    if (ce == cudaSuccess && rc == SPMVB_OK) cudaEventElapsedTime(&ms_out[i], ev[2 * i], ev[2 * i + 1]);
  }
  if (rc) return rc;
  if (ce != cudaSuccess) return fail(SPMVB_E_CUDA, std::string("time_spmv: ") + cudaGetErrorString(ce));
  return SPMVB_OK;
}

int spmvb_engine_enqueue_steps(spmvb_engine *e, int steps, int flags) {
  const int flush_l2 = flags & 1, inner_events = !(flags & 2);
  Engine *E = (Engine *)e;
  if (!E || steps < 1) return fail(SPMVB_E_ARG, "enqueue_steps");
  CUDA_TRY(cudaSetDevice(E->device));
  if (flush_l2 && !E->d_flush) {
    E->flush_words = (size_t)256 * 1024 * 1024 / 16;
    CUDA_TRY(cudaMalloc((void **)&E->d_flush, E->flush_words * 16));
  }
  for (auto &x : E->ev) cudaEventDestroy(x);
  E->ev.assign(2 * (size_t)steps + 2, nullptr);
  for (auto &x : E->ev) CUDA_TRY(cudaEventCreate(&x));
  E->ev_steps = inner_events ? steps : 0;
  const void *x = E->d_x;
  void *y = E->d_y;
  CUDA_TRY(cudaEventRecord(E->ev[0], E->stream));
  for (int i = 0; i < steps; i++) {
    if (flush_l2) l2_flush_kernel<<<E->sms * 4, 256, 0, E->stream>>>(E->d_flush, E->flush_words);
    int rc = zero_y(E, y, E->stream);
    if (rc) return rc;
    if (inner_events) CUDA_TRY(cudaEventRecord(E->ev[2 + 2 * i], E->stream));
    rc = E->is_double ? launch_spmv<double>(E, (const double *)x, (double *)y, E->stream, 0)
                      : launch_spmv<float>(E, (const float *)x, (float *)y, E->stream, 0);
    if (rc) return rc;
    if (inner_events) CUDA_TRY(cudaEventRecord(E->ev[3 + 2 * i], E->stream));
  }
  CUDA_TRY(cudaEventRecord(E->ev[1], E->stream));
  return SPMVB_OK;
}

int spmvb_engine_steps_done(spmvb_engine *e) {
  Engine *E = (Engine *)e;
  if (!E || E->ev.empty()) return 1;
  return cudaEventQuery(E->ev[1]) == cudaSuccess ? 1 : 0;
}

int spmvb_engine_collect_steps(spmvb_engine *e, float *total_ms, float *kernel_ms) {
  Engine *E = (Engine *)e;
  if (!E || E->ev.empty()) return fail(SPMVB_E_ARG, "collect_steps: nothing enqueued");
  CUDA_TRY(cudaSetDevice(E->device));
  CUDA_TRY(cudaEventSynchronize(E->ev[1]));
  if (total_ms) CUDA_TRY(cudaEventElapsedTime(total_ms, E->ev[0], E->ev[1]));
  if (kernel_ms)
    for (int i = 0; i < E->ev_steps; i++) CUDA_TRY(cudaEventElapsedTime(&kernel_ms[i], E->ev[2 + 2 * i], E->ev[3 + 2 * i]));
  for (auto &x : E->ev) cudaEventDestroy(x);
  E->ev.clear();
  return SPMVB_OK;
}

int spmvb_engine_scale_copy(spmvb_engine *e, const void *src_dev, void *dst_dev, uint32_t n, double scale,
                            void *stream) {
  Engine *E = (Engine *)e;
  if (!E || !src_dev || !dst_dev) return fail(SPMVB_E_ARG, "scale_copy");
  CUDA_TRY(cudaSetDevice(E->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : E->stream;
  if (E->is_double)
    scale_copy_kernel<double><<<E->sms * 4, 256, 0, st>>>((const double *)src_dev, (double *)dst_dev, n, scale);
  else
    scale_copy_kernel<float><<<E->sms * 4, 256, 0, st>>>((const float *)src_dev, (float *)dst_dev, n, scale);
  E->launches++;
  CUDA_TRY(cudaGetLastError());
  return SPMVB_OK;
}

int spmvb_engine_scale_rsqrt(spmvb_engine *e, const void *src_dev, void *dst_dev, uint32_t n, const double *sumsq_dev,
                             void *stream) {
  Engine *E = (Engine *)e;
  if (!E || !src_dev || !dst_dev || !sumsq_dev) return fail(SPMVB_E_ARG, "scale_rsqrt");
  CUDA_TRY(cudaSetDevice(E->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : E->stream;
  if (E->is_double)
    scale_rsqrt_kernel<double><<<E->sms * 4, 256, 0, st>>>((const double *)src_dev, (double *)dst_dev, n, sumsq_dev);
  else
    scale_rsqrt_kernel<float><<<E->sms * 4, 256, 0, st>>>((const float *)src_dev, (float *)dst_dev, n, sumsq_dev);
  E->launches++;
  CUDA_TRY(cudaGetLastError());
  return SPMVB_OK;
}

int spmvb_engine_sumsq(spmvb_engine *e, const void *src_dev, uint32_t n, double *out_dev, void *stream) {
  Engine *E = (Engine *)e;
  if (!E || !src_dev || !out_dev) return fail(SPMVB_E_ARG, "sumsq");
  CUDA_TRY(cudaSetDevice(E->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : E->stream;
  CUDA_TRY(cudaMemsetAsync(out_dev, 0, sizeof(double), st));
  if (E->is_double)
    sumsq_kernel<double><<<E->sms * 2, 256, 0, st>>>((const double *)src_dev, n, out_dev);
  else
    sumsq_kernel<float><<<E->sms * 2, 256, 0, st>>>((const float *)src_dev, n, out_dev);
  E->launches++;
  CUDA_TRY(cudaGetLastError());
  return SPMVB_OK;
}

int spmvb_engine_power_iter(spmvb_engine *e, int iters, double *norm_out) {
  Engine *E = (Engine *)e;
  if (!E || iters < 1) return fail(SPMVB_E_ARG, "power_iter");
  if (E->rows != E->cols) return fail(SPMVB_E_ARG, "power_iter needs a square matrix");
  CUDA_TRY(cudaSetDevice(E->device));
  // everything stays on the device: the norm's square is consumed by the scale kernel where it was summed, and only
  // the last one comes back
  CUDA_TRY(cudaEventRecord(E->ev_t0, E->stream));
  for (int it = 0; it < iters; it++) {
    int rc = do_spmv(E, nullptr, nullptr, 0, E->stream);
    if (rc) return rc;
    rc = spmvb_engine_sumsq(e, E->d_y, E->rows, E->d_scalar, E->stream);
    if (rc) return rc;
    rc = spmvb_engine_scale_rsqrt(e, E->d_y, E->d_x, E->rows, E->d_scalar, E->stream);
    if (rc) return rc;
  }
  CUDA_TRY(cudaEventRecord(E->ev_t1, E->stream));
  CUDA_TRY(cudaMemcpyAsync(E->h_scalar, E->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CUDA_TRY(cudaStreamSynchronize(E->stream));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, E->ev_t0, E->ev_t1) == cudaSuccess) E->last_iter_ms = ms / (float)iters;
  if (norm_out) *norm_out = std::sqrt(E->h_scalar[0]);
  return SPMVB_OK;
}

// bounds-checked build only: the violation counters of the kernels {chunk index, row-map index, y row, x index, x window
// offset} since the library was loaded; -1 in a release build
int spmvb_debug_bounds_errors(uint64_t *out5) {
#ifdef SPMVB_CHECK_BOUNDS
  unsigned long long h[5] = {0, 0, 0, 0, 0};
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(h, g_bounds_errors, sizeof h));
  for (int i = 0; i < 5; i++) out5[i] = h[i];
  return 1;
#else
  (void)out5;
  return -1;
#endif
}

// for tests: the ELL image the engine streams (device -> host); returns its size in bytes, 0 when the engine streams
// something else
int64_t spmvb_engine_ell_image(spmvb_engine *e, void *out, uint64_t max_bytes) {
  Engine *E = (Engine *)e;
  if (!E) return fail(SPMVB_E_ARG, "ell_image");
  if (!E->ell) return 0;
  const uint64_t bytes = (uint64_t)E->ell_slices * E->ell_slice_bytes;
  if (out && max_bytes >= bytes) {
    CUDA_TRY(cudaSetDevice(E->device));
    CUDA_TRY(cudaMemcpyAsync(out, E->d_ell, bytes, cudaMemcpyDeviceToHost, E->stream));
    CUDA_TRY(cudaStreamSynchronize(E->stream));
  }
  return (int64_t)bytes;
}

float spmvb_engine_last_iter_ms(const spmvb_engine *e) { return e ? ((const Engine *)e)->last_iter_ms : 0.f; }

int spmvb_engine_device_layout(const spmvb_engine *e, uint64_t *out) {
  const Engine *E = (const Engine *)e;
  if (!E || !out) return fail(SPMVB_E_ARG, "device_layout");
  out[0] = (uint64_t)E->dev_cu; out[1] = (uint64_t)E->dev_vf; out[2] = E->cdb; out[3] = E->cu_major ? 1u : 0u;
  out[4] = E->n_pairs; out[5] = E->n_chunks; out[6] = E->zero_all ? UINT64_MAX : (uint64_t)E->n_zero_rows;
  out[7] = E->stream_bytes;
  if (ell_on(E)) {  // no pairs, no chunks, no rows to clear: slices
    out[0] = 1; out[1] = 1; out[4] = 0; out[5] = E->ell_slices; out[6] = 0; out[7] = (uint64_t)E->ell_slices * E->ell_slice_bytes;
  } out[8] = (uint64_t)E->n_tiles; out[9] = E->tall ? 1u : 0u; out[10] = (uint64_t)E->xs_cfg;
  out[11] = (uint64_t)(E->tune_ms[0] * 1000.f); out[12] = (uint64_t)(E->tune_ms[1] * 1000.f);
  out[13] = E->wide ? 1u : 0u; out[14] = (uint64_t)(E->tune_ms[2] * 1000.f); out[15] = (uint64_t)E->blocks;
  out[16] = ell_on(E) ? 1u : 0u; out[17] = (uint64_t)(E->tune_ms[3] * 1000.f); out[18] = E->ell_width;
  out[19] = (uint64_t)(E->ell_tile_slice.empty() ? 0 : E->ell_tile_slice.size() - 1);
  return SPMVB_OK;
}

int spmvb_engine_cg(spmvb_engine *e, const void *b_host, void *x_host, int max_iters, double rel_tol, int *iters_out,
                    double *relres_out) {
  Engine *E = (Engine *)e;
  if (!E || !b_host || !x_host || max_iters < 1 || !(rel_tol >= 0.0)) return fail(SPMVB_E_ARG, "cg");
  if (E->rows != E->cols) return fail(SPMVB_E_ARG, "cg needs a square matrix");
  CUDA_TRY(cudaSetDevice(E->device));
  if (E->is_double) return cg_impl<double>(E, (const double *)b_host, (double *)x_host, max_iters, rel_tol, iters_out, relres_out);
  return cg_impl<float>(E, (const float *)b_host, (float *)x_host, max_iters, rel_tol, iters_out, relres_out);
}

}  // extern "C"
