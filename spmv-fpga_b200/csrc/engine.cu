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
#include "layout.h"
#include "layout_gpu.cuh"
#include "spmv_kernels.cuh"

namespace spmvb {

enum Variant { kVariantDefault = 0, kVariantDirect = 1, kVariantOcc4 = 6, kVariantOcc3 = 7, kVariantXs = 8 };


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
  uint32_t occ_run_log2 = 3, xs_run_log2 = 1;  // run lengths of the kernels (>= the layout's zero-list granularity)
  uint32_t n_items = 0;
  double xs_windowed_frac = 0.0;  // share of the chunks whose x window fits shared memory
  bool tall = false;               // x and y both exceed the L2 cache: explicit L2 eviction policies
  bool cu_major = false;           // pieces in CU-major device order (row tiles)
  int auto_variant = kVariantOcc3; // what variant 0 resolves to (chosen from the layout at creation)
  float tune_ms[2] = {0.f, 0.f};   // autotune timings: OCC, XS
  uint32_t n_zero_rows = 0, run_log2 = 2;
  bool zero_all = true;
  void *d_x = nullptr, *d_y = nullptr;
  double *d_scalar = nullptr;
  uint4 *d_flush = nullptr;
  size_t flush_words = 0;
  void *h_stage = nullptr;  // pinned staging for get_y(accumulate)
  size_t h_stage_bytes = 0;
  cudaStream_t stream = nullptr;
  int sms = 148;
  uint64_t launches = 0;
  int grid_cache[9][2] = {};  // [variant][is_double] -> grid size
  // asynchronous step timing (bench): events of the last enqueue_steps()
  std::vector<cudaEvent_t> ev;
  int ev_steps = 0;
  // GPU-built engines (spmvb_engine_create_from_csr): milliseconds of the CSR upload, of the build kernels (CUDA
  // events) and of the whole call (host clock)
  float build_ms[3] = {0.f, 0.f, 0.f};
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
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, (KArgs)args...);
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

template <typename VT>
static int launch_xs(Engine *E, const VT *x, VT *y, cudaStream_t st, int accumulate) {
  const uint4 *stream = reinterpret_cast<const uint4 *>(E->d_stream);
  auto kern = spmv_xs_kernel<VT, kXsWarps, kXsCap>;
  const size_t smem = (size_t)kXsCap + (size_t)kXsWarps * 2 * (VTraits<VT>::kGroupWords * 16 * 32 + 16) + kXsWarps * 16 + 16;
  int &grid = E->grid_cache[kVariantXs][sizeof(VT) == 8];
  if (grid == 0) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    grid = E->sms;
  }
  if (E->n_items == 0) return SPMVB_OK;
  CUDA_TRY(launch_pdl(kern, grid, kXsWarps * 32, smem, st, stream, (const uint32_t *)E->d_rowmap, x, y,
                      (const XsItem *)E->d_items, (const uint32_t *)E->d_cta_first, E->cdb, E->xs_run_log2,
                      (accumulate ? 4u : 0u) | (E->tall ? 8u : 0u)));
  return SPMVB_OK;
}

template <typename VT>
static int launch_spmv(Engine *E, const VT *x, VT *y, cudaStream_t st, int accumulate) {
  const uint4 *stream = reinterpret_cast<const uint4 *>(E->d_stream);
  constexpr int WARPS = 8;
  int variant = E->variant == kVariantDefault ? E->auto_variant : E->variant;
  if (E->n_chunks == 0) return SPMVB_OK;
  if (E->n_chunks >= 0x7FFFFFFFull) return fail(SPMVB_E_RANGE, "too many chunks for one engine");
  int rc = SPMVB_OK;
  if (variant == kVariantDirect) {
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
  if (E->is_double) return launch_spmv<double>(E, (const double *)x, (double *)y, st, accumulate);
  return launch_spmv<float>(E, (const float *)x, (float *)y, st, accumulate);
}

// variant 0: pick between the two production kernels from the structure of the layout (deterministic).  The
// shared-memory x window pays when nearly every entry is its own (row, block) pair with a scattered column - then the
// global gathers touch one sector per entry - and it is required for tall matrices (CU-major order, x streamed once per
// row tile).  Banded and power-law matrices keep the global gathers with more resident warps.  Measured on B200:
// Laplacian 54 vs 69 us, R-MAT 0.28 vs 0.38 ms, uniform 0.55 vs 0.43 ms, 1 B-nnz uniform 27 vs 8.9 ms (OCC vs XS).
// SPMVB_AUTOTUNE=1 times both kernels on the actual matrix instead (not under a profiler: the timings are noise there).
static int autotune(Engine *E) {
  E->auto_variant = kVariantOcc3;
  if (E->n_chunks == 0 || E->xs_windowed_frac < 0.5) return SPMVB_OK;
  const double pairs_per_chunk = (double)E->n_pairs / (double)E->n_chunks;
  if (E->cu_major || pairs_per_chunk > 160.0) E->auto_variant = kVariantXs;
  if (!getenv("SPMVB_AUTOTUNE")) return SPMVB_OK;
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

}  // namespace spmvb

using namespace spmvb;

namespace {
template <typename VT>
int cg_impl(Engine *E, const VT *b_host, VT *x_host, int max_iters, double rel_tol, int *iters_out, double *relres_out) {
  const uint32_t n = E->rows;
  const size_t bytes = (size_t)n * sizeof(VT);
  VT *d_r = nullptr, *d_xs = nullptr;
  double *s = nullptr;  // s[0], s[2]: r.r (roles swap every iteration); s[1]: p.q
  cudaStream_t st = E->stream;
  const int grid = E->sms * 4;
  int rc = SPMVB_OK, it = 0;
  double bnorm2 = 0.0, rr_host = 0.0;
  auto body = [&]() -> int {
    CUDA_TRY(cudaMalloc((void **)&d_r, bytes));
    CUDA_TRY(cudaMalloc((void **)&d_xs, bytes));
    CUDA_TRY(cudaMalloc((void **)&s, 4 * sizeof(double)));
    VT *p = (VT *)E->d_x, *q = (VT *)E->d_y;
    // x0 = 0: r = b, p = b (the padding of p beyond n stays zero), rr = b.b
    CUDA_TRY(cudaMemsetAsync(d_xs, 0, bytes, st));
    CUDA_TRY(cudaMemcpyAsync(d_r, b_host, bytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(E->d_x, 0, E->x_len * sizeof(VT), st));
    CUDA_TRY(cudaMemcpyAsync(p, d_r, bytes, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemsetAsync(s, 0, 4 * sizeof(double), st));
    dot_kernel<VT><<<grid, 256, 0, st>>>(d_r, d_r, n, s + 0);
    E->launches++;
    CUDA_TRY(cudaMemcpyAsync(&bnorm2, s + 0, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    rr_host = bnorm2;
    if (bnorm2 == 0.0) return SPMVB_OK;  // b = 0: x = 0
    const double stop2 = rel_tol * rel_tol * bnorm2;
    int cur = 0;  // index of the current r.r
    constexpr int kCheckEvery = 8;
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
        CUDA_TRY(cudaMemcpyAsync(&rr_host, s + cur, sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (!(rr_host > stop2)) break;  // converged (or NaN: give up)
      }
    }
    CUDA_TRY(cudaMemcpyAsync(x_host, d_xs, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SPMVB_OK;
  };
  rc = body();
  cudaFree(d_r); cudaFree(d_xs); cudaFree(s);
  if (rc) return rc;
  if (bnorm2 == 0.0) memset(x_host, 0, bytes);
  if (iters_out) *iters_out = it;
  if (relres_out) *relres_out = bnorm2 > 0.0 ? std::sqrt(rr_host / bnorm2) : 0.0;
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
  E->cu_major = L->cu_major;
  E->tall = (uint64_t)L->rows * L->vb > ((uint64_t)48 << 20) && (uint64_t)L->cols * L->vb > ((uint64_t)48 << 20);
  if (const char *v = getenv("SPMVB_TALL")) E->tall = atoi(v) != 0;
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
  if (const char *v = getenv("SPMVB_OCC_RUN_LOG2")) E->occ_run_log2 = (uint32_t)atoi(v);
  if (const char *v = getenv("SPMVB_XS_RUN_LOG2")) E->xs_run_log2 = (uint32_t)atoi(v);
  // rows split across run boundaries are only cleared at the layout's granularity: runs must be multiples of it
  E->occ_run_log2 = std::min(8u, std::max(E->occ_run_log2, E->run_log2));
  E->xs_run_log2 = std::min(8u, std::max(E->xs_run_log2, E->run_log2));
  CUDA_TRY(cudaMalloc(&E->d_x, E->x_len * E->vb));
  CUDA_TRY(cudaMalloc(&E->d_y, (size_t)E->rows * E->vb));
  return SPMVB_OK;
}

// work plan of the XS kernel, cleared vectors, kernel choice; the image must be on the device (or on its way, on E->stream)
static int engine_finish(Engine *E, const Layout *L) {
  std::vector<XsItem> items;
  std::vector<uint32_t> cta_first;
  build_xs_items(L, E->sms, E->xs_run_log2, items, cta_first);
  E->n_items = (uint32_t)items.size();
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

int spmvb_engine_create(const spmvb_layout *l, int device, int variant, spmvb_engine **out) {
  const Layout *L = (const Layout *)l;
  if (!L || !out) return fail(SPMVB_E_ARG, "engine_create: NULL");
  *out = nullptr;
  if (!L->stream || !L->rowmap)
    return fail(SPMVB_E_ARG, "engine_create: this layout was built on a GPU and lives in its engine; fetch it first");
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
    CUDA_TRY(cudaMalloc((void **)&E->d_rowmap, (std::max<uint64_t>(L->n_pairs, 1) + 1) * 4));
    CUDA_TRY(cudaMalloc((void **)&E->d_zero_rows, std::max<size_t>(L->zero_rows.size(), 1) * 4));
    if (L->n_chunks) {
      CUDA_TRY(cudaMemcpy2DAsync(E->d_stream, slot, L->stream, L->chunk_bytes, L->chunk_bytes, L->n_chunks,
                                 cudaMemcpyHostToDevice, E->stream));
      CUDA_TRY(cudaMemcpy2DAsync(E->d_stream + L->chunk_bytes, slot, L->chunks, sizeof(ChunkMeta), sizeof(ChunkMeta),
                                 L->n_chunks, cudaMemcpyHostToDevice, E->stream));
    }
    CUDA_TRY(cudaMemcpyAsync(E->d_rowmap, L->rowmap, L->n_pairs * 4, cudaMemcpyHostToDevice, E->stream));
    CUDA_TRY(cudaMemcpyAsync(E->d_zero_rows, L->zero_rows.data(), L->zero_rows.size() * 4, cudaMemcpyHostToDevice, E->stream));
    return engine_finish(E, L);
  };
  rc = upload();
  if (rc) { spmvb_engine_free((spmvb_engine *)E); return rc; }
  *out = (spmvb_engine *)E;
  return SPMVB_OK;
}

// create_csr_hw_matrix on the GPU: the CSR goes to the device (unless it is there already), the layout is built there
// (layout_gpu_steps.h) straight into the engine's image.  *layout_out gets every host-side table (csr_hw_matrix
// fields, chunk metadata, rows to clear); its two big arrays - the pieces and the row map - stay on the device until
// spmvb_engine_fetch_layout asks for them.
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
    CudaBackend be;
    be.st = E->stream; be.sms = E->sms;
    LbImage img;
    int r = lb_build(be, rows, cols, nnz, csr_on_device ? row_ptr : d_rp, csr_on_device ? col_ind : d_ci,
                     csr_on_device ? values : d_va, n_cu, vf, is_double, cols_div_blocks, &L, &img);
    if (r) return r;
    E->d_stream = img.image; E->d_rowmap = img.rowmap; E->d_zero_rows = img.zero_rows;
    CUDA_TRY(cudaEventRecord(ev[2], E->stream));
    be.trace("(end of build)");
    r = engine_adopt_layout(E, L);
    if (r) return r;
    be.trace("x / y allocation");
    r = engine_finish(E, L);
    if (r) return r;
    be.trace("XS plan + kernel choice");
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
  if (L->n_chunks != E->n_chunks || L->n_pairs != E->n_pairs || L->stream_bytes != E->stream_bytes || L->rows != E->rows)
    return fail(SPMVB_E_ARG, "fetch_layout: this layout does not belong to this engine");
  CUDA_TRY(cudaSetDevice(E->device));
  CudaBackend be;
  be.st = E->stream; be.sms = E->sms;
  LbImage img;
  img.image = E->d_stream; img.rowmap = E->d_rowmap;
  return lb_fetch_host(be, L, img);
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
  if (E->h_stage) cudaFreeHost(E->h_stage);
  for (auto &x : E->ev) cudaEventDestroy(x);
  if (E->stream) cudaStreamDestroy(E->stream);
  delete E;
}

int spmvb_engine_set_variant(spmvb_engine *e, int variant) {
  if (!e || variant < 0 || variant > 8) return fail(SPMVB_E_ARG, "variant");
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
  constexpr int kPieces = 8;
  cudaEvent_t ev[kPieces];
  uint32_t cutp[kPieces + 1];
  for (int p = 0; p <= kPieces; p++) cutp[p] = (uint32_t)((uint64_t)m * p / kPieces);
  for (int p = 0; p < kPieces; p++) {
    CUDA_TRY(cudaEventCreateWithFlags(&ev[p], cudaEventDisableTiming));
    const size_t o = (size_t)cutp[p] * E->vb, len = (size_t)(cutp[p + 1] - cutp[p]) * E->vb;
    if (len) CUDA_TRY(cudaMemcpyAsync((uint8_t *)E->h_stage + o, (const uint8_t *)E->d_y + o, len, cudaMemcpyDeviceToHost, E->stream));
    CUDA_TRY(cudaEventRecord(ev[p], E->stream));
  }
  cudaError_t werr = cudaSuccess;
  for (int p = 0; p < kPieces; p++) {
    cudaError_t r = cudaEventSynchronize(ev[p]);
    if (r != cudaSuccess) werr = r;
    cudaEventDestroy(ev[p]);
    if (werr != cudaSuccess) continue;
    const int64_t a = cutp[p], b = cutp[p + 1];
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
  if (werr != cudaSuccess) return fail(SPMVB_E_CUDA, std::string("get_y: ") + cudaGetErrorString(werr));
  return SPMVB_OK;
}

int spmvb_engine_spmv_host(spmvb_engine *e, const void *x_host, uint32_t n, void *y_host, int accumulate) {
  Engine *E = (Engine *)e;
  if (!E) return fail(SPMVB_E_ARG, "spmv_host");
  int rc = spmvb_engine_set_x(e, x_host, n);
  if (rc) return rc;
  rc = do_spmv(E, nullptr, nullptr, 0, E->stream);
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
  std::vector<cudaEvent_t> ev(2 * (size_t)iters);
  for (auto &x : ev) CUDA_TRY(cudaEventCreate(&x));
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
  for (auto &x : ev) cudaEventDestroy(x);
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
  double nrm = 0.0;
  for (int it = 0; it < iters; it++) {
    int rc = do_spmv(E, nullptr, nullptr, 0, E->stream);
    if (rc) return rc;
    rc = spmvb_engine_sumsq(e, E->d_y, E->rows, E->d_scalar, E->stream);
    if (rc) return rc;
    double ss = 0.0;
    CUDA_TRY(cudaMemcpyAsync(&ss, E->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    CUDA_TRY(cudaStreamSynchronize(E->stream));
    nrm = std::sqrt(ss);
    rc = spmvb_engine_scale_copy(e, E->d_y, E->d_x, E->rows, nrm > 0 ? 1.0 / nrm : 0.0, E->stream);
    if (rc) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(E->stream));
  if (norm_out) *norm_out = nrm;
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
