// sm_100a SpMV kernels over the hw_matrix stream (the reference's HLS dataflow pipeline, src/spmv.cpp:6-205,
// fused with the host accumulation accum_results, src/csr_hw.cpp:1531-1565).
//
// Mapping of the reference stages:
//   read_data_submatrix  (spmv.cpp:6-34)    -> chunk fetch: ld.global.v4 per lane (variant DIRECT) or a per-warp
//                                              cp.async.bulk (TMA 1-D) ring with mbarriers (variant RING / XSMEM)
//   stream_data_col_ind  (spmv.cpp:36-49)   -> in-register unpack of 8 x (15-bit column | end-of-row bit)
//   stream_data_values   (spmv.cpp:51-64)   -> in-register reinterpretation of the value words
//   compute_results      (spmv.cpp:66-104)  -> per-lane left-to-right multiply/add over its 8 entries (mul and add
//                                              separately rounded, like the HLS cores) + warp segmented scan keyed on
//                                              the end-of-row bit
//   write_back_results + accum_results      -> red.global.add to y[rowmap[rank]] (rank = running count of
//                                              end-of-row bits), no partial-y buffers, no host pass
//   x slice copy L0      (spmv.cpp:182-192) -> variant XSMEM: cp.async.bulk of the block's x slice into shared
//                                              memory; other variants gather x through L1/L2
//
// Work unit: a "chunk" = 32 consecutive 8-entry groups of one piece (one group per lane).  Every warp owns a
// contiguous range of chunks, carries the open row sum in registers from chunk to chunk, and only the two ends of its
// range depend on atomics for correctness across warps.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "layout.h"

namespace spmvb {

template <typename VT> struct VTraits;
template <> struct VTraits<double> {
  static constexpr int kValWords = 4;    // 16-byte value words per group
  static constexpr int kGroupWords = 5;  // RATIO_col_val, src/util.h:67
};
template <> struct VTraits<float> {
  static constexpr int kValWords = 2;
  static constexpr int kGroupWords = 3;
};

__device__ __forceinline__ double vmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float vmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double vadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float vadd(float a, float b) { return __fadd_rn(a, b); }

template <typename VT> __device__ __forceinline__ VT value_of(const uint4 *vw, int s);
template <> __device__ __forceinline__ double value_of<double>(const uint4 *vw, int s) {
  const uint4 w = vw[s >> 1];
  return (s & 1) ? __hiloint2double((int)w.w, (int)w.z) : __hiloint2double((int)w.y, (int)w.x);
}
template <> __device__ __forceinline__ float value_of<float>(const uint4 *vw, int s) {
  const uint4 w = vw[s >> 2];
  const uint32_t u = (s & 3) == 0 ? w.x : (s & 3) == 1 ? w.y : (s & 3) == 2 ? w.z : w.w;
  return __uint_as_float(u);
}

__device__ __forceinline__ uint32_t idx16(const uint4 &iw, int s) {
  const uint32_t u = (s >> 1) == 0 ? iw.x : (s >> 1) == 1 ? iw.y : (s >> 1) == 2 ? iw.z : iw.w;
  return (s & 1) ? (u >> 16) : (u & 0xFFFFu);
}

// red.global.add (result unused -> RED, no return trip)
__device__ __forceinline__ void y_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void y_add(float *p, float v) { atomicAdd(p, v); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// 1-D bulk async copy global -> shared (TMA engine; SASS UBLKCP), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// One chunk for one warp: lane `lane` owns group `lane`.  iw / vw already hold the lane's index word and value
// words.  XS: x slice addressable at xs[col] (global pointer to the block's slice, or shared memory).
// `carry` (warp-uniform) is the open row sum entering the chunk; `open` says whether entries after the last
// end-of-row bit exist (so a flush at the end of the warp's range is due).  Returns through references.
template <typename VT, typename XS>
__device__ __forceinline__ void process_chunk(const uint4 &iw, const uint4 *vw, const ChunkMeta &m, XS xs,
                                              const uint32_t *__restrict__ rowmap, VT *__restrict__ y, int lane,
                                              VT &carry, bool &open, uint32_t &next_rank, bool &first_head_pending) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t valid = m.valid & 0x3FFu;
  const bool consecutive = (m.valid & kChunkRowsConsecutive) != 0;
  const int nvalid = min(8, max(0, (int)valid - lane * 8));

  // end-of-row bits of this lane's valid entries
  uint32_t eor = 0;
#pragma unroll
  for (int s = 0; s < 8; s++) eor |= ((idx16(iw, s) >> 15) & 1u) << s;
  eor &= (1u << nvalid) - 1u;
  const int n_eor = __popc(eor);

  // exclusive prefix of n_eor across lanes -> rank of this lane's first segment end
  int pre = n_eor;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(FULL, pre, d);
    if (lane >= d) pre += o;
  }
  const int total_eor = __shfl_sync(FULL, pre, 31);
  const uint32_t rank_t = m.rank0 + (uint32_t)(pre - n_eor);

  // gather + multiply; products of invalid slots are exactly 0
  VT prod[8];
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const uint32_t ci = idx16(iw, s) & 0x7FFFu;
    VT xv = (s < nvalid) ? xs[ci] : VT(0);
    VT v = (s < nvalid) ? value_of<VT>(vw, s) : VT(0);
    prod[s] = vmul(v, xv);
  }

  // per-lane sequential segmented sum (compute_results order inside a lane)
  VT acc = (lane == 0) ? carry : VT(0);
  VT head = VT(0);
  bool seen = false;
  int seg = 0;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    acc = vadd(acc, prod[s]);
    if ((eor >> s) & 1u) {
      if (!seen) {
        head = acc;
        seen = true;
      } else {
        const uint32_t rk = rank_t + (uint32_t)seg;
        const uint32_t row = consecutive ? m.row_first + (rk - m.rank0) : rowmap[rk];
        y_add(&y[row], acc);
      }
      seg++;
      acc = VT(0);
    }
  }

  // warp segmented inclusive scan of the open tails; a lane with an end-of-row bit restarts the segment
  VT v = acc;
  bool f = seen;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    VT pv = __shfl_up_sync(FULL, v, d);
    int pf = __shfl_up_sync(FULL, (int)f, d);
    if (lane >= d && !f) {
      v = vadd(pv, v);
      f = pf != 0;
    }
  }
  VT cin = __shfl_up_sync(FULL, v, 1);
  if (lane == 0) cin = VT(0);  // lane 0 already absorbed `carry`

  const uint32_t seen_mask = __ballot_sync(FULL, seen);
  if (seen) {
    const VT tot = vadd(cin, head);
    const uint32_t row = consecutive ? m.row_first + (rank_t - m.rank0) : rowmap[rank_t];
    (void)first_head_pending;
    y_add(&y[row], tot);
  }
  carry = __shfl_sync(FULL, v, 31);

  // is a row still open after this chunk?
  const int trail = nvalid - (eor ? (32 - __clz(eor)) : 0);  // valid entries after the lane's last end-of-row bit
  const uint32_t trail_mask = __ballot_sync(FULL, trail > 0);
  if (trail_mask | seen_mask) {
    const int hi_trail = trail_mask ? 31 - __clz(trail_mask) : -1;
    const int hi_seen = seen_mask ? 31 - __clz(seen_mask) : -1;
    open = hi_trail >= 0 && hi_trail >= hi_seen;
  }
  next_rank = m.rank0 + (uint32_t)total_eor;
}

// ------------------------------------------------------------------------------------------------------------------
// Variant DIRECT: every lane loads its group straight from global memory (5 / 3 x ld.global.v4).
template <typename VT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
    spmv_direct_kernel(const uint4 *__restrict__ stream, const ChunkMeta *__restrict__ meta,
                       const uint32_t *__restrict__ rowmap, const VT *__restrict__ x, VT *__restrict__ y,
                       unsigned long long n_chunks, uint32_t cdb) {
  constexpr int GW = VTraits<VT>::kGroupWords;
  constexpr int VW = VTraits<VT>::kValWords;
  const int lane = threadIdx.x & 31;
  const unsigned long long w = (unsigned long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const unsigned long long W = (unsigned long long)gridDim.x * WARPS;
  const unsigned long long c0 = n_chunks * w / W, c1 = n_chunks * (w + 1) / W;
  VT carry = VT(0);
  bool open = false, fhp = true;
  uint32_t next_rank = 0;
  for (unsigned long long c = c0; c < c1; c++) {
    const uint4 mraw = __ldg(reinterpret_cast<const uint4 *>(meta) + c);
    ChunkMeta m;
    m.rank0 = mraw.x; m.block = mraw.y; m.valid = mraw.z; m.row_first = mraw.w;
    const uint4 *g = stream + (c * 32 + lane) * GW;
    uint4 iw = __ldg(g);
    uint4 vw[VW];
#pragma unroll
    for (int i = 0; i < VW; i++) vw[i] = __ldg(g + 1 + i);
    const VT *xs = x + (size_t)m.block * cdb;
    process_chunk<VT, const VT *>(iw, vw, m, xs, rowmap, y, lane, carry, open, next_rank, fhp);
  }
  if (open && lane == 0) y_add(&y[rowmap[next_rank]], carry);
}

// ------------------------------------------------------------------------------------------------------------------
// Variant RING: per-warp ring of STAGES chunks filled by cp.async.bulk (one elected lane issues, an mbarrier per
// stage counts the bytes); lanes read their group with conflict-free LDS.128 (lane stride 80 B / 48 B).
template <typename VT, int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32)
    spmv_ring_kernel(const uint4 *__restrict__ stream, const ChunkMeta *__restrict__ meta,
                     const uint32_t *__restrict__ rowmap, const VT *__restrict__ x, VT *__restrict__ y,
                     unsigned long long n_chunks, uint32_t cdb) {
  constexpr int GW = VTraits<VT>::kGroupWords;
  constexpr int VW = VTraits<VT>::kValWords;
  constexpr uint32_t CHUNK_BYTES = GW * 16 * 32;
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint8_t *ring = smem + (size_t)warp * STAGES * CHUNK_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)WARPS * STAGES * CHUNK_BYTES) + warp * STAGES;
  const unsigned long long w = (unsigned long long)blockIdx.x * WARPS + warp;
  const unsigned long long W = (unsigned long long)gridDim.x * WARPS;
  const unsigned long long c0 = n_chunks * w / W, c1 = n_chunks * (w + 1) / W;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      if (c0 + s < c1) {
        const uint32_t bar = smem_u32(&bars[s]);
        mbar_expect_tx(bar, CHUNK_BYTES);
        bulk_g2s(smem_u32(ring + (size_t)s * CHUNK_BYTES), stream + (c0 + s) * 32 * GW, CHUNK_BYTES, bar);
      }
    }
  }
  VT carry = VT(0);
  bool open = false, fhp = true;
  uint32_t next_rank = 0;
  int stage = 0;
  uint32_t parity = 0;
  uint4 mraw = (c0 < c1) ? __ldg(reinterpret_cast<const uint4 *>(meta) + c0) : make_uint4(0, 0, 0, 0);
  for (unsigned long long c = c0; c < c1; c++) {
    ChunkMeta m;
    m.rank0 = mraw.x; m.block = mraw.y; m.valid = mraw.z; m.row_first = mraw.w;
    if (c + 1 < c1) mraw = __ldg(reinterpret_cast<const uint4 *>(meta) + c + 1);
    const uint32_t bar = smem_u32(&bars[stage]);
    mbar_wait(bar, parity);
    const uint4 *g = reinterpret_cast<const uint4 *>(ring + (size_t)stage * CHUNK_BYTES) + lane * GW;
    uint4 iw = g[0];
    uint4 vw[VW];
#pragma unroll
    for (int i = 0; i < VW; i++) vw[i] = g[1 + i];
    __syncwarp();  // every lane has its group in registers: the stage may be refilled
    if (lane == 0 && c + STAGES < c1) {
      mbar_expect_tx(bar, CHUNK_BYTES);
      bulk_g2s(smem_u32(ring + (size_t)stage * CHUNK_BYTES), stream + (c + STAGES) * 32 * GW, CHUNK_BYTES, bar);
    }
    if (++stage == STAGES) { stage = 0; parity ^= 1; }
    const VT *xs = x + (size_t)m.block * cdb;
    process_chunk<VT, const VT *>(iw, vw, m, xs, rowmap, y, lane, carry, open, next_rank, fhp);
  }
  if (open && lane == 0) y_add(&y[rowmap[next_rank]], carry);
}

// zero-fill / scale helpers for the iterated caller
template <typename VT>
__global__ void scale_copy_kernel(const VT *__restrict__ src, VT *__restrict__ dst, uint32_t n, double scale) {
  const VT s = (VT)scale;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i] * s;
}

template <typename VT>
__global__ void sumsq_kernel(const VT *__restrict__ src, uint32_t n, double *__restrict__ out) {
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = (double)src[i];
    acc += v * v;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  __shared__ double part[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    acc = lane < (int)(blockDim.x >> 5) ? part[lane] : 0.0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if (lane == 0) atomicAdd(out, acc);
  }
}

__global__ void l2_flush_kernel(uint4 *__restrict__ buf, size_t n_words) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x)
    buf[i] = make_uint4((uint32_t)i, 0, 0, 0);
}

}  // namespace spmvb
