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

#include <type_traits>

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

// predicated variant: the RED is the only instruction under the predicate (no branch, no reconvergence barrier)
__device__ __forceinline__ void y_add_if(double *p, double v, uint32_t e) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.global.add.f64 [%0], %1;\n\t}" ::"l"(p), "d"(v), "r"(e)
               : "memory");
}
__device__ __forceinline__ void y_add_if(float *p, float v, uint32_t e) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.global.add.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"(e)
               : "memory");
}

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
// bulk prefetch global -> L2 (no shared memory, no completion tracking): deepens the bytes in flight to DRAM beyond
// what the shared-memory ring can hold
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// One chunk for one warp: lane `lane` owns group `lane`.  iw holds the lane's index word, vw its value words and xv
// its 8 gathered x values (MUL) or its 8 products value * x[col] (!MUL, vw unused).  Warp-uniform state threaded through the chunks of a run:
//   carry      open row sum entering the chunk (absorbed by lane 0)
//   open       entries after the last end-of-row bit exist (a hand-over is due at the end of the run)
//   next_rank  rank of that open row
//   head_red   the run began in the middle of a row: that row's end must be an atomic even in a `sole` chunk
// A chunk with fewer than 256 real entries is the last one of its piece: only its end-of-row bits need masking (the
// padding slots follow every real entry of their lane, so they never reach a row sum that is written) and it always
// ends closed.  `consec`: the chunk's rows are row_first, row_first + 1, ... (no row-map loads).
// `sole` chunks (every row lives in one column block only) write y with plain stores: nothing else ever touches
// those rows, so they need no zero-fill and no read-modify-write.
template <typename VT, bool MUL>
__device__ __forceinline__ void process_chunk(const uint4 &iw, const uint4 &mraw, const uint4 *vw, const VT *xv,
                                              const uint32_t *__restrict__ rowmap, VT *__restrict__ y, int lane,
                                              VT &carry, bool &open, uint32_t &next_rank, bool sole, bool &head_red) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t rank0 = mraw.x, valid = mraw.z & 0x3FFu, row_first = mraw.w;
  const bool consec = (mraw.z & kChunkRowsConsecutive) != 0;
  const bool partial = valid != (uint32_t)kChunkEntries;

  // end-of-row bits: gather the high byte of the 8 slots, then compress bit 7 of each byte with a multiply
  const uint32_t h0 = __byte_perm(iw.x, iw.y, 0x7531), h1 = __byte_perm(iw.z, iw.w, 0x7531);
  uint32_t eor = ((((h0 >> 7) & 0x01010101u) * 0x01020408u) >> 24) | (((((h1 >> 7) & 0x01010101u) * 0x01020408u) >> 20) & 0xF0u);
  if (partial) eor &= (1u << min(8, max(0, (int)valid - lane * 8))) - 1u;
  const int n_eor = __popc(eor);

  // inclusive prefix of n_eor across lanes -> rank of this lane's first row end
  int pre = n_eor;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(FULL, pre, d);
    if (lane >= d) pre += o;
  }
  const uint32_t total_eor = (uint32_t)__shfl_sync(FULL, pre, 31);
  const uint32_t rank_t = rank0 + (uint32_t)(pre - n_eor);

  // per-lane running sums of the products that restart after each end-of-row bit (add separately rounded from the
  // multiply, like the HLS cores: spmv.cpp:84-97)
  VT seg[8];
  VT acc = (lane == 0) ? carry : VT(0);
#pragma unroll
  for (int s = 0; s < 8; s++) {
    // MUL: xv holds the gathered x values and the multiply happens here, after the index work and the rank
    // shuffles above, which gives the gathers that much more time to land; otherwise xv already holds products
    acc = vadd(acc, MUL ? vmul(value_of<VT>(vw, s), xv[s]) : xv[s]);
    seg[s] = acc;
    if ((eor >> s) & 1u) acc = VT(0);
  }
  const bool seen = eor != 0;
  const uint32_t seen_mask = __ballot_sync(FULL, seen);

  // carry-in of every lane = open tails of the lanes before it back to the last lane that closed a row
  VT cin;
  if (seen_mask == FULL) {  // short rows: every lane closes at least one row
    cin = __shfl_up_sync(FULL, acc, 1);
    carry = __shfl_sync(FULL, acc, 31);
  } else {
    VT v = acc;
    bool f = seen;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const VT pv = __shfl_up_sync(FULL, v, d);
      const int pf = __shfl_up_sync(FULL, (int)f, d);
      if (lane >= d && !f) {
        v = vadd(pv, v);
        f = pf != 0;
      }
    }
    cin = __shfl_up_sync(FULL, v, 1);
    carry = __shfl_sync(FULL, v, 31);
  }
  if (lane == 0) cin = VT(0);  // lane 0 already absorbed the incoming carry

  const uint32_t first_bit = eor & (0u - eor);  // the lane's first row end also closes what earlier lanes left open

  // which row ends must be atomics: all of them unless the chunk is `sole`; then only the run's dangling first row
  uint32_t redm = 0xFFu;
  if (sole) redm = (head_red && (seen_mask & ((2u << lane) - 1u)) == (1u << lane)) ? first_bit : 0u;
  if (seen_mask) head_red = false;  // the row the run started in is closed by this chunk

  // Two copies of the emission loop on purpose: with a single loop and `consec ? row : rowmap[row]` the compiler
  // emits a predicated-off LDG whose scoreboard slot is shared with the x gathers already in flight for the next
  // chunk, and every store then waits for them (measured: +20 % kernel time).
  if (consec) {
    uint32_t row = row_first + (rank_t - rank0);
#pragma unroll
    for (int s = 0; s < 8; s++) {
      if ((eor >> s) & 1u) {
        VT v = seg[s];
        if ((first_bit >> s) & 1u) v = vadd(cin, v);
        if ((redm >> s) & 1u) y_add(y + row, v);
        else y[row] = v;
        row++;
      }
    }
  } else {
    uint32_t rk = rank_t;
#pragma unroll
    for (int s = 0; s < 8; s++) {
      if ((eor >> s) & 1u) {
        const uint32_t row = rowmap[rk];
        VT v = seg[s];
        if ((first_bit >> s) & 1u) v = vadd(cin, v);
        if ((redm >> s) & 1u) y_add(y + row, v);
        else y[row] = v;
        rk++;
      }
    }
  }

  // is a row still open after this chunk?  (a partial chunk ends its piece, and pieces end on a row end)
  open = !partial && ((__shfl_sync(FULL, eor, 31) >> 7) & 1u) == 0;
  if (partial) carry = VT(0);
  next_rank = rank0 + total_eor;
}

// The 8 x gathers of a lane.  Index 0 of a zero-padded slot is a valid address (x holds blocks * cols_div_blocks
// values), so the loads are unconditional and partial chunks mask the products instead.
// The loads are volatile asm so that they stay where they are written: issued one chunk ahead of their use.  (As plain
// loads the compiler sank them to the multiply when registers got tight, which silently removed the prefetch.)
__device__ __forceinline__ double ldg_x(const double *p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_x(const float *p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
template <typename VT>
__device__ __forceinline__ void gather_x(const uint4 &iw, const VT *__restrict__ x, uint32_t xbase, VT *xv) {
#pragma unroll
  for (int s = 0; s < 8; s++) xv[s] = ldg_x(x + (xbase + (idx16(iw, s) & 0x7FFFu)));  // 32-bit element index
}

// ------------------------------------------------------------------------------------------------------------------
// Variant DIRECT: every lane loads its group straight from global memory (5 / 3 x ld.global.v4); contiguous chunk
// range per warp, atomics for every row end.  Kept as the simple baseline the RING kernel is measured against.
template <typename VT, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    spmv_direct_kernel(const uint4 *__restrict__ stream,
                       const uint32_t *__restrict__ rowmap, const VT *__restrict__ x, VT *__restrict__ y,
                       unsigned long long n_chunks, uint32_t cdb) {
  constexpr int GW = VTraits<VT>::kGroupWords;
  constexpr int VW = VTraits<VT>::kValWords;
  const int lane = threadIdx.x & 31;
  const unsigned long long w = (unsigned long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const unsigned long long W = (unsigned long long)gridDim.x * WARPS;
  const unsigned long long c0 = n_chunks * w / W, c1 = n_chunks * (w + 1) / W;
  VT carry = VT(0);
  bool open = false;
  uint32_t next_rank = 0;
  for (unsigned long long c = c0; c < c1; c++) {
    const uint4 *slot = stream + c * (32 * GW + 1);  // device slot = chunk words + its ChunkMeta
    const uint4 mraw = __ldg(slot + 32 * GW);
    const uint4 *g = slot + lane * GW;
    uint4 iw = __ldg(g);
    uint4 vw[VW];
#pragma unroll
    for (int i = 0; i < VW; i++) vw[i] = __ldg(g + 1 + i);
    VT xv[8];
    gather_x<VT>(iw, x, mraw.y * cdb, xv);
    bool head_red = false;
    process_chunk<VT, true>(iw, mraw, vw, xv, rowmap, y, lane, carry, open, next_rank, false, head_red);
  }
  if (open && lane == 0) y_add(&y[rowmap[next_rank]], carry);
}

// ------------------------------------------------------------------------------------------------------------------
// Variant RING: per-warp ring of STAGES chunks filled by cp.async.bulk (one elected lane issues, an mbarrier per
// stage counts the bytes); lanes read their group with conflict-free LDS.128 (lane stride 80 B / 48 B).
// Software pipeline per warp, chunk i being summed while chunk i+1's x values are in flight:
//   iteration i:  wait stage(i+1) -> LDS its index word -> issue its 8 x gathers (not consumed until i+1)
//                 LDS chunk i's value words -> refill stage(i-1) with chunk i-1+STAGES -> segmented sums of chunk i
// Chunk assignment: warp w of W walks runs of R = 2^run_log2 consecutive chunks, run q of the warp being global run
// q*W + w.  All warps therefore sweep one contiguous window of the stream together (DRAM page locality; contiguous
// per-warp ranges measured ~25 % slower), while inside a run the open row sum stays in registers.
template <typename VT, int WARPS, int STAGES, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    spmv_ring_kernel(const uint4 *__restrict__ stream,
                     const uint32_t *__restrict__ rowmap, const VT *__restrict__ x, VT *__restrict__ y,
                     uint32_t n_chunks, uint32_t cdb, uint32_t run_log2, uint32_t flags) {
  // flags bit 2: atomics everywhere (y += A x semantics: plain stores would overwrite the caller's y)
  static_assert((STAGES & (STAGES - 1)) == 0 && STAGES >= 4,
                "prefetch one chunk ahead + refill one chunk behind needs >= 3 stages (power of two: 4)");
  constexpr int GW = VTraits<VT>::kGroupWords;
  constexpr int VW = VTraits<VT>::kValWords;
  constexpr uint32_t CHUNK_BYTES = GW * 16 * 32;
  constexpr uint32_t SLOT = CHUNK_BYTES + 16;  // a device slot: the chunk's words followed by its ChunkMeta
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t ring = smem_u32(smem) + (uint32_t)warp * STAGES * SLOT;   // shared-window addresses
  const uint32_t bars = smem_u32(smem) + WARPS * STAGES * SLOT + (uint32_t)warp * STAGES * 8;
  const uint32_t my = ring + lane * (GW * 16);                                    // this lane's group in stage 0
  const uint32_t w = blockIdx.x * WARPS + warp, W = gridDim.x * WARPS;
  const uint32_t R = 1u << run_log2;
  const uint32_t total_runs = (n_chunks + R - 1) >> run_log2;
  if (w >= total_runs) return;
  const uint32_t my_runs = (total_runs - w + W - 1) / W;
  uint32_t n = my_runs << run_log2;  // chunks this warp walks
  {
    const uint32_t over = ((my_runs - 1) * W + w + 1) << run_log2;  // a partial last run can only be the global last
    if (over > n_chunks) n -= over - n_chunks;
  }
  const uint32_t jump = (W - 1) << run_log2;  // extra chunk distance when stepping from one run into the next
  // global chunk index of the warp's (i + k)-th chunk, given ci = index of its i-th, for 0 <= k <= R
  auto ahead = [&](uint32_t ci, uint32_t i, uint32_t k) -> uint32_t {
    return ci + k + ((((i & (R - 1)) + k) >> run_log2) ? jump : 0u);
  };
  const bool force_red = (flags & 4u) != 0;
  auto lds128 = [](uint32_t a) -> uint4 {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
  };
  auto issue = [&](uint32_t stage, uint32_t chunk) {  // lane 0 only
    const uint32_t bar = bars + stage * 8;
    mbar_expect_tx(bar, SLOT);
    bulk_g2s(ring + stage * SLOT, stream + (size_t)chunk * (32 * GW + 1), SLOT, bar);
  };

  uint32_t c_cur = w << run_log2;  // global index of chunk i (i = 0)
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) mbar_init(bars + s * 8, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++)
      if ((uint32_t)s < n) issue(s, ahead(c_cur, 0, s));
  }

  VT carry = VT(0);
  bool open = false, head_red = false;
  uint32_t next_rank = 0;

  // registers of the chunk being prefetched / processed: [i & 1]
  uint4 iw[2], mr[2];
  VT xv[2][8];
  // prologue: prefetch chunk 0
  {
    mbar_wait(bars, 0);
    mr[0] = lds128(ring + CHUNK_BYTES);
    iw[0] = lds128(my);
    gather_x<VT>(iw[0], x, mr[0].y * cdb, xv[0]);
  }
  // one pipeline step; CUR is a compile-time 0/1 so that the double-buffered registers are never indexed dynamically
  auto step = [&](auto CUR, uint32_t i) {
    constexpr int cur = decltype(CUR)::value, nxt = cur ^ 1;
    // ---- prefetch chunk i+1: its stage has landed or is about to; leaves the 8 gathers in flight
    if (i + 1 < n) {
      const uint32_t st = (i + 1) & (STAGES - 1);
      mbar_wait(bars + st * 8, ((i + 1) / STAGES) & 1u);
      mr[nxt] = lds128(ring + st * SLOT + CHUNK_BYTES);  // the chunk's meta travels with it: no separate global load
      iw[nxt] = lds128(my + st * SLOT);
      gather_x<VT>(iw[nxt], x, mr[nxt].y * cdb, xv[nxt]);
    }
    // ---- values of chunk i and its products with the x values gathered one step ago
    VT prod[8];
    {
      uint4 vw[VW];
      const uint32_t a = my + (i & (STAGES - 1)) * SLOT + 16;
#pragma unroll
      for (int k = 0; k < VW; k++) vw[k] = lds128(a + k * 16);
#pragma unroll
      for (int s = 0; s < 8; s++) prod[s] = vmul(value_of<VT>(vw, s), xv[cur][s]);
    }
    // Refill the stage consumed in the PREVIOUS iteration: its registers went through process_chunk, whose warp
    // shuffles order every lane's LDS before this point.  (Refilling the stage just read is a race: the bulk copy
    // runs in the async proxy and can overtake LDS still queued in the LSU - seen as rare wrong rows on B200.)
    __syncwarp();
    if (lane == 0 && i >= 1 && i - 1 + STAGES < n) issue((i - 1) & (STAGES - 1), ahead(c_cur, i, STAGES - 1));
    // ---- segmented sums + y updates of chunk i
    const uint32_t pos = i & (R - 1);
    const bool sole = (mr[cur].z & kChunkSole) != 0 && !force_red;
    if (pos == 0) head_red = (mr[cur].z & kChunkStartsMid) != 0;  // stays set until the run's first row end
    process_chunk<VT, false>(iw[cur], mr[cur], nullptr, prod, rowmap, y, lane, carry, open, next_rank, sole, head_red);
    if (pos == R - 1 || i + 1 == n) {  // the row left open continues in another warp's run: hand over atomically
      if (open && lane == 0) y_add(&y[rowmap[next_rank]], carry);
      carry = VT(0);
      open = false;
    }
    c_cur = ahead(c_cur, i, 1);
  };
  // two chunks per trip without a join in between: the compiler then keeps the prefetched x values where the loads
  // put them (a conditional second half made it copy them right after issue, i.e. wait for them - 30 % slower)
  uint32_t i = 0;
  for (; i + 1 < n; i += 2) {
    step(std::integral_constant<int, 0>{}, i);
    step(std::integral_constant<int, 1>{}, i + 1);
  }
  if (i < n) step(std::integral_constant<int, 0>{}, i);
}

// ------------------------------------------------------------------------------------------------------------------
// Variant OCC: the same TMA ring and run-interleaved walk, but no software prefetch of x: two stages per warp and
// few enough registers for MINB resident CTAs per SM, so that the x gather latency of one warp is covered by the
// other warps of the SM sub-partition (classic occupancy-based hiding; nothing stays in flight in registers across
// the row-sum code, so it does not depend on how ptxas assigns scoreboard slots).
//   step i:  wait stage(i) -> LDS group + meta -> 8 gathers -> products -> segmented sums / y updates
//            -> refill stage(i) with chunk i+2 (every lane's LDS is ordered before it by process_chunk's shuffles)
template <typename VT, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    spmv_occ_kernel(const uint4 *__restrict__ stream, const uint32_t *__restrict__ rowmap,
                    const VT *__restrict__ x, VT *__restrict__ y, uint32_t n_chunks, uint32_t cdb,
                    uint32_t run_log2, uint32_t flags) {
  constexpr int STAGES = 2;
  constexpr int GW = VTraits<VT>::kGroupWords;
  constexpr int VW = VTraits<VT>::kValWords;
  constexpr uint32_t CHUNK_BYTES = GW * 16 * 32;
  constexpr uint32_t SLOT = CHUNK_BYTES + 16;
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t ring = smem_u32(smem) + (uint32_t)warp * STAGES * SLOT;
  const uint32_t bars = smem_u32(smem) + WARPS * STAGES * SLOT + (uint32_t)warp * STAGES * 8;
  const uint32_t my = ring + lane * (GW * 16);
  const uint32_t w = blockIdx.x * WARPS + warp, W = gridDim.x * WARPS;
  const uint32_t R = 1u << run_log2;
  const uint32_t total_runs = (n_chunks + R - 1) >> run_log2;
  if (w >= total_runs) return;
  const uint32_t my_runs = (total_runs - w + W - 1) / W;
  uint32_t n = my_runs << run_log2;
  {
    const uint32_t over = ((my_runs - 1) * W + w + 1) << run_log2;
    if (over > n_chunks) n -= over - n_chunks;
  }
  const uint32_t jump = (W - 1) << run_log2;
  auto ahead = [&](uint32_t ci, uint32_t i, uint32_t k) -> uint32_t {
    return ci + k + ((((i & (R - 1)) + k) >> run_log2) ? jump : 0u);
  };
  const bool force_red = (flags & 4u) != 0;
  auto lds128 = [](uint32_t a) -> uint4 {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
  };
  auto issue = [&](uint32_t stage, uint32_t chunk) {  // lane 0 only
    const uint32_t bar = bars + stage * 8;
    mbar_expect_tx(bar, SLOT);
    bulk_g2s(ring + stage * SLOT, stream + (size_t)chunk * (32 * GW + 1), SLOT, bar);
  };
  uint32_t c_cur = w << run_log2;
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (lane == 0) {
    issue(0, c_cur);
    if (1 < n) issue(1, ahead(c_cur, 0, 1));
  }
  VT carry = VT(0);
  bool open = false, head_red = false;
  uint32_t next_rank = 0;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t st = i & 1u;
    mbar_wait(bars + st * 8, (i >> 1) & 1u);
    const uint32_t base = my + st * SLOT;
    const uint4 mraw = lds128(ring + st * SLOT + CHUNK_BYTES);
    const uint4 iw = lds128(base);
    VT xv[8];
    gather_x<VT>(iw, x, mraw.y * cdb, xv);
    uint4 vw[VW];
#pragma unroll
    for (int k = 0; k < VW; k++) vw[k] = lds128(base + 16 + k * 16);
    const uint32_t pos = i & (R - 1);
    const bool sole = (mraw.z & kChunkSole) != 0 && !force_red;
    if (pos == 0) head_red = (mraw.z & kChunkStartsMid) != 0;
    process_chunk<VT, true>(iw, mraw, vw, xv, rowmap, y, lane, carry, open, next_rank, sole, head_red);
    if (pos == R - 1 || i + 1 == n) {
      if (open && lane == 0) y_add(&y[rowmap[next_rank]], carry);
      carry = VT(0);
      open = false;
    }
    __syncwarp();
    if (lane == 0 && i + 2 < n) issue(st, ahead(c_cur, i, 2));
    c_cur = ahead(c_cur, i, 1);
  }
}

// y[rows[i]] = 0 for the rows that receive atomics or no update at all (Layout::zero_rows)
template <typename VT>
__global__ void zero_rows_kernel(VT *__restrict__ y, const uint32_t *__restrict__ rows, uint32_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[rows[i]] = VT(0);
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
