// sm_100a SpMV kernels over the hw_matrix stream (the reference's HLS dataflow pipeline, src/spmv.cpp:6-205,
// fused with the host accumulation accum_results, src/csr_hw.cpp:1531-1565).
//
// Mapping of the reference stages:
//   read_data_submatrix  (spmv.cpp:6-34)    -> chunk fetch: a per-warp ring of slots filled by cp.async.bulk (TMA 1-D,
//                                              SASS UBLKCP) with an mbarrier per slot and an evict-first L2 policy;
//                                              lanes read their group with conflict-free LDS.128 (kernels OCC, XS);
//                                              ld.global.v4 per lane in the DIRECT baseline
//   stream_data_col_ind  (spmv.cpp:36-49)   -> in-register unpack of 8 x (15-bit column | end-of-row bit)
//   stream_data_values   (spmv.cpp:51-64)   -> in-register reinterpretation of the value words
//   compute_results      (spmv.cpp:66-104)  -> per-lane left-to-right multiply/add over its 8 entries (mul and add
//                                              separately rounded, like the HLS cores) + warp segmented scan keyed on
//                                              the end-of-row bit
//   write_back_results + accum_results      -> st.global / red.global.add to y[rowmap[rank]] (rank = running count of
//                                              end-of-row bits), no partial-y buffers, no host pass
//   x slice copy L0      (spmv.cpp:182-192) -> kernel XS: cp.async.bulk of the x window of a work item into shared
//                                              memory, gathers with LDS; kernel OCC gathers x through L1/L2
//
// Work unit: a "chunk" = 32 consecutive 8-entry groups of one piece (one group per lane).  Warps walk runs of
// consecutive chunks (walk_chunks), carry the open row sum in registers inside a run, and hand a row that continues
// into another warp's run over through an atomic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "ell.h"
#include "layout.h"

namespace spmvb {

// ---- bounds-checked build (make check: lib/libspmvb_check.so, -DSPMVB_CHECK_BOUNDS).  compute-sanitizer is not available
// on every pool, so the kernels can check their own addresses: every index that reaches memory - chunk slot, row-map
// entry, y row, x element (global gather) or x window offset (shared-memory gather) - is compared with the limits the
// engine stores in g_limits before a launch; a violation is counted per class and the access is redirected to
// element 0.  The release build compiles all of this away.
struct CheckLimits { unsigned long long n_chunks, n_pairs, rows, x_len; };
#ifdef SPMVB_CHECK_BOUNDS
__device__ CheckLimits g_limits;
__device__ unsigned long long g_bounds_errors[5];  // chunk index, row-map index, y row, x index, x window offset
__device__ __forceinline__ uint32_t check_bound(int cls, uint32_t idx, unsigned long long limit) {
  if ((unsigned long long)idx < limit) return idx;
  atomicAdd(&g_bounds_errors[cls], 1ull);
  return 0u;
}
#define SPMVB_BOUND(cls, idx, limit) check_bound(cls, (uint32_t)(idx), (unsigned long long)(limit))
#else
#define SPMVB_BOUND(cls, idx, limit) (idx)
#endif

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
// L2 eviction policy for data that is read exactly once (the matrix stream): evict first, so that it does not push
// x and the y range being updated out of the L2 cache
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// y update / x gather with an explicit L2 policy ("tall" matrices: x and y both exceed the L2 cache; the y range in
// flight must stay resident, x has no reuse left once the pieces are walked CU-major)
__device__ __forceinline__ void y_add_hint(double *p, double v, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void y_add_hint(float *p, float v, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ double ldg_x_hint(const double *p, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ldg_x_hint(const float *p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine; SASS UBLKCP), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// programmatic dependent launch: wait for the previous kernel of the stream / let the next one start early
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
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
// COAL (wide images, whose rows ascend through a chunk): the row sums are parked in `scratch` (shared memory, one
// VT per row end of the chunk, in rank order) and written out by the whole warp, lane q taking row ends q, q + 32, ...:
// consecutive rows per request - 4 fp64 rows per 32-byte sector - instead of one sector per lane, and the row ids of
// non-consecutive chunks come as coalesced loads.
__device__ __forceinline__ void sts_v(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ double lds_v(uint32_t a, double) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds_v(uint32_t a, float) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}

template <typename VT, bool MUL, bool COAL = false>
__device__ __forceinline__ void process_chunk(const uint4 &iw, const uint4 &mraw, const uint4 *vw, const VT *xv,
                                              const uint32_t *__restrict__ rowmap, VT *__restrict__ y, int lane,
                                              VT &carry, bool &open, uint32_t &next_rank, bool sole, bool &head_red,
                                              uint64_t y_policy = 0, uint64_t stream_policy = 0, uint32_t scratch = 0,
                                              const uint32_t *rows_pre = nullptr) {
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

  // Row ids of this lane's row ends, all requested now, back to back, so that their latency overlaps the products
  // and the row sums (fetched one by one inside the update loop below they were 35 % of the XS kernel's stall time
  // on R-MAT: load -> wait -> RED -> next load ...).
  uint32_t rows[8];
  if (!COAL && !consec) {
    uint32_t rk = rank_t;
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const uint32_t e = (eor >> s) & 1u;
      rows[s] = 0;
#ifdef SPMVB_CHECK_BOUNDS
      if (e) rk = SPMVB_BOUND(1, rk, g_limits.n_pairs);
#endif
      if (y_policy)  // tall matrix: the row map is read once, it must not push the y tile out of the L2 cache
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.global.nc.L2::cache_hint.u32 %0, [%1], %3;\n\t}"
                     : "+r"(rows[s])
                     : "l"(rowmap + rk), "r"(e), "l"(stream_policy));
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.global.nc.u32 %0, [%1];\n\t}"
                     : "+r"(rows[s])
                     : "l"(rowmap + rk), "r"(e));
      rk += e;
    }
  }

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

  if (COAL) {
    const bool head = head_red;  // the chunk's first row end closes the row the run started in: an atomic even if `sole`
    if (seen_mask) head_red = false;
    uint32_t total_eor_emit = total_eor;
    uint32_t q = (uint32_t)(pre - n_eor);
#pragma unroll
    for (int s = 0; s < 8; s++) {
      if ((eor >> s) & 1u) {
        VT v = seg[s];
        if ((first_bit >> s) & 1u) v = vadd(cin, v);
        sts_v(scratch + q * (uint32_t)sizeof(VT), v);
        q++;
      }
    }
    __syncwarp();
    if (stream_policy == ~0ull) total_eor_emit = 0;  // diagnostic (diag_flags 32): no y updates, wrong results
    // rows_pre[j] = row id of row end lane + 32 j, requested by the caller together with the x gathers (a load inside
    // this loop would put one memory latency between every two updates)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      q = (uint32_t)lane + 32u * (uint32_t)j;
      if (q < total_eor_emit) {
        const VT v = lds_v(scratch + q * (uint32_t)sizeof(VT), VT(0));
        const uint32_t row = SPMVB_BOUND(2, consec ? row_first + q : rows_pre[j], g_limits.rows);
        if (!sole || (q == 0 && head)) {
          if (y_policy) y_add_hint(y + row, v, y_policy);
          else y_add(y + row, v);
        } else {
          y[row] = v;
        }
      }
    }
    __syncwarp();  // the scratch is free for the next chunk
    open = !partial && ((__shfl_sync(FULL, eor, 31) >> 7) & 1u) == 0;
    if (partial) carry = VT(0);
    next_rank = rank0 + total_eor;
    return;
  }

  if (y_policy == ~0ull) eor = 0;  // diagnostic (engine option diag_flags bit 5): no y updates at all, wrong results
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
        const uint32_t rr = SPMVB_BOUND(2, row, g_limits.rows);
        if ((redm >> s) & 1u) y_add(y + rr, v);
        else y[rr] = v;
        row++;
      }
    }
  } else {
#pragma unroll
    for (int s = 0; s < 8; s++) {
      if ((eor >> s) & 1u) {
        const uint32_t row = SPMVB_BOUND(2, rows[s], g_limits.rows);
        VT v = seg[s];
        if ((first_bit >> s) & 1u) v = vadd(cin, v);
        if ((redm >> s) & 1u) {
          if (y_policy) y_add_hint(y + row, v, y_policy);
          else y_add(y + row, v);
        } else {
          y[row] = v;
        }
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
__device__ __forceinline__ void gather_x(const uint4 &iw, const VT *__restrict__ x, uint32_t xbase, VT *xv,
                                         uint64_t policy = 0) {
  if (policy) {
#pragma unroll
    for (int s = 0; s < 8; s++) xv[s] = ldg_x_hint(x + SPMVB_BOUND(3, xbase + (idx16(iw, s) & 0x7FFFu), g_limits.x_len), policy);
  } else {
#pragma unroll
    for (int s = 0; s < 8; s++)  // 32-bit element index
      xv[s] = ldg_x(x + SPMVB_BOUND(3, xbase + (idx16(iw, s) & 0x7FFFu), g_limits.x_len));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Variant DIRECT: every lane loads its group straight from global memory (5 / 3 x ld.global.v4); contiguous chunk
// range per warp, atomics for every row end.  Kept as the simple baseline the TMA-ring kernels (OCC, XS) are measured against.
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
    gather_x<VT>(iw, x, (mraw.y & kMetaBlockMask) * cdb, xv);
    bool head_red = false;
    process_chunk<VT, true>(iw, mraw, vw, xv, rowmap, y, lane, carry, open, next_rank, false, head_red);
  }
  if (open && lane == 0) y_add(&y[rowmap[next_rank]], carry);
}

// ------------------------------------------------------------------------------------------------------------------
// The warp-level chunk walk shared by the OCC and XS kernels: a 2-stage TMA ring per warp, no software prefetch of x.
//   step i:  wait stage -> LDS group + meta -> 8 x gathers (functor) -> LDS values -> segmented sums / y updates
//            -> refill the stage with chunk i+2 (every lane's LDS is ordered before it by process_chunk's shuffles)
// The walk covers the chunk domain [base, base + n_dom): warp w of W takes runs of R = 2^run_log2 consecutive chunks,
// run q of the warp being run q*W + w of the domain, so that the W warps sweep one contiguous window of the stream
// together (DRAM page locality) while the open row sum stays in registers inside a run.
// `t` is the warp's running slot counter: it carries the ring stage / mbarrier phase from one domain to the next.
// (Two ways of bringing a chunk's row ids through shared memory instead of the lanes' eight scattered global loads were
// built and measured in round 2 - a second bulk copy on the chunk's mbarrier, and coalesced loads by the whole warp parked
// in the chunk's own ring slot: 3.74 vs 3.65 ms on the 0.5 B-nnz uniform matrix for the first, 8.68 vs 8.27 ms on the
// 1 B-nnz one and 1.73 vs 1.55 ms on R-MAT for the second.  The kernel is bound by its instruction chain, not by the
// L1 tag stage of those loads; the hoisted loads of process_chunk stay.)
template <typename VT, typename Gather>
__device__ __forceinline__ void walk_chunks(const uint4 *__restrict__ stream, const uint32_t *__restrict__ rowmap,
                                            VT *__restrict__ y, uint32_t ring, uint32_t bars, int lane, uint32_t base,
                                            uint32_t n_dom, uint32_t w, uint32_t W, uint32_t run_log2, bool force_red,
                                            uint32_t &t, uint64_t y_policy, bool dep_wait, Gather gather) {
  constexpr int GW = VTraits<VT>::kGroupWords;
  constexpr int VW = VTraits<VT>::kValWords;
  constexpr uint32_t CHUNK_BYTES = GW * 16 * 32;
  constexpr uint32_t SLOT = CHUNK_BYTES + 16;
  constexpr uint32_t STAGE = SLOT;
  const uint32_t R = 1u << run_log2;
  const uint32_t total_runs = (n_dom + R - 1) >> run_log2;
  if (w >= total_runs) return;
  const uint32_t my_runs = (total_runs - w + W - 1) / W;
  uint32_t n = my_runs << run_log2;  // chunks this warp walks
  {
    const uint32_t over = ((my_runs - 1) * W + w + 1) << run_log2;  // a partial last run can only be the domain's last
    if (over > n_dom) n -= over - n_dom;
  }
  const uint32_t jump = (W - 1) << run_log2;  // extra chunk distance when stepping from one run into the next
  auto ahead = [&](uint32_t ci, uint32_t i, uint32_t k) -> uint32_t {  // index of chunk i + k given chunk i (any k, R)
    return ci + k + ((((i & (R - 1)) + k) >> run_log2) * jump);
  };
  auto lds128 = [](uint32_t a) -> uint4 {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
  };
  const uint64_t stream_policy = l2_policy_evict_first();
  auto issue = [&](uint32_t slot, uint32_t chunk) {  // lane 0 only
    chunk = SPMVB_BOUND(0, chunk, g_limits.n_chunks);
    const uint32_t bar = bars + (slot & 1u) * 8;
    mbar_expect_tx(bar, SLOT);
    bulk_g2s_hint(ring + (slot & 1u) * STAGE, stream + (size_t)chunk * (32 * GW + 1), SLOT, bar, stream_policy);
  };
  const uint32_t my = ring + lane * (GW * 16);
  uint32_t c_cur = base + (w << run_log2);
  if (lane == 0) {
    issue(t, c_cur);
    if (1 < n) issue(t + 1, ahead(c_cur, 0, 1));
  }
  // the matrix stream never changes, so its first chunks are already in flight; x and y may still be written by
  // the previous kernel of the stream (row clearing, x <- y / ||y|| of an iterated caller)
  if (dep_wait) grid_dep_wait();
  VT carry = VT(0);
  bool open = false, head_red = false;
  uint32_t next_rank = 0;
  for (uint32_t i = 0; i < n; i++, t++) {
    const uint32_t st = t & 1u;
    mbar_wait(bars + st * 8, (t >> 1) & 1u);
    const uint32_t g = my + st * STAGE;
    const uint4 mraw = lds128(ring + st * STAGE + CHUNK_BYTES);
    const uint4 iw = lds128(g);
    VT xv[8];
    gather(iw, mraw, xv);
    uint4 vw[VW];
#pragma unroll
    for (int k = 0; k < VW; k++) vw[k] = lds128(g + 16 + k * 16);
    const uint32_t pos = i & (R - 1);
    const bool sole = (mraw.z & kChunkSole) != 0 && !force_red;
    if (pos == 0) head_red = (mraw.z & kChunkStartsMid) != 0;  // stays set until the run's first row end
    process_chunk<VT, true>(iw, mraw, vw, xv, rowmap, y, lane, carry, open, next_rank, sole, head_red, y_policy,
                            stream_policy);
    if (pos == R - 1 || i + 1 == n) {  // the row left open continues in another warp's run: hand over atomically
      if (open && lane == 0) {
        // in a `consecutive` chunk the open row follows from the rank (the flag covers it): no dependent row-map load
        uint32_t row = (mraw.z & kChunkRowsConsecutive) ? mraw.w + (next_rank - mraw.x)
                                                        : rowmap[SPMVB_BOUND(1, next_rank, g_limits.n_pairs)];
        row = SPMVB_BOUND(2, row, g_limits.rows);
        y_add(&y[row], carry);
      }
      carry = VT(0);
      open = false;
    }
    __syncwarp();
    if (lane == 0 && i + 2 < n) issue(t + 2, ahead(c_cur, i, 2));
    c_cur = ahead(c_cur, i, 1);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Variant OCC: x gathered from global memory (L1/L2); few enough registers for MINB resident CTAs per SM, so that
// the gather latency of one warp is covered by the other warps of its SM sub-partition.  Nothing stays in flight in
// registers across the row-sum code, so the kernel does not depend on how ptxas assigns scoreboard slots.
template <typename VT, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    spmv_occ_kernel(const uint4 *__restrict__ stream, const uint32_t *__restrict__ rowmap,
                    const VT *__restrict__ x, VT *__restrict__ y, uint32_t n_chunks, uint32_t cdb,
                    uint32_t run_log2, uint32_t flags) {
  constexpr uint32_t SLOT = VTraits<VT>::kGroupWords * 16 * 32 + 16;
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t ring = smem_u32(smem) + (uint32_t)warp * 2 * SLOT;
  const uint32_t bars = smem_u32(smem) + WARPS * 2 * SLOT + (uint32_t)warp * 16;
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    fence_barrier_init();
  }
  __syncwarp();
  uint32_t t = 0;
  // flags bit 3: "tall" matrix (x and y both larger than the L2 cache): y updates evict-last, x gathers evict-first
  const uint64_t y_policy = (flags & 8u) ? l2_policy_evict_last() : 0ull;
  const uint64_t x_policy = (flags & 8u) ? l2_policy_evict_first() : 0ull;
  walk_chunks<VT>(stream, rowmap, y, ring, bars, lane, 0u, n_chunks, blockIdx.x * WARPS + warp,
                         gridDim.x * WARPS, run_log2, (flags & 4u) != 0, t, y_policy, true,
                         [&](const uint4 &iw, const uint4 &mraw, VT *xv) {
                           gather_x<VT>(iw, x, (mraw.y & kMetaBlockMask) * cdb, xv, x_policy);
                         });
}

// ------------------------------------------------------------------------------------------------------------------
// Variant XS: the reference's "x slice in on-chip memory" (spmv.cpp:180-192: every compute unit copies the block's x
// slice to BRAM before streaming the block).  The stream is cut into work items = a range of chunks of ONE column
// block whose entries touch a window of at most X_CAP bytes of that block's x slice.  Every CTA (one per SM) owns one
// contiguous, equally long range of the stream = a few consecutive items; per item one thread TMA-bulk-copies the x window into shared memory (cp.async.bulk, several
// pieces on one mbarrier) while the warps already fetch their first chunks, then all warps walk the item's chunks and
// gather x with LDS (2-4 wavefronts per gather instead of 12-32 L1 sectors).  Items whose window does not fit (wide
// blocks of an irregular matrix in fp64) fall back to global gathers.

__device__ __forceinline__ double lds_x(uint32_t a, double) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_x(uint32_t a, float) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}

template <typename VT, int WARPS, uint32_t X_CAP, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    spmv_xs_kernel(const uint4 *__restrict__ stream, const uint32_t *__restrict__ rowmap, const VT *__restrict__ x,
                   VT *__restrict__ y, const XsItem *__restrict__ items, const uint32_t *__restrict__ cta_first,
                   uint32_t cdb, uint32_t run_log2, uint32_t flags) {
  constexpr uint32_t SLOT = VTraits<VT>::kGroupWords * 16 * 32 + 16;
  constexpr uint32_t STAGE = SLOT;
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t xbuf = smem_u32(smem);
  const uint32_t ring = xbuf + X_CAP + (uint32_t)warp * 2 * STAGE;
  const uint32_t bars = xbuf + X_CAP + WARPS * 2 * STAGE + (uint32_t)warp * 16;
  const uint32_t xbar = xbuf + X_CAP + WARPS * 2 * STAGE + WARPS * 16;
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    if (warp == 0) mbar_init(xbar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  const uint64_t y_policy = (flags & 8u) ? l2_policy_evict_last() : 0ull;  // "tall" matrix: keep the y range in L2
  grid_dep_wait();  // the first thing an item does is copy a window of x, which the previous kernel may have written
  uint32_t t = 0, k = 0;  // k counts the windows loaded so far (phase of xbar)
  const uint32_t it_end = __ldg(cta_first + blockIdx.x + 1);
  for (uint32_t it = __ldg(cta_first + blockIdx.x); it < it_end; it++) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(items + it));
    const uint4 b = __ldg(reinterpret_cast<const uint4 *>(items + it) + 1);
    const uint32_t chunk_begin = a.x, chunk_count = a.y, x_off = a.z, x_bytes = a.w, col_base = b.x;
    __syncthreads();  // every warp is done gathering from the previous window
    if (threadIdx.x == 0 && x_bytes && (flags & 16u)) {  // diagnostic: no window traffic (results are wrong)
      mbar_expect_tx(xbar, 16u);
      bulk_g2s(xbuf, x, 16u, xbar);
    } else if (threadIdx.x == 0 && x_bytes) {
      fence_proxy_async();
      mbar_expect_tx(xbar, x_bytes);
      const uint8_t *src = reinterpret_cast<const uint8_t *>(x + x_off);
      if (y_policy) {  // tall matrix: x is streamed once per row tile - it must not displace the tile's y range in L2
        const uint64_t xpol = l2_policy_evict_first();
        for (uint32_t o = 0; o < x_bytes; o += 16384u) bulk_g2s_hint(xbuf + o, src + o, min(16384u, x_bytes - o), xbar, xpol);
      } else {
        for (uint32_t o = 0; o < x_bytes; o += 16384u) bulk_g2s(xbuf + o, src + o, min(16384u, x_bytes - o), xbar);
      }
    }
    if (x_bytes) {
      bool waited = false;
      walk_chunks<VT>(stream, rowmap, y, ring, bars, lane, chunk_begin, chunk_count, (uint32_t)warp,
                              (uint32_t)WARPS, run_log2, (flags & 4u) != 0, t, (flags & 32u) ? ~0ull : y_policy, false,
                              [&](const uint4 &iw, const uint4 &mraw, VT *xv) {
                                if (!waited) {  // first chunk of the item: the window must have landed
                                  mbar_wait(xbar, k & 1u);
                                  waited = true;
                                }
                                const uint32_t valid = mraw.z & 0x3FFu;
                                const uint32_t xs = xbuf - col_base * (uint32_t)sizeof(VT);
                                auto win = [&](uint32_t col) -> uint32_t {  // shared-memory address of x[col] of the block
#ifdef SPMVB_CHECK_BOUNDS
                                  const uint32_t off = (col - col_base) * (uint32_t)sizeof(VT);
                                  return xbuf + SPMVB_BOUND(4, off, x_bytes);
#else
                                  return xs + col * (uint32_t)sizeof(VT);
#endif
                                };
                                if (valid == (uint32_t)kChunkEntries) {
#pragma unroll
                                  for (int s = 0; s < 8; s++) xv[s] = lds_x(win(idx16(iw, s) & 0x7FFFu), VT(0));
                                } else {  // padding slots carry column 0, which may lie outside the window
                                  const int nv = min(8, max(0, (int)valid - lane * 8));
#pragma unroll
                                  for (int s = 0; s < 8; s++) xv[s] = s < nv ? lds_x(win(idx16(iw, s) & 0x7FFFu), VT(0)) : VT(0);
                                }
                              });
      if (!waited) mbar_wait(xbar, k & 1u);  // warps without work still consume the phase
      k++;
    } else {
      walk_chunks<VT>(stream, rowmap, y, ring, bars, lane, chunk_begin, chunk_count, (uint32_t)warp,
                              (uint32_t)WARPS, run_log2, (flags & 4u) != 0, t, 0ull, false,
                              [&](const uint4 &iw, const uint4 &mraw, VT *xv) {
                                gather_x<VT>(iw, x, (mraw.y & kMetaBlockMask) * cdb, xv);
                              });
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Variant WIDE: the kernel of the wide image (layout.h: Layout::wide).  The column blocks are as wide as the L2 cache
// can hold of x (up to 2^23 columns) and the rows ascend through a block, so the scattered 8-byte access of every
// entry is the x gather - ld.global.nc out of an L2-resident range, the cheapest scattered access a B200 has
// (tools/access_probe.cu: 288 G/s against 193 G/s for red.global.add) - while the row sums of a (row, block) pair are
// formed in registers and leave as coalesced requests (process_chunk<COAL>).  Same machinery otherwise: per-warp TMA
// ring of chunk slots, runs of chunks with the open row carried in registers.  A slot is the planar chunk of
// layout.h (kWide*) + its ChunkMeta; lane l reads its index word, its 8 high column bytes and its value words with
// conflict-free LDS.128 / LDS.64.
//   flags bit 2: accumulate (all updates atomics); bit 3: x gathers evict-last; bit 4: y updates + row map evict-first;
//   bits 5, 6: diagnostics (no y updates / gathers confined to a small window of x)
template <typename VT, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    spmv_wide_kernel(const uint4 *__restrict__ stream, const uint32_t *__restrict__ rowmap, const VT *__restrict__ x,
                     VT *__restrict__ y, uint32_t chunk_base, uint32_t n_chunks, uint32_t cdb, uint32_t run_log2,
                     uint32_t flags) {
  constexpr int VW = VTraits<VT>::kValWords;
  constexpr uint32_t CHUNK_BYTES = kWideValOff + 32u * 8u * (uint32_t)sizeof(VT);
  constexpr uint32_t SLOT = CHUNK_BYTES + 16u;
  constexpr uint32_t SCRATCH = (uint32_t)kChunkEntries * (uint32_t)sizeof(VT);
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t ring = smem_u32(smem) + (uint32_t)warp * 2 * SLOT;
  const uint32_t scratch = smem_u32(smem) + WARPS * 2 * SLOT + (uint32_t)warp * SCRATCH;
  const uint32_t bars = smem_u32(smem) + WARPS * (2 * SLOT + SCRATCH) + (uint32_t)warp * 16;
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    fence_barrier_init();
  }
  __syncwarp();
  const uint64_t x_policy = (flags & 8u) ? l2_policy_evict_last() : 0ull;
  const uint64_t y_policy = (flags & 16u) ? l2_policy_evict_first() : 0ull;
  const uint64_t stream_policy = l2_policy_evict_first();
  const bool force_red = (flags & 4u) != 0;

  // the walk of walk_chunks: warp w of W takes run q*W + w of the domain [chunk_base, chunk_base + n_chunks)
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
    return ci + k + ((((i & (R - 1)) + k) >> run_log2) * jump);
  };
  auto lds128 = [](uint32_t a) -> uint4 {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
  };
  auto issue = [&](uint32_t slot, uint32_t chunk) {  // lane 0 only
    chunk = SPMVB_BOUND(0, chunk, g_limits.n_chunks);
    const uint32_t bar = bars + (slot & 1u) * 8;
    mbar_expect_tx(bar, SLOT);
    bulk_g2s_hint(ring + (slot & 1u) * SLOT, reinterpret_cast<const uint8_t *>(stream) + (size_t)chunk * SLOT, SLOT, bar,
                  stream_policy);
  };
  uint32_t c_cur = chunk_base + (w << run_log2);
  uint32_t t = 0;
  if (lane == 0) {
    issue(t, c_cur);
    if (1 < n) issue(t + 1, ahead(c_cur, 0, 1));
  }
  grid_dep_wait();  // x and y may still be written by the previous kernel of the stream
  VT carry = VT(0);
  bool open = false, head_red = false;
  uint32_t next_rank = 0;
  for (uint32_t i = 0; i < n; i++, t++) {
    const uint32_t st = t & 1u;
    mbar_wait(bars + st * 8, (t >> 1) & 1u);
    const uint32_t slot = ring + st * SLOT;
    const uint4 mraw = lds128(slot + CHUNK_BYTES);
    const uint4 iw = lds128(slot + 16u * (uint32_t)lane);
    uint2 hb;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(hb.x), "=r"(hb.y) : "r"(slot + kWideHiOff + 8u * (uint32_t)lane));
    const VT *xb = x + (size_t)(mraw.y & kMetaBlockMask) * cdb;
    VT xv[8];
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const uint32_t hi = ((s < 4 ? hb.x : hb.y) >> (8 * (s & 3))) & 0xFFu;
      uint32_t col = (hi << 15) | (idx16(iw, s) & 0x7FFFu);
      if (flags & 64u) col &= 0xFFFFu;  // diagnostic (diag_flags 64): every gather hits a 512 KB window, wrong results
#ifdef SPMVB_CHECK_BOUNDS
      const VT *p = x + SPMVB_BOUND(3, (uint32_t)((size_t)(mraw.y & kMetaBlockMask) * cdb + col), g_limits.x_len);
#else
      const VT *p = xb + col;
#endif
      xv[s] = x_policy ? ldg_x_hint(p, x_policy) : ldg_x(p);
    }
    // row ids of the chunk's row ends (lane + 32 j of them), in flight together with the gathers
    uint32_t rows_pre[8];
    if (!(mraw.z & kChunkRowsConsecutive)) {
      const uint32_t n_ends = mraw.y >> kMetaRowsShift;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t q = (uint32_t)lane + 32u * (uint32_t)j;
        rows_pre[j] = 0;
        if (q < n_ends) {
          const uint32_t rk = SPMVB_BOUND(1, mraw.x + q, g_limits.n_pairs);
          if (y_policy) asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(rows_pre[j]) : "l"(rowmap + rk), "l"(stream_policy));
          else asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(rows_pre[j]) : "l"(rowmap + rk));
        }
      }
    }
    uint4 vw[VW];
#pragma unroll
    for (int k = 0; k < VW; k++) vw[k] = lds128(slot + kWideValOff + (uint32_t)k * kWidePlane + 16u * (uint32_t)lane);
    const uint32_t pos = i & (R - 1);
    const bool sole = (mraw.z & kChunkSole) != 0 && !force_red;
    if (pos == 0) head_red = (mraw.z & kChunkStartsMid) != 0;
    process_chunk<VT, true, true>(iw, mraw, vw, xv, rowmap, y, lane, carry, open, next_rank, sole, head_red, y_policy,
                                  (flags & 32u) ? ~0ull : 0ull, scratch, rows_pre);
    if (pos == R - 1 || i + 1 == n) {  // the row left open continues in another warp's run: hand over atomically
      if (open && lane == 0) {
        uint32_t row = (mraw.z & kChunkRowsConsecutive) ? mraw.w + (next_rank - mraw.x)
                                                        : rowmap[SPMVB_BOUND(1, next_rank, g_limits.n_pairs)];
        row = SPMVB_BOUND(2, row, g_limits.rows);
        y_add(&y[row], carry);
      }
      carry = VT(0);
      open = false;
    }
    __syncwarp();
    if (lane == 0 && i + 2 < n) issue(t + 2, ahead(c_cur, i, 2));
    c_cur = ahead(c_cur, i, 1);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Variant ELL: the kernel of the sliced-ELLPACK image (ell.h) for regular matrices.  One warp per slice of 32
// consecutive rows, lane l owns row l: slot s of the 32 rows is ONE gather request over neighbouring columns (2-3
// distinct lines of x on a stencil instead of 12 with the reference's row-after-row groups), the row sum never leaves
// the lane's registers (left to right, multiply and add separately rounded like compute_results, spmv.cpp:84-97), and
// y is written once, 32 consecutive rows per store.  Slices arrive through a per-warp ring of STAGES TMA bulk copies
// (cp.async.bulk + mbarrier, evict-first); warp w of W takes slices w, w + W, ... so that all warps sweep the stream
// together.  flags bit 2: accumulate (y += A x; rows are exclusive, so a plain read-modify-write).
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

template <typename VT, int WARPS, int STAGES, int WMAX, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    spmv_ell_kernel(const uint8_t *__restrict__ image, const VT *__restrict__ x, VT *__restrict__ y, uint32_t rows,
                    uint32_t slice_begin, uint32_t n_slices, uint32_t width, uint32_t slice_bytes, uint32_t flags) {
  static_assert(STAGES >= 3, "slices i and i + 1 are both in use while the next ones arrive");
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t ring = smem_u32(smem) + (uint32_t)warp * STAGES * slice_bytes;
  const uint32_t bars = smem_u32(smem) + (uint32_t)WARPS * STAGES * slice_bytes + (uint32_t)warp * STAGES * 8;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) mbar_init(bars + s * 8, 1);
    fence_barrier_init();
  }
  __syncwarp();
  const uint32_t w = blockIdx.x * WARPS + warp, W = gridDim.x * WARPS;
  if (w >= n_slices) return;
  const uint32_t n = (n_slices - w + W - 1) / W;
  const uint64_t stream_policy = l2_policy_evict_first();
  auto issue = [&](uint32_t k) {  // lane 0 only: the warp's k-th slice into stage k % STAGES
    const uint32_t st = k % STAGES;
    const uint64_t slice = (uint64_t)slice_begin + w + (uint64_t)k * W;
    mbar_expect_tx(bars + st * 8, slice_bytes);
    bulk_g2s_hint(ring + st * slice_bytes, image + slice * slice_bytes, slice_bytes, bars + st * 8, stream_policy);
  };
  if (lane == 0)
    for (uint32_t k = 0; k < (uint32_t)STAGES && k < n; k++) issue(k);
  grid_dep_wait();  // x and y may still be written by the previous kernel of the stream
  // the warp's k-th slice has landed: its header, and the x gathers of this lane's row (one request per slot)
  auto fetch = [&](uint32_t k, uint4 &head, VT *xv) {
    const uint32_t st = k % STAGES;
    mbar_wait(bars + st * 8, (k / STAGES) & 1u);
    const uint32_t rec = ring + st * slice_bytes;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(head.x), "=r"(head.y), "=r"(head.z), "=r"(head.w) : "r"(rec));
#pragma unroll
    for (int s = 0; s < WMAX; s++)
      if ((uint32_t)s < width) {
        const uint32_t off = lds_u16(rec + 16u + (uint32_t)s * 64u + 2u * (uint32_t)lane);
        xv[s] = ldg_x(x + SPMVB_BOUND(3, head.x + off, g_limits.x_len));
      }
  };
  // software pipeline: the gathers of slice i + 1 are in flight while slice i is summed and stored
  uint4 head_n;
  VT xn[WMAX];
  fetch(0, head_n, xn);
  for (uint32_t i = 0; i < n; i++) {
    const uint4 head = head_n;
    VT xc[WMAX];
#pragma unroll
    for (int s = 0; s < WMAX; s++) xc[s] = xn[s];
    if (i + 1 < n) fetch(i + 1, head_n, xn);
    const uint32_t vals = ring + (i % STAGES) * slice_bytes + 16u + width * 64u + (uint32_t)lane * (uint32_t)sizeof(VT);
    VT acc = VT(0);
#pragma unroll
    for (int s = 0; s < WMAX; s++)
      if ((uint32_t)s < width) acc = vadd(acc, vmul(lds_v(vals + (uint32_t)s * 32u * (uint32_t)sizeof(VT), VT(0)), xc[s]));
    const uint32_t row = head.z + (uint32_t)lane;
    if (row < rows) {
      const uint32_t rr = SPMVB_BOUND(2, row, g_limits.rows);
      y[rr] = (flags & 4u) ? vadd(y[rr], acc) : acc;
    }
    __syncwarp();  // every lane's loads from this stage have been consumed: it can be refilled
    if (lane == 0 && i + STAGES < n) issue(i + STAGES);
  }
}

// y[rows[i]] = 0 for the rows that receive atomics or no update at all (Layout::zero_rows)
template <typename VT>
__global__ void zero_rows_kernel(VT *__restrict__ y, const uint32_t *__restrict__ rows, uint32_t n) {
  grid_dep_launch();  // the SpMV kernel may start its prologue now; it waits for this grid before touching y
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

// dst[i] = src[i] / sqrt(*sumsq): the normalisation of the iterated caller with the norm still on the device
template <typename VT>
__global__ void scale_rsqrt_kernel(const VT *__restrict__ src, VT *__restrict__ dst, uint32_t n,
                                   const double *__restrict__ sumsq) {
  const double ss = *sumsq;
  const VT s = (VT)(ss > 0.0 ? 1.0 / sqrt(ss) : 0.0);
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

// ---- conjugate gradients (the iterated caller on the Laplacian): vector work of one iteration in
// three kernels around the SpMV.  Scalars stay on the device: s[0], s[2] = r.r of the current / next iteration (the
// host swaps their roles every iteration), s[1] = p.q.  Sums are accumulated in double.
__device__ __forceinline__ void block_sum_to(double acc, double *out) {
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

template <typename VT>
__global__ void dot_kernel(const VT *__restrict__ a, const VT *__restrict__ b, uint32_t n, double *__restrict__ out) {
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc += (double)a[i] * (double)b[i];
  block_sum_to(acc, out);
}

// x += alpha p; r -= alpha q; *rr_next += r.r   with alpha = *rr / *pq
template <typename VT>
__global__ void cg_update_kernel(VT *__restrict__ x, VT *__restrict__ r, const VT *__restrict__ p,
                                 const VT *__restrict__ q, uint32_t n, const double *__restrict__ rr,
                                 const double *__restrict__ pq, double *__restrict__ rr_next) {
  const double d = *pq;
  const VT alpha = (VT)(d != 0.0 ? *rr / d : 0.0);
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const VT ri = r[i] - alpha * q[i];
    r[i] = ri;
    acc += (double)ri * (double)ri;
  }
  block_sum_to(acc, rr_next);
}

// p = r + beta p   with beta = *rr_next / *rr
template <typename VT>
__global__ void cg_direction_kernel(VT *__restrict__ p, const VT *__restrict__ r, uint32_t n,
                                    const double *__restrict__ rr, const double *__restrict__ rr_next) {
  const double d = *rr;
  const VT beta = (VT)(d != 0.0 ? *rr_next / d : 0.0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = r[i] + beta * p[i];
}

__global__ void l2_flush_kernel(uint4 *__restrict__ buf, size_t n_words) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x)
    buf[i] = make_uint4((uint32_t)i, 0, 0, 0);
}

}  // namespace spmvb
