// CUDA side of the GPU layout builder: the backend that runs the steps of layout_gpu_steps.h as kernels on a stream
// (grid-stride launches, CUB for the scans, the stable radix sort by column block and the zero-row compaction).
// Included by engine.cu only.
#pragma once
#include <cuda_runtime.h>
#include <omp.h>

#include <cstdio>
#include <cstdlib>

#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "../../include/spmvb.h"
#include "layout_gpu_steps.h"

namespace spmvb {

template <class Body>
__global__ void __launch_bounds__(256) lb_kernel(uint64_t n, LbCtx c) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    Body::run(i, c);
}

struct LbBit0 {
  __host__ __device__ __forceinline__ uint32_t operator()(uint8_t v) const { return v & 1u; }
};
struct LbCastU32 {
  __host__ __device__ __forceinline__ uint32_t operator()(uint8_t v) const { return v; }
};
struct LbCastU64 {
  __host__ __device__ __forceinline__ uint64_t operator()(uint32_t v) const { return v; }
};
struct LbNonZero {
  __host__ __device__ __forceinline__ unsigned long long operator()(uint8_t v) const { return v != 0; }
};

struct CudaBackend {
  cudaStream_t st = nullptr;
  int sms = 148;
  cudaError_t e = cudaSuccess;
  std::string where;
  void *scratch = nullptr;  // CUB temporary storage (arena memory), grown on demand
  size_t scratch_bytes = 0;
  unsigned long long *d_count = nullptr;
  // Temporaries are carved out of a few large device allocations (dozens of cudaMalloc / cudaFree calls cost more
  // than all the build kernels together); they go back in one piece when the backend is destroyed.
  std::vector<void *> arenas;
  uint8_t *cur = nullptr;
  size_t cur_left = 0;

  ~CudaBackend() {
    for (void *a : arenas) cudaFree(a);  // cudaFree waits for the work that still uses the memory
  }
  void chk(cudaError_t r, const char *what) {
    if (e == cudaSuccess && r != cudaSuccess) { e = r; where = what; }
  }
  bool ok() const { return e == cudaSuccess; }
  int code() const { return SPMVB_E_CUDA; }
  std::string error() const { return "GPU layout build: " + where + ": " + cudaGetErrorString(e); }

  // option build_trace = 1: wait for the stream after every stage and print the time it took (diagnostics)
  bool tracing = options().build_trace > 0;
  double t_last = omp_get_wtime();
  void trace(const char *what) {
    if (!tracing) return;
    cudaStreamSynchronize(st);
    const double t = omp_get_wtime();
    fprintf(stderr, "[layout build] %-34s %8.3f ms\n", what, (t - t_last) * 1e3);
    t_last = t;
  }
  void reserve(uint64_t bytes) {
    if (!ok()) return;
    void *p = nullptr;
    chk(cudaMalloc(&p, (size_t)bytes), "cudaMalloc (arena)");
    if (!ok()) return;
    arenas.push_back(p);
    cur = (uint8_t *)p; cur_left = (size_t)bytes;
  }
  void *alloc(size_t bytes) {
    bytes = ((bytes < 16 ? 16 : bytes) + 255) & ~(size_t)255;
    if (bytes > cur_left) reserve(bytes > ((size_t)8 << 20) ? bytes : ((size_t)8 << 20));
    if (!ok()) return nullptr;
    void *p = cur;
    cur += bytes; cur_left -= bytes;
    return p;
  }
  void release(void *) {}
  void *alloc_output(size_t bytes) {
    void *p = nullptr;
    if (ok()) chk(cudaMalloc(&p, bytes < 16 ? 16 : bytes), "cudaMalloc");
    return p;
  }
  void release_output(void *p) { if (p) cudaFree(p); }
  void fill(void *p, int byte, size_t bytes) { if (ok() && bytes) chk(cudaMemsetAsync(p, byte, bytes, st), "memset"); }
  void to_host(void *dst, const void *src, size_t bytes) {
    if (!ok() || !bytes) return;
    chk(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st), "copy to host");
    chk(cudaStreamSynchronize(st), "kernels before a copy to host");
  }
  void to_device(void *dst, const void *src, size_t bytes) {
    if (ok() && bytes) chk(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st), "copy to device");
  }
  void to_host_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height) {
    if (!ok() || !width || !height) return;
    chk(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, st), "2-D copy to host");
    chk(cudaStreamSynchronize(st), "2-D copy to host");
  }
  template <class Body>
  void launch(uint64_t n, const LbCtx &c) {
    if (!ok() || n == 0) return;
    const uint64_t want = (n + 255) / 256, cap = (uint64_t)sms * 16;
    lb_kernel<Body><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(n, c);
    chk(cudaGetLastError(), "kernel launch");
  }
  bool temp(size_t bytes) {
    if (bytes <= scratch_bytes) return true;
    scratch = alloc(bytes);
    scratch_bytes = ok() ? bytes : 0;
    return ok();
  }
  void inclusive_sum_bit0_u32(const uint8_t *in, uint32_t *out, uint64_t n) {
    if (!ok() || !n) return;
    cub::TransformInputIterator<uint32_t, LbBit0, const uint8_t *> it(in, LbBit0());
    size_t bytes = 0;
    chk(cub::DeviceScan::InclusiveSum(nullptr, bytes, it, out, (int)n, st), "scan size");
    if (temp(bytes)) chk(cub::DeviceScan::InclusiveSum(scratch, bytes, it, out, (int)n, st), "inclusive scan");
  }
  void exclusive_sum_u32_u64(const uint32_t *in, uint64_t *out, uint64_t n) {
    if (!ok() || !n) return;
    cub::TransformInputIterator<uint64_t, LbCastU64, const uint32_t *> it(in, LbCastU64());
    size_t bytes = 0;
    chk(cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, (int)n, st), "scan size");
    if (temp(bytes)) chk(cub::DeviceScan::ExclusiveSum(scratch, bytes, it, out, (int)n, st), "exclusive scan");
  }
  void exclusive_sum_u8_u32(const uint8_t *in, uint32_t *out, uint64_t n) {
    if (!ok() || !n) return;
    cub::TransformInputIterator<uint32_t, LbCastU32, const uint8_t *> it(in, LbCastU32());
    size_t bytes = 0;
    chk(cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, (int)n, st), "scan size");
    if (temp(bytes)) chk(cub::DeviceScan::ExclusiveSum(scratch, bytes, it, out, (int)n, st), "exclusive scan");
  }
  void sort_pairs(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, uint64_t n, int bits) {
    if (!ok() || !n) return;
    size_t bytes = 0;
    chk(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, bits, st), "sort size");
    if (temp(bytes)) chk(cub::DeviceRadixSort::SortPairs(scratch, bytes, kin, kout, vin, vout, (int)n, 0, bits, st), "radix sort");
  }
  void sort_pairs_u64(const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout, uint64_t n, int bits) {
    if (!ok() || !n) return;
    size_t bytes = 0;
    chk(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, bits, st), "sort size");
    if (temp(bytes)) chk(cub::DeviceRadixSort::SortPairs(scratch, bytes, kin, kout, vin, vout, (int)n, 0, bits, st), "radix sort (entries)");
  }
  bool counter() {
    if (!d_count) d_count = (unsigned long long *)alloc(16);
    return ok();
  }
  uint64_t count_nonzero_u8(const uint8_t *in, uint64_t n) {
    if (!ok() || !n || !counter()) return 0;
    cub::TransformInputIterator<unsigned long long, LbNonZero, const uint8_t *> it(in, LbNonZero());
    size_t bytes = 0;
    chk(cub::DeviceReduce::Sum(nullptr, bytes, it, d_count, (int)n, st), "reduce size");
    if (temp(bytes)) chk(cub::DeviceReduce::Sum(scratch, bytes, it, d_count, (int)n, st), "reduce");
    unsigned long long h = 0;
    to_host(&h, d_count, 8);
    return h;
  }
  uint64_t select_flagged_iota(const uint8_t *flags, uint32_t *out, uint64_t n) {
    if (!ok() || !n || !counter()) return 0;
    cub::CountingInputIterator<uint32_t> iota(0);
    unsigned long long *d_n = d_count + 1;
    size_t bytes = 0;
    chk(cub::DeviceSelect::Flagged(nullptr, bytes, iota, flags, out, d_n, (int)n, st), "select size");
    if (temp(bytes)) chk(cub::DeviceSelect::Flagged(scratch, bytes, iota, flags, out, d_n, (int)n, st), "select");
    unsigned long long h = 0;
    to_host(&h, d_n, 8);
    return h;
  }
};

}  // namespace spmvb
