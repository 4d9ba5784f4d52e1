// Random-access rates of one B200, measured: the second roofline of an irregular SpMV.  Every non-zero of an irregular
// matrix costs ONE scattered 8-byte access that cannot be made local (either the x gather or the y update; see
// DESIGN.md section 3.4), so next to HBM bytes/s the kernel is bounded by scattered accesses/s:
//   red   red.global.add.f64 to uniformly random rows of an M-element vector (the fused accum_results of an irregular
//         matrix: one update per (row, block) pair)
//   ldg   ld.global.nc.f64 gathers from random elements of an M-element vector (x through L1/L2)
//   lds   ld.shared.f64 gathers from a 128 KB window (x staged in shared memory)
// for M from L2-resident to HBM-resident sizes.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a (Makefile
// target `tools`); run on the GPU box: spmv-fpga_b200/lib/access_probe > profiles/r2/access_probe.txt
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <int MODE>  // 0 red, 1 ldg, 2 red with evict-last policy
__global__ void __launch_bounds__(256) probe(double *__restrict__ v, uint64_t m_mask, uint64_t per_thread, double *sink) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t pol = 0;
  if (MODE == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  double acc = 0.0;
  uint64_t k = mix64(tid);
  for (uint64_t i = 0; i < per_thread; i += 8) {
    uint64_t idx[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { k = mix64(k + i + j); idx[j] = k & m_mask; }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (MODE == 0) atomicAdd(v + idx[j], 1.0);
      else if (MODE == 2) asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(v + idx[j]), "d"(1.0), "l"(pol) : "memory");
      else { double t; asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(t) : "l"(v + idx[j])); acc += t; }
    }
  }
  if (acc == 12345.678) *sink = acc;
}

__global__ void __launch_bounds__(512) probe_lds(const double *__restrict__ x, uint64_t per_thread, double *sink) {
  extern __shared__ double win[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) win[i] = x[i];
  __syncthreads();
  double acc = 0.0;
  uint64_t k = mix64((uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
  for (uint64_t i = 0; i < per_thread; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; j++) { k = mix64(k + i + j); acc += win[k & 16383]; }
  }
  if (acc == 12345.678) *sink = acc;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t max_elems = (size_t)1 << 27;  // 1 GiB of doubles
  double *v, *sink;
  cudaMalloc(&v, max_elems * 8);
  cudaMalloc(&sink, 8);
  cudaMemset(v, 0, max_elems * 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = sms * 8, block = 256;
  const uint64_t per_thread = 1024;
  const double total = (double)grid * block * per_thread;
  printf("# B200 random 8-byte access rates, %d SMs, %d x %d threads x %llu accesses (uniformly random addresses)\n", sms, grid, block,
         (unsigned long long)per_thread);
  printf("%-28s %12s %14s\n", "vector size", "mode", "G accesses/s");
  for (int lg = 20; lg <= 27; lg++) {
    const uint64_t mask = ((uint64_t)1 << lg) - 1;
    for (int mode = 0; mode < 3; mode++) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a);
        if (mode == 0) probe<0><<<grid, block>>>(v, mask, per_thread, sink);
        else if (mode == 1) probe<1><<<grid, block>>>(v, mask, per_thread, sink);
        else probe<2><<<grid, block>>>(v, mask, per_thread, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
      }
      printf("%4llu MB (2^%d doubles)       %12s %14.1f\n", (unsigned long long)(8ull << lg >> 20), lg,
             mode == 0 ? "red.f64" : mode == 1 ? "ldg.nc.f64" : "red.f64+evl", total / best / 1e6);
    }
  }
  cudaFuncSetAttribute(probe_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(a);
    probe_lds<<<sms, 512, 131072>>>(v, per_thread * 4, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  printf("%-28s %12s %14.1f\n", "128 KB window (shared)", "lds.f64", (double)sms * 512 * per_thread * 4 / best / 1e6);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
