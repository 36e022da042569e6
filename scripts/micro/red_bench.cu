// Micro-benchmark: throughput of RED.ADD.F64 as a function of how many lanes of a warp
// instruction fall into the same 32 B sector (1 = scattered like the per-lane block emission,
// 4 = fully coalesced runs).  Build: nvcc -arch=sm_100a -O3 -o red_bench red_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int GROUP>  // GROUP consecutive lanes write consecutive doubles, groups are scattered
__global__ void k_red(double* out, size_t n, int iters) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
    // each (warp, it) touches a pseudo-random region; within it, lane groups are contiguous
    size_t w = (tid >> 5) + (size_t)it * (nthreads >> 5);
    size_t base = (w * 2654435761ull) % (n / 1024) * 1024;
    size_t off = (size_t)(lane / GROUP) * 24 + (lane % GROUP);  // groups 24 doubles (6 sectors) apart
    atomicAdd(out + base + off, 1.0);
  }
}
template <int G>
float run(double* d, size_t n, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k_red<G><<<148 * 16, 256>>>(d, n, 8);
  cudaEventRecord(a);
  k_red<G><<<148 * 16, 256>>>(d, n, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}
int main() {
  for (size_t mb : {64, 4096}) {
    size_t n = mb * 1024 * 1024 / 8;
    double* d;
    cudaMalloc(&d, n * 8);
    cudaMemset(d, 0, n * 8);
    const int iters = 256;
    const double ops = 148.0 * 16 * 256 * iters;
    printf("array %zu MB\n", mb);
    printf("  group 1 (32 sectors/instr): %.1f G RED/s\n", ops / run<1>(d, n, iters) / 1e6);
    printf("  group 2                   : %.1f G RED/s\n", ops / run<2>(d, n, iters) / 1e6);
    printf("  group 4 (8 sectors/instr) : %.1f G RED/s\n", ops / run<4>(d, n, iters) / 1e6);
    printf("  group 6 (runs of 6)       : %.1f G RED/s\n", ops / run<6>(d, n, iters) / 1e6);
    printf("  group 32 (contiguous)     : %.1f G RED/s\n", ops / run<32>(d, n, iters) / 1e6);
    cudaFree(d);
  }
  return 0;
}
