// Micro-benchmark: FP64 add-reduction into global memory through the TMA engine
// (cp.reduce.async.bulk.global.shared::cta ... .add.f64) against per-lane RED.ADD.F64, for the scatter of
// short contiguous runs (a 6-dof node run of a CSC column is 48 B).  Each issuing lane sends one run of SIZE
// bytes from the warp's shared staging area to a pseudo-random 16 B-aligned place of a large array.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_red_bench bulk_red_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_red_add_f64(double* dst, const double* src_smem, int bytes) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(src_smem);
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(s), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// SIZE bytes per op, NL issuing lanes per warp and round; every round rewrites the staging area (as a real
// emitter would) and waits until the previous round's reads of it are done.
template <int SIZE, int NL>
__global__ void k_bulk(double* out, size_t n, int iters) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  constexpr int ND = SIZE / 8;
  double* st = sm + (size_t)wib * 32 * ND;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (int it = 0; it < iters; ++it) {
    bulk_wait_read0();
    __syncwarp();
    for (int k = 0; k < ND; ++k) st[k * 32 + lane] = 1.0;  // staging writes (conflict-free)
    fence_async_smem();
    __syncwarp();
    if (lane < NL) {
      size_t w = warp + (size_t)it * nwarps;
      size_t base = ((w * 2654435761ull + lane * 40503ull) % (n / 64)) * 64;  // 512 B granules
      bulk_red_add_f64(out + base + 2 * (lane & 3), st + lane * ND, SIZE);
    }
    bulk_commit();
  }
  bulk_wait0();
}

template <int ND>  // the same runs with per-lane RED: ND consecutive doubles per run, 32 / ND runs... (lane -> element)
__global__ void k_red(double* out, size_t n, int iters) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (int it = 0; it < iters; ++it) {
    size_t w = warp + (size_t)it * nwarps;
    const int run = lane / ND, k = lane % ND;
    if (run * ND + ND <= 32) {
      size_t base = ((w * 2654435761ull + run * 40503ull) % (n / 64)) * 64;
      atomicAdd(out + base + k, 1.0);
    }
  }
}

template <int SIZE, int NL>
void run_bulk(double* d, size_t n) {
  const int iters = 128, block = 128;
  const int grid = 148 * 8;
  const size_t smem = (size_t)(block / 32) * 32 * (SIZE / 8) * 8;
  cudaFuncSetAttribute(k_bulk<SIZE, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k_bulk<SIZE, NL><<<grid, block, smem>>>(d, n, 4);
  cudaEventRecord(a);
  k_bulk<SIZE, NL><<<grid, block, smem>>>(d, n, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)grid * (block / 32) * NL * iters;
  cudaError_t e = cudaGetLastError();
  printf("  bulk add.f64 %3d B x %2d lanes/warp: %7.2f G ops/s  %7.1f G adds/s  %6.2f SM-cycles/op  (%s)\n", SIZE, NL, ops / ms / 1e6,
         ops * (SIZE / 8) / ms / 1e6, ms * 1e-3 * 1.965e9 * 148 / ops, cudaGetErrorString(e));
}
template <int ND>
void run_red(double* d, size_t n) {
  const int iters = 128;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k_red<ND><<<148 * 16, 256>>>(d, n, 4);
  cudaEventRecord(a);
  k_red<ND><<<148 * 16, 256>>>(d, n, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double adds = 148.0 * 16 * 8 * (32 / ND * ND) * iters;
  printf("  RED.ADD.F64 runs of %d: %7.1f G adds/s\n", ND, adds / ms / 1e6);
}
int main() {
  size_t n = (size_t)3 * 1024 * 1024 * 1024 / 8;
  double* d;
  cudaMalloc(&d, n * 8);
  cudaMemset(d, 0, n * 8);
  run_red<6>(d, n);
  run_red<1>(d, n);
  run_bulk<16, 32>(d, n);
  run_bulk<32, 32>(d, n);
  run_bulk<48, 32>(d, n);
  run_bulk<48, 16>(d, n);
  run_bulk<48, 8>(d, n);
  run_bulk<96, 32>(d, n);
  run_bulk<96, 16>(d, n);
  run_bulk<192, 16>(d, n);
  run_bulk<288, 8>(d, n);
  // correctness: total of the array = number of adds issued (every add contributes 1.0)
  cudaDeviceSynchronize();
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
