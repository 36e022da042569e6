// FP64 throughput on B200: DFMA (CUDA cores) vs DMMA (mma.sync.m8n8k4.f64) vs both interleaved.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu ; run: ./dmma_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
  for (int k = 0; k < NACC; ++k) c[k][0] = c[k][1] = 0.0;
  double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) dmma(c[k][0], c[k][1], a, b);
  }
  double s = 0;
  for (int k = 0; k < NACC; ++k) s += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
  double c[NACC];
  for (int k = 0; k < NACC; ++k) c[k] = k;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NACC; ++k) c[k] = fma(c[k], a, b);
  }
  double s = 0;
  for (int k = 0; k < NACC; ++k) s += c[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// per iteration: NM dmma + NF dfma, independent chains
template <int NM, int NF>
__global__ void k_mix(double* out, int iters, double a0, double b0) {
  double c[NM][2], f[NF];
  for (int k = 0; k < NM; ++k) c[k][0] = c[k][1] = 0.0;
  for (int k = 0; k < NF; ++k) f[k] = k;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NM; ++k) dmma(c[k][0], c[k][1], a, b);
#pragma unroll
    for (int k = 0; k < NF; ++k) f[k] = fma(f[k], a, b);
  }
  double s = 0;
  for (int k = 0; k < NM; ++k) s += c[k][0] + c[k][1];
  for (int k = 0; k < NF; ++k) s += f[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
  const int iters = 20000;
  for (int wps : {4, 8, 16, 32}) {  // warps per SM
    const int threads = 128, blocks = sms * wps / 4;
    const double nthr = (double)blocks * threads;
    float ms = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("warps/SM %2d  DFMA x8 chains : %7.2f TFLOP/s\n", wps, 2.0 * nthr * iters * 8 / ms / 1e9);
    ms = timeit([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0, 1.0); });
    printf("warps/SM %2d  DMMA x8 chains : %7.2f TFLOP/s  (%.2f cycles per DMMA per SMSP at %d MHz)\n", wps,
           2.0 * (nthr / 32) * iters * 8 * 256 / ms / 1e9, ms * 1e-3 * p.clockRate * 1e3 / ((double)iters * 8 * wps / 4), p.clockRate / 1000);
    ms = timeit([&] { k_dmma<2><<<blocks, threads>>>(out, iters, 1.0, 1.0); });
    printf("warps/SM %2d  DMMA x2 chains : %7.2f TFLOP/s\n", wps, 2.0 * (nthr / 32) * iters * 2 * 256 / ms / 1e9);
    ms = timeit([&] { k_mix<4, 8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("warps/SM %2d  mix 4 DMMA + 8 DFMA : %7.2f TFLOP/s total (DMMA part %.2f, DFMA part %.2f)\n", wps,
           (2.0 * (nthr / 32) * iters * 4 * 256 + 2.0 * nthr * iters * 8) / ms / 1e9, 2.0 * (nthr / 32) * iters * 4 * 256 / ms / 1e9,
           2.0 * nthr * iters * 8 / ms / 1e9);
  }
  return 0;
}
