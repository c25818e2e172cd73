// FP64 peak probe for B200 (sm_100a): DMMA.8x8x4 issue rate vs DFMA issue rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5 + threadIdx.x * 1e-9;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(a, c[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s SMs %d clock %d kHz\n", prop.name, sms, prop.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 4096;
  for (int warps_per_sm = 4; warps_per_sm <= 32; warps_per_sm *= 2) {
    int threads = 256; int blocks_per_sm = warps_per_sm * 32 / threads; if (blocks_per_sm < 1) { threads = warps_per_sm * 32; blocks_per_sm = 1; }
    int grid = sms * blocks_per_sm;
    float ms;
    // DMMA, 16 independent accumulators per warp
    dmma_loop<16><<<grid, threads>>>(out, 64, 1.0); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); dmma_loop<16><<<grid, threads>>>(out, iters, 1.0); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 256 * 16 * (double)iters * (grid * (threads / 32));
    printf("DMMA.8x8x4 nacc=16 warps/SM=%2d : %8.2f TFLOP/s  (%.3f ms)\n", warps_per_sm, flops / ms / 1e9, ms);
    dmma_loop<4><<<grid, threads>>>(out, 64, 1.0); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); dmma_loop<4><<<grid, threads>>>(out, iters * 4, 1.0); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    flops = 2.0 * 256 * 4 * (double)iters * 4 * (grid * (threads / 32));
    printf("DMMA.8x8x4 nacc= 4 warps/SM=%2d : %8.2f TFLOP/s  (%.3f ms)\n", warps_per_sm, flops / ms / 1e9, ms);
    // DFMA
    dfma_loop<16><<<grid, threads>>>(out, 64, 1.0); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); dfma_loop<16><<<grid, threads>>>(out, iters * 4, 1.0); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    flops = 2.0 * 32 * 16 * (double)iters * 4 * (grid * (threads / 32));
    printf("DFMA       nacc=16 warps/SM=%2d : %8.2f TFLOP/s  (%.3f ms)\n", warps_per_sm, flops / ms / 1e9, ms);
  }
  // single-warp DMMA latency probe: dependent chain
  {
    float ms;
    CK(cudaEventRecord(e0)); dmma_loop<1><<<1, 32>>>(out, 1 << 16, 1.0); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("DMMA dependent chain: %.2f ns per DMMA (1 warp)\n", ms * 1e6 / (1 << 16));
    CK(cudaEventRecord(e0)); dmma_loop<16><<<1, 32>>>(out, 1 << 14, 1.0); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("DMMA 16 indep, 1 warp: %.2f ns per DMMA\n", ms * 1e6 / ((1 << 14) * 16));
    CK(cudaEventRecord(e0)); dmma_loop<16><<<1, 128>>>(out, 1 << 14, 1.0); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("DMMA 16 indep, 4 warps (1/SMSP): %.2f ns per DMMA per warp\n", ms * 1e6 / ((1 << 14) * 16));
  }
  return 0;
}
