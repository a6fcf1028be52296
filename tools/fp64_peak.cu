// FP64 roofline denominators for B200 (sm_100a): raw DFMA-pipe and DMMA-pipe issue peaks.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
// Not part of the product path; results are recorded in profiles/ and DESIGN.md.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-9 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int TILES>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
  double c0[TILES], c1[TILES];
#pragma unroll
  for (int t = 0; t < TILES; ++t) { c0[t] = threadIdx.x * 1e-9; c1[t] = t; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < TILES; ++t) dmma(c0[t], c1[t], a, b);
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < TILES; ++t) s += c0[t] + c1[t];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms %d\n", p.name, sms);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 1 << 16;
  for (int bps = 1; bps <= 8; bps *= 2) {
    for (int threads = 128; threads <= 256; threads *= 2) {
      int grid = sms * bps;
      float ms = time_ms([&] { dfma_kernel<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double flops = 2.0 * 8 * iters * (double)grid * threads;
      printf("DFMA chains=8 blocks/SM=%d threads=%d : %.3f ms  %.2f TFLOP/s\n", bps, threads, ms, flops / ms * 1e-9);
      ms = time_ms([&] { dmma_kernel<8><<<grid, threads>>>(out, iters / 8, 1.0000001, 1e-9); }, 5);
      flops = 512.0 * 8 * (iters / 8) * (double)grid * (threads / 32);
      printf("DMMA tiles=8  blocks/SM=%d threads=%d : %.3f ms  %.2f TFLOP/s\n", bps, threads, ms, flops / ms * 1e-9);
      ms = time_ms([&] { dmma_kernel<16><<<grid, threads>>>(out, iters / 8, 1.0000001, 1e-9); }, 5);
      flops = 512.0 * 16 * (iters / 8) * (double)grid * (threads / 32);
      printf("DMMA tiles=16 blocks/SM=%d threads=%d : %.3f ms  %.2f TFLOP/s\n", bps, threads, ms, flops / ms * 1e-9);
    }
  }
  // sustained: ~3 s of DMMA back to back
  {
    int grid = sms * 2;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    for (; n < 400; ++n) dmma_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 512.0 * 16 * iters * (double)grid * 8 * n;
    printf("DMMA sustained %d launches: %.1f ms  %.2f TFLOP/s\n", n, ms, flops / ms * 1e-9);
  }
  return 0;
}
