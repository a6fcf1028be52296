// DMMA issue-rate ceiling for the solve kernel's instruction mix (complex 16x32 warp tile per k4 step:
// 2 A fragments x 4 B fragments x 4 real DMMAs), with and without the LDS.128 fragment loads.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE, int NT>
__global__ void __launch_bounds__(256, 2) mix_kernel(double* out, int iters, const double2* gsrc) {
  extern __shared__ double2 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = gsrc[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double cre[2][NT][2], cim[2][NT][2];
  for (int a = 0; a < 2; ++a) for (int b = 0; b < NT; ++b) for (int e = 0; e < 2; ++e) { cre[a][b][e] = 0; cim[a][b][e] = 0; }
  double2 a[2], b[NT];
  for (int m = 0; m < 2; ++m) a[m] = sm[m * 32 + lane];
  for (int t = 0; t < NT; ++t) b[t] = sm[256 + t * 32 + lane];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (MODE == 2) {
#pragma unroll
        for (int m = 0; m < 2; ++m) a[m] = sm[((j * 8 + 2 * (warp & 3) + m) * 32 + lane + it * 64) & 4095];
#pragma unroll
        for (int t = 0; t < NT; ++t) b[t] = sm[(1024 + (t * 8 + (lane >> 2)) * 20 + ((4 * j + (lane & 3)) ^ (it & 15)) + it * 16) & 4095];
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int t = 0; t < NT; ++t) dmma(cre[m][t][0], cre[m][t][1], a[m].x, b[t].x);
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int t = 0; t < NT; ++t) dmma(cim[m][t][0], cim[m][t][1], a[m].x, b[t].y);
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int t = 0; t < NT; ++t) dmma(cre[m][t][0], cre[m][t][1], -a[m].y, b[t].y);
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int t = 0; t < NT; ++t) dmma(cim[m][t][0], cim[m][t][1], a[m].y, b[t].x);
      if (MODE == 1) {  // perturb operands in registers so that they differ per step without memory traffic
        a[0].x += 1e-9; a[1].y -= 1e-9; b[0].x += 1e-9;
      }
    }
  }
  double s = 0;
  for (int m = 0; m < 2; ++m) for (int t = 0; t < NT; ++t) for (int e = 0; e < 2; ++e) s += cre[m][t][e] + cim[m][t][e];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int NT>
void run(const char* name, double* out, double2* src, int grid, int threads) {
  const int iters = 20000;
  CK(cudaFuncSetAttribute(mix_kernel<MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  mix_kernel<MODE, NT><<<grid, threads, 65536>>>(out, 100, src);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0));
    mix_kernel<MODE, NT><<<grid, threads, 65536>>>(out, iters, src);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  double flops = 512.0 * (2 * NT * 4) * 4 * iters * (double)grid * (threads / 32);
  printf("%-44s grid=%d threads=%d NT=%d : %8.3f ms %6.2f TFLOP/s\n", name, grid, threads, NT, best, flops / best * 1e-9);
}
int main() {
  double* out; double2* src;
  CK(cudaMalloc(&out, 8 << 20)); CK(cudaMalloc(&src, 4096 * 16)); CK(cudaMemset(src, 0, 4096 * 16));
  for (int bps = 1; bps <= 2; ++bps) {
    run<0, 4>("fixed operands (registers)", out, src, 148 * bps, 256);
    run<1, 4>("varying operands (registers)", out, src, 148 * bps, 256);
    run<2, 4>("operands via LDS.128 each k4 step", out, src, 148 * bps, 256);
    run<2, 2>("operands via LDS.128, NT=2", out, src, 148 * bps, 256);
    run<2, 1>("operands via LDS.128, NT=1", out, src, 148 * bps, 256);
  }
  return 0;
}
