// Shared device helpers for libpsagpu (sm_100a only): mbarrier / bulk-copy (TMA engine) PTX wrappers,
// FP64 tensor-core MMA, small complex helpers, deterministic block reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace psa {

// ---- async-proxy plumbing -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time; test_wait never does)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 16-byte asynchronous copy global -> shared (SASS: LDGSTS.E.BYPASS.128), per-thread addresses, and its completion
// hooked onto an mbarrier (the arrival was pre-counted in the barrier's init count: ".noinc")
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// L2 eviction-priority policies for the two streams of the solve kernel: the packed T tiles are shared by every CTA
// (keep them: evict_last), the X rows are private to one CTA and re-read only a whole block row later (evict_first)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void cp_async16_hint(void* smem_dst, const void* gmem_src, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
               "l"(policy)
               : "memory");
}
// the same with the shared-memory destination given as a 32-bit shared-window address
__device__ __forceinline__ void cp_async16_hint_s(uint32_t smem_dst, const void* gmem_src, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "l"(policy)
               : "memory");
}
// order generic-proxy writes (st.global / st.shared) before later async-proxy accesses (bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- FP64 tensor core: D(8x8) += A(8x4, row) * B(4x8, col)   (SASS: DMMA.8x8x4) ---------------------
// lane l holds A[l>>2][l&3], B[l&3][l>>2], C[l>>2][2*(l&3)] and C[l>>2][2*(l&3)+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- complex helpers (double2 = (re, im)) ---------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// The diagonal-block solves of the square path exist in several choreographies that must agree bit for bit: their
// pivot arithmetic is spelled with explicit roundings, so that the compiler's choice of which product of
// a.x*b.x - a.y*b.y to fuse cannot differ between them.
__device__ __forceinline__ double2 cmul_fixed(double2 a, double2 b) {
  return make_double2(__fma_rn(a.x, b.x, -__dmul_rn(a.y, b.y)), __fma_rn(a.x, b.y, __dmul_rn(a.y, b.x)));
}
// 1 / d for the pivot d = t_ii - z
__device__ __forceinline__ double2 cinv_fixed(double2 d) {
  const double inv = __drcp_rn(__fma_rn(d.x, d.x, __dmul_rn(d.y, d.y)));
  return make_double2(__dmul_rn(d.x, inv), -__dmul_rn(d.y, inv));
}
// a -= b * c
__device__ __forceinline__ void cmsub(double2& a, double2 b, double2 c) {
  a.x = fma(-b.x, c.x, a.x);
  a.x = fma(b.y, c.y, a.x);
  a.y = fma(-b.x, c.y, a.y);
  a.y = fma(-b.y, c.x, a.y);
}
__device__ __forceinline__ double2 crecip(double2 d) {
  double inv = 1.0 / (d.x * d.x + d.y * d.y);
  return make_double2(d.x * inv, -d.y * inv);
}

// ---- deterministic block reduction (fixed tree; result broadcast to all threads) --------------------
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* red /* >= THREADS/32 doubles of smem */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect red[] from a previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) t += red[w];
  return t;
}
template <int THREADS>
__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = red[0];
#pragma unroll
  for (int w = 1; w < THREADS / 32; ++w) t = fmax(t, red[w]);
  return t;
}

}  // namespace psa
