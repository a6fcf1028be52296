// Rectangular-Hessenberg (:HessQR) path kernels.
//
// Reference: /root/reference/src/compute.jl:571-592 -- per grid point, shift the diagonal of the (n+1) x n
// Hessenberg matrix (573), sweep n-1 Givens rotations over row pairs (581-587), keep triu(T1[1:n,1:n]) (592),
// then run the inverse-Lanczos recurrence on that private triangular factor (602).  Every shift owns its own
// R, so this path is BLAS-2 and bound by memory bandwidth (HBM / L2), not by the tensor pipe.
#pragma once
#include <type_traits>

#include "psa_common.cuh"

namespace psa {

constexpr int HQ_MAX_N = 1024;   // one thread per column in the QR sweep
constexpr int HS_THREADS = 128;  // solve kernel (small CTAs: more slots in flight per SM)
constexpr int HB = 32;           // diagonal block of the per-shift triangular solves

// Per-slot storage of the private triangular factor: Rc (column-major n x n), Rr (row-major copy, so that BOTH
// triangular solves stream contiguous vectors) and rinv (reciprocal diagonal).
inline size_t hess_R_elems(int n) { return 2 * (size_t)n * n + (size_t)((n + 31) / 32 * 32); }
// Slots in flight: enough CTAs to keep HBM busy and to amortise the per-tick launches, bounded by a 24 GB budget
// for the private triangular factors.
inline long long hess_slot_capacity(int n, int num_sms) {
  const long long by_mem = (24LL << 30) / ((long long)hess_R_elems(n) * 16 + 1);
  long long w = 64LL * num_sms;
  if (w > by_mem) w = by_mem;
  if (n > 256) {
    // the warp-per-slot streaming kernels for large n keep 8 (n <= 512) / 4 (n <= 1024) warps resident per SM: whole
    // waves only, a partly filled last wave costs a full one (each warp's solve is a latency-bound chain of 2n steps)
    const long long wave = (n <= 512 ? 8LL : 4LL) * num_sms;
    if (w >= wave) return w / wave * wave;
  }
  return w < 64 ? 64 : w / 64 * 64;
}

// zlartg-convention Givens (Julia `givens(f,g,1,2)`, compute.jl:585): [c s; -conj(s) c][f; g] = [r; 0], c real
__device__ __forceinline__ void givens_cs(double2 f, double2 g, double& c, double2& s) {
  if (g.x == 0.0 && g.y == 0.0) {
    c = 1.0;
    s = make_double2(0.0, 0.0);
    return;
  }
  // The rotation is on the critical path of the sweep (one thread computes it, everybody waits), so no hypot() here:
  // an exact power-of-two scaling to the unit range guards the squares against overflow and underflow, then
  // c = |f| / d and s = (f / |f|) conj(g) / d with d = sqrt(|f|^2 + |g|^2) come from two reciprocal square roots.
  const double mx = fmax(fmax(fabs(f.x), fabs(f.y)), fmax(fabs(g.x), fabs(g.y)));
  int ex;
  (void)frexp(mx, &ex);
  const double sc = ldexp(1.0, -max(ex, -1000));  // (subnormal inputs: stay finite)
  f.x *= sc;
  f.y *= sc;
  g.x *= sc;
  g.y *= sc;
  const double ag2 = fma(g.x, g.x, g.y * g.y);
  const double af2 = fma(f.x, f.x, f.y * f.y);
  if (af2 == 0.0) {
    const double ia = rsqrt(ag2);
    c = 0.0;
    s = make_double2(g.x * ia, -g.y * ia);
    return;
  }
  const double iaf = rsqrt(af2), id = rsqrt(af2 + ag2);
  c = af2 * iaf * id;
  const double2 ph = make_double2(f.x * iaf, f.y * iaf);
  const double2 cg = make_double2(g.x * id, -g.y * id);
  s = cmul(ph, cg);
}

// One CTA per freshly scheduled slot, one thread per column c.  The thread keeps the running entry of row
// jj ("carry") in a register: rotation jj turns (carry, T1[jj+1,c]) into the final R[jj,c] and the next carry.
// Column jj's owner produces the rotation; it is broadcast through (double-buffered) shared memory.
// full_sweep = 0: :HessQR, rotations jj = 1..n-1 only (the reference's loop bound, compute.jl:581).
// full_sweep = 1: :rect_qr on Hessenberg input: the reference calls qr(T1) (compute.jl:589), which also
//   eliminates H[n+1,n]; for an upper Hessenberg matrix that is one more rotation, and R is unique up to unit row
//   phases, which leave (R^H R) and hence sigma unchanged.
__global__ void __launch_bounds__(HQ_MAX_N) hess_qr_kernel(const int* __restrict__ fresh, const int* __restrict__ n_fresh, int m, int n,
                               const double2* __restrict__ H, const double2* __restrict__ zs,
                               double2* __restrict__ Rall, int full_sweep) {
  __shared__ double g_c[2];
  __shared__ double2 g_s[2];
  const int c = threadIdx.x;
  // grid-stride over the fresh slots: the launch is sized from an upper bound (the host may be a few ticks behind)
  for (int f = blockIdx.x; f < *n_fresh; f += gridDim.x) {
  __syncthreads();  // g_c / g_s of the previous slot are no longer read
  const int slot = fresh[f];
  const double2 z = zs[slot];
  const size_t rstride = 2 * (size_t)n * n + (size_t)((n + 31) / 32 * 32);
  double2* R = Rall + (size_t)slot * rstride;   // Rc: R[i,k] at k*n + i
  double2* Rr = R + (size_t)n * n;               // Rr: R[j,k] at j*n + k
  double2* rinv = Rr + (size_t)n * n;            // 1 / R[k,k]
  const bool live = c < n;
  const double2* Hc = H + (size_t)(live ? c : 0) * m;
  double2 carry = make_double2(0.0, 0.0);
  if (live) {
    carry = Hc[0];
    if (c == 0) {
      carry.x -= z.x;
      carry.y -= z.y;
    }
  }
  // H[jj+1, c] is fetched one rotation ahead (a strided read of the shared H: L2 latency off the critical path)
  double2 bnext = (live && n > 1) ? Hc[1] : make_double2(0.0, 0.0);
  for (int jj = 0; jj < n - 1; ++jj) {
    double2 b = make_double2(0.0, 0.0);
    if (live && c >= jj) {
      b = bnext;
      if (jj + 1 == c) {
        b.x -= z.x;
        b.y -= z.y;
      }
    }
    if (live && c >= jj + 1 && jj + 2 < m) bnext = Hc[jj + 2];
    if (c == jj) {
      double cs;
      double2 sn;
      givens_cs(carry, b, cs, sn);
      g_c[jj & 1] = cs;
      g_s[jj & 1] = sn;
    }
    __syncthreads();
    if (live && c >= jj) {
      const double cs = g_c[jj & 1];
      const double2 sn = g_s[jj & 1];
      const double2 a = carry;
      // R[jj,c] = cs*a + sn*b ; carry = -conj(sn)*a + cs*b
      double2 r = cmul(sn, b);
      r.x = fma(cs, a.x, r.x);
      r.y = fma(cs, a.y, r.y);
      R[(size_t)c * n + jj] = r;
      Rr[(size_t)jj * n + c] = r;
      if (c == jj) rinv[jj] = crecip(r);
      const double2 sc = make_double2(-sn.x, sn.y);  // -conj(sn)
      double2 nc = cmul(sc, a);
      nc.x = fma(cs, b.x, nc.x);
      nc.y = fma(cs, b.y, nc.y);
      carry = nc;
    }
  }
  if (c == n - 1) {
    if (full_sweep) {
      const double2 b = Hc[n];  // row n+1 of the (n+1) x n matrix
      double cs;
      double2 sn;
      givens_cs(carry, b, cs, sn);
      double2 r = cmul(sn, b);
      r.x = fma(cs, carry.x, r.x);
      r.y = fma(cs, carry.y, r.y);
      carry = r;
    }
    R[(size_t)c * n + (n - 1)] = carry;  // HessQR: H[n+1,n] is never rotated in (reference quirk)
    Rr[(size_t)(n - 1) * n + c] = carry;
    rinv[n - 1] = crecip(carry);
  }
  }
}

// One CTA per active slot: w = R^{-H} q (forward, dot form) then v = R^{-1} w (backward, axpy form); v -> Wv.
struct HessSolveParams {
  int n, n_pad;
  const int* active;
  const int* n_active;
  const double2* R;
  const double2* Q0;
  const double2* Q1;
  double2* Wv;
  const int* sel;
  const int* iter;     // 0 = fresh slot: q is still the start vector of its row batch (read in place)
  const int* batch;
  const double2* q0;   // [nbatch][n_pad]
};

__device__ __forceinline__ const double2* hess_q_ptr(const HessSolveParams& P, int slot) {
  return P.iter[slot] == 0 ? P.q0 + (size_t)P.batch[slot] * P.n_pad
                           : (P.sel[slot] ? P.Q1 : P.Q0) + (size_t)slot * P.n_pad;
}

__global__ void __launch_bounds__(HS_THREADS) hess_solve_kernel(HessSolveParams P) {
  extern __shared__ __align__(16) unsigned char hsm[];
  double2* w = reinterpret_cast<double2*>(hsm);                       // n_pad
  double2* tri = w + P.n_pad;                                         // HB x (HB+1), tri[col*(HB+1) + row]
  double2* part = tri + HB * (HB + 1);                                // HB
  if ((int)blockIdx.x >= *P.n_active) return;
  const int slot = P.active[blockIdx.x];
  const int n = P.n, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double2* R = P.R + (size_t)slot * (2 * (size_t)n * n + (size_t)((n + 31) / 32 * 32));
  const double2* q = hess_q_ptr(P, slot);
  const int nb = (n + HB - 1) / HB;

  // ---- forward: (R^H) w = q ----
  for (int J = 0; J < nb; ++J) {
    const int j0 = J * HB, jw = min(HB, n - j0);
    // diagonal block -> smem (coalesced down each column)
    for (int e = tid; e < HB * HB; e += HS_THREADS) {
      const int col = e >> 5, row = e & 31;
      tri[col * (HB + 1) + row] = (col < jw && row <= col) ? R[(size_t)(j0 + col) * n + j0 + row] : make_double2(0.0, 0.0);
    }
    // partial dots against the already known w[0:j0]
    for (int jj = warp; jj < jw; jj += HS_THREADS / 32) {
      const double2* col = R + (size_t)(j0 + jj) * n;
      double sr = 0.0, si = 0.0;
      for (int i = lane; i < j0; i += 32) {
        const double2 r = col[i], x = w[i];
        sr = fma(r.x, x.x, sr);
        sr = fma(r.y, x.y, sr);   // conj(r)*x
        si = fma(r.x, x.y, si);
        si = fma(-r.y, x.x, si);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
      }
      if (lane == 0) part[jj] = make_double2(sr, si);
    }
    __syncthreads();
    if (warp == 0) {
      double2 val = make_double2(0.0, 0.0);
      if (lane < jw) {
        const double2 qq = q[j0 + lane], pp = part[lane];
        val = make_double2(qq.x - pp.x, qq.y - pp.y);
      }
      for (int i = 0; i < jw; ++i) {
        if (lane == i) {
          double2 d = tri[i * (HB + 1) + i];
          d.y = -d.y;
          val = cmul(val, crecip(d));
        }
        const double wx = __shfl_sync(0xffffffffu, val.x, i), wy = __shfl_sync(0xffffffffu, val.y, i);
        if (lane > i && lane < jw) {
          double2 r = tri[lane * (HB + 1) + i];
          r.y = -r.y;
          cmsub(val, r, make_double2(wx, wy));
        }
      }
      if (lane < jw) w[j0 + lane] = val;
    }
    __syncthreads();
  }
  // ---- backward: R v = w (in place in smem) ----
  for (int J = nb - 1; J >= 0; --J) {
    const int j0 = J * HB, jw = min(HB, n - j0);
    for (int e = tid; e < HB * HB; e += HS_THREADS) {
      const int col = e >> 5, row = e & 31;
      tri[col * (HB + 1) + row] = (col < jw && row <= col) ? R[(size_t)(j0 + col) * n + j0 + row] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    if (warp == 0) {
      double2 val = lane < jw ? w[j0 + lane] : make_double2(0.0, 0.0);
      for (int k = jw - 1; k >= 0; --k) {
        if (lane == k) val = cmul(val, crecip(tri[k * (HB + 1) + k]));
        const double vx = __shfl_sync(0xffffffffu, val.x, k), vy = __shfl_sync(0xffffffffu, val.y, k);
        if (lane < k) cmsub(val, tri[k * (HB + 1) + lane], make_double2(vx, vy));
      }
      if (lane < jw) w[j0 + lane] = val;
    }
    __syncthreads();
    // rows above the block: w[i] -= sum_k R[i, j0+k] v[j0+k]
    for (int i = tid; i < j0; i += HS_THREADS) {
      double2 acc = w[i];
      for (int k = 0; k < jw; ++k) cmsub(acc, R[(size_t)(j0 + k) * n + i], w[j0 + k]);
      w[i] = acc;
    }
    __syncthreads();
  }
  double2* out = P.Wv + (size_t)slot * P.n_pad;
  for (int i = tid; i < n; i += HS_THREADS) out[i] = w[i];
}

// Warp-per-slot variant for n <= 256: the vector lives in registers (8 elements per lane), both solves are
// "axpy form" (one contiguous row of Rr / column of Rc per step), the pivot value is broadcast with a shuffle and
// the diagonal reciprocals come from the QR kernel.  The rows stream through a per-warp ring in shared memory
// filled by cp.async (LDGSTS) HW_DEPTH steps ahead, so a step costs a fraction of a memory latency and ~16 warps
// per SM keep HBM busy; there is no block-level synchronisation at all.
constexpr int HW_NPL = 8;        // vector elements per lane
constexpr int HW_THREADS = 128;  // 4 slots per CTA
constexpr int HW_DEPTH = 4;      // rows in flight per warp

inline size_t hess_warp_smem(int n) { return (size_t)(HW_THREADS / 32) * HW_DEPTH * ((n + 31) / 32 * 32) * sizeof(double2); }

__global__ void __launch_bounds__(HW_THREADS) hess_solve_warp_kernel(HessSolveParams P) {
  extern __shared__ __align__(16) unsigned char hwsm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = blockIdx.x * (HW_THREADS / 32) + warp;
  if (a >= *P.n_active) return;
  const int slot = P.active[a];
  const int n = P.n, n32 = (n + 31) / 32 * 32;
  const size_t rstride = 2 * (size_t)n * n + (size_t)n32;
  const double2* Rc = P.R + (size_t)slot * rstride;
  const double2* Rr = Rc + (size_t)n * n;
  const double2* rinv = Rr + (size_t)n * n;
  const double2* q = hess_q_ptr(P, slot);
  double2* ring = reinterpret_cast<double2*>(hwsm) + (size_t)warp * HW_DEPTH * n32;
  const double2 zero = make_double2(0.0, 0.0);

  double2 x[HW_NPL], ri[HW_NPL];
#pragma unroll
  for (int e = 0; e < HW_NPL; ++e) {
    const int i = 32 * e + lane;
    x[e] = i < n ? q[i] : zero;
    ri[e] = i < n ? rinv[i] : zero;
  }
  // ---- forward: (R^H) w = q.  After w_j is final, row j of R updates the entries k > j. ----
  auto fetch_row = [&](int j) {  // entries k > j of row j -> ring slot j % DEPTH (always commits a group)
    if (j < n) {
      double2* dst = ring + (size_t)(j % HW_DEPTH) * n32;
#pragma unroll
      for (int e = 0; e < HW_NPL; ++e) {
        const int k = 32 * e + lane;
        if (k > j && k < n) cp_async16(dst + k, Rr + (size_t)j * n + k);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int d = 0; d < HW_DEPTH; ++d) fetch_row(d);
#pragma unroll
  for (int je = 0; je < HW_NPL; ++je) {
    for (int jl = 0; jl < 32; ++jl) {
      const int j = 32 * je + jl;
      if (j >= n) break;
      cp_async_wait<HW_DEPTH - 1>();
      __syncwarp();
      const double2* row = ring + (size_t)(j % HW_DEPTH) * n32;
      const double xr = __shfl_sync(0xffffffffu, x[je].x, jl), xi = __shfl_sync(0xffffffffu, x[je].y, jl);
      const double rr = __shfl_sync(0xffffffffu, ri[je].x, jl), rim = __shfl_sync(0xffffffffu, ri[je].y, jl);
      const double2 wj = cmul(make_double2(xr, xi), make_double2(rr, -rim));  // / conj(R[j,j])
      if (lane == jl) x[je] = wj;
#pragma unroll
      for (int e = 0; e < HW_NPL; ++e)
        if (e >= je) {
          const int k = 32 * e + lane;
          if (k > j && k < n) {
            const double2 r = row[k];
            cmsub(x[e], make_double2(r.x, -r.y), wj);  // conj(R[j,k]) * w_j
          }
        }
      __syncwarp();
      fetch_row(j + HW_DEPTH);
    }
  }
  cp_async_wait<0>();
  __syncwarp();
  // ---- backward: R v = w.  After v_k is final, column k of R updates the entries i < k. ----
  auto fetch_col = [&](int k) {  // entries i < k of column k -> ring slot k % DEPTH
    if (k >= 0) {
      double2* dst = ring + (size_t)(k % HW_DEPTH) * n32;
#pragma unroll
      for (int e = 0; e < HW_NPL; ++e) {
        const int i = 32 * e + lane;
        if (i < k) cp_async16(dst + i, Rc + (size_t)k * n + i);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int d = 0; d < HW_DEPTH; ++d) fetch_col(n - 1 - d);
#pragma unroll
  for (int ke = HW_NPL - 1; ke >= 0; --ke) {
    for (int kl = 31; kl >= 0; --kl) {
      const int k = 32 * ke + kl;
      if (k >= n) continue;
      cp_async_wait<HW_DEPTH - 1>();
      __syncwarp();
      const double2* col = ring + (size_t)(k % HW_DEPTH) * n32;
      const double xr = __shfl_sync(0xffffffffu, x[ke].x, kl), xi = __shfl_sync(0xffffffffu, x[ke].y, kl);
      const double rr = __shfl_sync(0xffffffffu, ri[ke].x, kl), rim = __shfl_sync(0xffffffffu, ri[ke].y, kl);
      const double2 vk = cmul(make_double2(xr, xi), make_double2(rr, rim));
      if (lane == kl) x[ke] = vk;
#pragma unroll
      for (int e = 0; e < HW_NPL; ++e)
        if (e <= ke) {
          const int i = 32 * e + lane;
          if (i < k) cmsub(x[e], col[i], vk);
        }
      __syncwarp();
      fetch_col(k - HW_DEPTH);
    }
  }
  cp_async_wait<0>();
  double2* out = P.Wv + (size_t)slot * P.n_pad;
#pragma unroll
  for (int e = 0; e < HW_NPL; ++e) {
    const int i = 32 * e + lane;
    if (i < n) out[i] = x[e];
  }
}

// Warp-per-slot variant for 256 < n <= 1024 (NPL = 16: n <= 512, NPL = 32: n <= 1024).  Same streaming scheme as
// hess_solve_warp_kernel -- the vector in registers, rows of Rr / columns of Rc through a per-warp cp.async ring, no
// block-level synchronisation -- restructured so that the code does not grow with NPL^2 and the register file holds
// 32 elements per lane: the 32-element block that contains the pivots of the current 32 steps lives in one named
// register pair (xe); the reciprocal diagonal travels with its row through the ring instead of sitting in registers;
// the block loop is a run-time loop, only the update over the NPL blocks is unrolled (predicated on block > pivot block).
// A step is bound by the latency of its row (fetched DEPTH - 1 steps ahead), and the rows of a triangular factor get
// shorter: the ring is re-cut in LEVELS -- once at most 1/2, 1/4, 1/8 of the blocks remain, the same shared memory holds
// 6, 12, 24 rows in flight instead of 3 (one drain of the ring per level).
// Same arithmetic per element, in the same order, as the small-n kernel.
template <int NPL>
struct HessBig {
  static constexpr int THREADS = 128;               // 4 slots per CTA
  static constexpr int DEPTH0 = 3;                  // rows in flight per warp at level 0
  static constexpr int LEVELS = 4;
  static int ring_elems(int n) { return DEPTH0 * ((n + 31) / 32 * 32) + 32; }
  static size_t smem(int n) { return (size_t)(THREADS / 32) * ring_elems(n) * sizeof(double2); }
};

template <int NPL>
__global__ void __launch_bounds__(HessBig<NPL>::THREADS) hess_solve_big_kernel(HessSolveParams P) {
  using HBK = HessBig<NPL>;
  extern __shared__ __align__(16) unsigned char hbsm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int a = blockIdx.x * (HBK::THREADS / 32) + warp;
  if (a >= *P.n_active) return;
  const int slot = P.active[a];
  const int n = P.n, n32 = (n + 31) / 32 * 32, nblk = n32 / 32;
  const size_t rstride = 2 * (size_t)n * n + (size_t)n32;
  const double2* Rc = P.R + (size_t)slot * rstride;
  const double2* Rr = Rc + (size_t)n * n;
  const double2* rinv = Rr + (size_t)n * n;
  const double2* q = hess_q_ptr(P, slot);
  double2* ring = reinterpret_cast<double2*>(hbsm) + (size_t)warp * (HBK::DEPTH0 * n32 + 32);
  const double2 zero = make_double2(0.0, 0.0);

  double2 x[NPL];
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int i = 32 * e + lane;
    x[e] = i < n ? q[i] : zero;
  }
  auto get_block = [&](int b) {
    double2 v = zero;
#pragma unroll
    for (int e = 0; e < NPL; ++e)
      if (e == b) v = x[e];
    return v;
  };
  auto put_block = [&](int b, double2 v) {
#pragma unroll
    for (int e = 0; e < NPL; ++e)
      if (e == b) x[e] = v;
  };
  // level of a phase position with `rem` blocks left to stream: the deepest L with rem <= nblk >> L (and a non-empty cut)
  auto level_of = [&](int rem) {
    int L = 0;
    while (L + 1 < HBK::LEVELS && (nblk >> (L + 1)) >= 1 && rem <= (nblk >> (L + 1))) ++L;
    return L;
  };

  // ---- forward: (R^H) w = q.  After w_j is final, row j of R updates the entries k > j. ----
  // One pivot block (32 rows) at ring geometry (DEPTH rows of `rs` elements; element k of a row at index k - 32 * kb).
  auto fwd_block = [&](auto depth_tag, int je, int rs, int kb, bool prime) {
    constexpr int DEPTH = decltype(depth_tag)::value;
    auto fetch_row = [&](int j) {  // entries k > j of row j (+ rinv[j]) -> ring slot j % DEPTH (always commits a group)
      if (j < n) {
        double2* dst = ring + (size_t)(j % DEPTH) * rs;
        const int e0 = j >> 5;
#pragma unroll
        for (int e = 0; e < NPL; ++e) {
          const int k = 32 * e + lane;
          if (e >= e0 && k > j && k < n) cp_async16(dst + (k - 32 * kb), Rr + (size_t)j * n + k);
        }
        if (lane == 0) cp_async16(dst + rs - 1, rinv + j);
      }
      cp_async_commit();
    };
    if (prime) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) fetch_row(32 * je + d);
    }
    double2 xe = get_block(je);
    const int jend = min(32, n - 32 * je);
    for (int jl = 0; jl < jend; ++jl) {
      const int j = 32 * je + jl;
      cp_async_wait<DEPTH - 1>();
      __syncwarp();
      const double2* row = ring + (size_t)(j % DEPTH) * rs - 32 * kb;
      const double2 rj = row[32 * kb + rs - 1];
      const double xr = __shfl_sync(0xffffffffu, xe.x, jl), xi = __shfl_sync(0xffffffffu, xe.y, jl);
      const double2 wj = cmul(make_double2(xr, xi), make_double2(rj.x, -rj.y));  // / conj(R[j,j])
      if (lane == jl) xe = wj;
      {
        const int k = 32 * je + lane;
        if (k > j && k < n) {
          const double2 r = row[k];
          cmsub(xe, make_double2(r.x, -r.y), wj);  // conj(R[j,k]) * w_j
        }
      }
#pragma unroll
      for (int e = 0; e < NPL; ++e)
        if (e > je) {
          const int k = 32 * e + lane;
          if (k < n) {
            const double2 r = row[k];
            cmsub(x[e], make_double2(r.x, -r.y), wj);
          }
        }
      __syncwarp();
      fetch_row(j + DEPTH);
    }
    put_block(je, xe);
  };
  {
    int cur = -1;
    for (int je = 0; je < nblk; ++je) {
      const int L = level_of(nblk - je);
      const bool prime = L != cur;
      if (prime && cur >= 0) {  // re-cut the ring: everything in flight belongs to the old geometry
        cp_async_wait<0>();
        __syncwarp();
      }
      cur = L;
      const int blocks = L == 0 ? nblk : (nblk >> L);  // blocks a row of this level can span
      const int rs = 32 * blocks + 1, kb = nblk - blocks;
      switch (L) {
        case 0: fwd_block(std::integral_constant<int, HBK::DEPTH0>{}, je, rs, kb, prime); break;
        case 1: fwd_block(std::integral_constant<int, 2 * HBK::DEPTH0>{}, je, rs, kb, prime); break;
        case 2: fwd_block(std::integral_constant<int, 4 * HBK::DEPTH0>{}, je, rs, kb, prime); break;
        default: fwd_block(std::integral_constant<int, 8 * HBK::DEPTH0>{}, je, rs, kb, prime); break;
      }
    }
  }
  cp_async_wait<0>();
  __syncwarp();
  // ---- backward: R v = w.  After v_k is final, column k of R updates the entries i < k. ----
  auto bwd_block = [&](auto depth_tag, int ke, int rs, bool prime) {
    constexpr int DEPTH = decltype(depth_tag)::value;
    auto fetch_col = [&](int k) {  // entries i < k of column k (+ rinv[k]) -> ring slot k % DEPTH
      if (k >= 0) {
        double2* dst = ring + (size_t)(k % DEPTH) * rs;
        const int e1 = k >> 5;
#pragma unroll
        for (int e = 0; e < NPL; ++e) {
          const int i = 32 * e + lane;
          if (e <= e1 && i < k) cp_async16(dst + i, Rc + (size_t)k * n + i);
        }
        if (lane == 0) cp_async16(dst + rs - 1, rinv + k);
      }
      cp_async_commit();
    };
    const int ktop = min(32 * ke + 31, n - 1);
    if (prime) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) fetch_col(ktop - d);
    }
    double2 xe = get_block(ke);
    for (int k = ktop; k >= 32 * ke; --k) {
      const int kl = k & 31;
      cp_async_wait<DEPTH - 1>();
      __syncwarp();
      const double2* col = ring + (size_t)(k % DEPTH) * rs;
      const double2 rk = col[rs - 1];
      const double xr = __shfl_sync(0xffffffffu, xe.x, kl), xi = __shfl_sync(0xffffffffu, xe.y, kl);
      const double2 vk = cmul(make_double2(xr, xi), rk);
      if (lane == kl) xe = vk;
      if (lane < kl) cmsub(xe, col[32 * ke + lane], vk);
#pragma unroll
      for (int e = 0; e < NPL; ++e)
        if (e < ke) cmsub(x[e], col[32 * e + lane], vk);
      __syncwarp();
      fetch_col(k - DEPTH);
    }
    put_block(ke, xe);
  };
  {
    int cur = -1;
    for (int ke = nblk - 1; ke >= 0; --ke) {
      const int L = level_of(ke + 1);
      const bool prime = L != cur;
      if (prime && cur >= 0) {
        cp_async_wait<0>();
        __syncwarp();
      }
      cur = L;
      const int blocks = L == 0 ? nblk : (nblk >> L);
      const int rs = 32 * blocks + 1;
      switch (L) {
        case 0: bwd_block(std::integral_constant<int, HBK::DEPTH0>{}, ke, rs, prime); break;
        case 1: bwd_block(std::integral_constant<int, 2 * HBK::DEPTH0>{}, ke, rs, prime); break;
        case 2: bwd_block(std::integral_constant<int, 4 * HBK::DEPTH0>{}, ke, rs, prime); break;
        default: bwd_block(std::integral_constant<int, 8 * HBK::DEPTH0>{}, ke, rs, prime); break;
      }
    }
  }
  cp_async_wait<0>();
  double2* out = P.Wv + (size_t)slot * P.n_pad;
#pragma unroll
  for (int e = 0; e < NPL; ++e) {
    const int i = 32 * e + lane;
    if (i < n) out[i] = x[e];
  }
}

struct HessTick {
  int m, n, n_pad;
  const double2* H;
  double2* R;
  const double2* q0;
  const int* fresh;
  int n_fresh;              // upper bound on the number of fresh slots (exact count: *n_fresh_dev)
  const int* n_fresh_dev;
  const int* active;
  int n_active;             // upper bound (exact count: *n_active_dev)
  const int* n_active_dev;
  double2 *Q0, *Q1, *Wv;
  const int* sel;
  const int* iter;
  const int* batch;
  const double2* z;
  int full_sweep;
  int num_sms;
};

inline size_t hess_solve_smem(int n_pad) { return ((size_t)n_pad + HB * (HB + 1) + HB) * sizeof(double2); }

inline int hess_set_attrs() {
  cudaError_t e = cudaFuncSetAttribute(hess_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(hess_solve_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hess_warp_smem(32 * HW_NPL));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(hess_solve_big_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HessBig<16>::smem(512));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(hess_solve_big_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HessBig<32>::smem(1024));
  return e == cudaSuccess ? 0 : -1;
}

// launches the tick's HessQR kernels; the caller recorded the start event, ev_b is recorded here. Returns #launches.
inline int hess_launch_tick(const HessTick& t, cudaStream_t stream, cudaEvent_t ev_b) {
  int launches = 0;
  if (t.n_fresh > 0) {
    const int threads = (t.n + 31) / 32 * 32;
    const int grid = t.n_fresh < 16 * t.num_sms ? t.n_fresh : 16 * t.num_sms;
    hess_qr_kernel<<<grid, threads, 0, stream>>>(t.fresh, t.n_fresh_dev, t.m, t.n, t.H, t.z, t.R, t.full_sweep);
    ++launches;
  }
  HessSolveParams sp;
  sp.n = t.n;
  sp.n_pad = t.n_pad;
  sp.active = t.active;
  sp.n_active = t.n_active_dev;
  sp.R = t.R;
  sp.Q0 = t.Q0;
  sp.Q1 = t.Q1;
  sp.Wv = t.Wv;
  sp.sel = t.sel;
  sp.iter = t.iter;
  sp.batch = t.batch;
  sp.q0 = t.q0;
  if (t.n <= 32 * HW_NPL)
    hess_solve_warp_kernel<<<(t.n_active + HW_THREADS / 32 - 1) / (HW_THREADS / 32), HW_THREADS, hess_warp_smem(t.n), stream>>>(sp);
  else if (getenv("PSA_HESS_CTA"))  // experiments: the CTA-per-slot kernel (A/B timing)
    hess_solve_kernel<<<t.n_active, HS_THREADS, hess_solve_smem(t.n_pad), stream>>>(sp);
  else if (t.n <= 512)
    hess_solve_big_kernel<16><<<(t.n_active + 3) / 4, HessBig<16>::THREADS, HessBig<16>::smem(t.n), stream>>>(sp);
  else
    hess_solve_big_kernel<32><<<(t.n_active + 3) / 4, HessBig<32>::THREADS, HessBig<32>::smem(t.n), stream>>>(sp);
  ++launches;
  cudaEventRecord(ev_b, stream);
  return launches;
}

}  // namespace psa
