// :SVD path -- the reference's branch for small matrices (n < minlancs4psa = 55):
//   Twork[diag] = diaga - zpt;  Z[j,k] = minimum(svd(Twork).S)          /root/reference/src/compute.jl:510-518
// (threaded variant _SVD_Server, compute.jl:363-386).  One CTA per grid point: the shifted matrix lives in shared
// memory and is diagonalised by one-sided (Hestenes) Jacobi rotations on column pairs, all disjoint pairs of a
// round in parallel (one warp per pair).  One-sided Jacobi delivers the small singular values to high relative
// accuracy, which is what the reference's LAPACK zgesdd is asked for here.
#pragma once
#include "psa_common.cuh"

namespace psa {

constexpr long long SVD_MAX_ELEMS = 12288;  // m*n complex128 elements that fit one CTA's shared memory (192 KiB)
constexpr int SVD_THREADS = 512;
constexpr int SVD_MAX_SWEEPS = 40;

inline size_t svd_smem(int m, int n) { return (size_t)m * n * sizeof(double2) + 16 + (SVD_THREADS / 32) * sizeof(double); }

struct SvdParams {
  int m, n;
  const double2* T;  // m x n column-major, ld = m (shared by all points)
  int P, rows_local, row0, row_step;
  const double* x;
  const double* y;
  double* Z;    // local point order
  int* iters;   // nullable: set to 0 (no Lanczos iterations on this branch)
  unsigned long long* counts;
};

__global__ void __launch_bounds__(SVD_THREADS) svd_kernel(SvdParams S) {
  extern __shared__ __align__(16) unsigned char ssm[];
  double2* A = reinterpret_cast<double2*>(ssm);
  int* flag = reinterpret_cast<int*>(ssm + (size_t)S.m * S.n * sizeof(double2));
  double* smin = reinterpret_cast<double*>(flag + 2);
  const int m = S.m, n = S.n, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = SVD_THREADS / 32;
  const int N = (n + 1) & ~1;  // players of the round-robin tournament (one dummy when n is odd)
  const double eps = 2.220446049250313e-16;

  for (int p = blockIdx.x; p < S.P; p += gridDim.x) {
    const int jr = p % S.rows_local, k = p / S.rows_local;
    const double zx = S.x[k], zy = S.y[S.row0 + jr * S.row_step];
    __syncthreads();
    for (int e = tid; e < m * n; e += SVD_THREADS) {
      double2 v = S.T[e];
      const int i = e % m, j = e / m;
      if (i == j) {
        v.x -= zx;
        v.y -= zy;
      }
      A[e] = v;
    }
    __syncthreads();
    for (int sweep = 0; sweep < SVD_MAX_SWEEPS; ++sweep) {
      if (tid == 0) flag[0] = 0;
      __syncthreads();
      for (int r = 0; r < N - 1; ++r) {
        for (int pr = warp; pr < N / 2; pr += NW) {
          int a_, b_;
          if (pr == 0) {
            a_ = N - 1;
            b_ = r;
          } else {
            a_ = (r + pr) % (N - 1);
            b_ = (r - pr + (N - 1)) % (N - 1);
          }
          const int cp = min(a_, b_), cq = max(a_, b_);
          if (cq >= n) continue;  // pair with the dummy player
          double2* Ap = A + (size_t)cp * m;
          double2* Aq = A + (size_t)cq * m;
          double sa = 0.0, sb = 0.0, cr = 0.0, ci = 0.0;
          for (int i = lane; i < m; i += 32) {
            const double2 u = Ap[i], v = Aq[i];
            sa = fma(u.x, u.x, fma(u.y, u.y, sa));
            sb = fma(v.x, v.x, fma(v.y, v.y, sb));
            cr = fma(u.x, v.x, fma(u.y, v.y, cr));   // conj(u) * v
            ci = fma(u.x, v.y, fma(-u.y, v.x, ci));
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
            cr += __shfl_xor_sync(0xffffffffu, cr, o);
            ci += __shfl_xor_sync(0xffffffffu, ci, o);
          }
          const double cabs = hypot(cr, ci);
          if (cabs > eps * sqrt(sa) * sqrt(sb) && cabs > 0.0) {
            if (lane == 0) flag[0] = 1;
            // phase-align q (q' = e^{-i phi} q so that p^H q' = |c|), then a real Jacobi rotation
            const double phr = cr / cabs, phi = ci / cabs;
            const double zeta = (sb - sa) / (2.0 * cabs);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
            for (int i = lane; i < m; i += 32) {
              const double2 u = Ap[i], v = Aq[i];
              const double2 vq = make_double2(v.x * phr + v.y * phi, v.y * phr - v.x * phi);  // e^{-i phi} v
              Ap[i] = make_double2(cs * u.x - sn * vq.x, cs * u.y - sn * vq.y);
              Aq[i] = make_double2(sn * u.x + cs * vq.x, sn * u.y + cs * vq.y);
            }
          }
        }
        __syncthreads();
      }
      if (flag[0] == 0) break;
      __syncthreads();
    }
    // singular values = column norms; keep the smallest
    double best = 1e308;
    for (int c = warp; c < n; c += NW) {
      const double2* Ac = A + (size_t)c * m;
      double mx = 0.0;
      for (int i = lane; i < m; i += 32) mx = fmax(mx, fmax(fabs(Ac[i].x), fabs(Ac[i].y)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      double s = 0.0;
      if (mx > 0.0) {
        for (int i = lane; i < m; i += 32) {
          const double a = Ac[i].x / mx, b = Ac[i].y / mx;
          s = fma(a, a, fma(b, b, s));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      }
      best = fmin(best, mx * sqrt(s));
    }
    __syncthreads();
    if (lane == 0) smin[warp] = best;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < NW; ++w) best = fmin(best, smin[w]);
      S.Z[p] = best;
      if (S.iters) S.iters[p] = 0;
      atomicAdd(&S.counts[3], 1ull);
    }
  }
}

}  // namespace psa
