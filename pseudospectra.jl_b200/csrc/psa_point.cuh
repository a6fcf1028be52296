// Point-resident Lanczos for SMALL square matrices (minlancs4psa <= n <= ~160): the whole of
//   `_psa_lanczos!`  (/root/reference/src/compute.jl:619-665)
// for one grid point runs inside ONE warp of a persistent kernel -- both triangular solves of every iteration
// (`v = F1 \ (F1' \ q)`, compute.jl:630), the recurrence (631-638), the tridiagonal eigenvalue (640-642) and the
// stopping test (658-661) -- so a grid needs one launch instead of one tick (3 launches) per Lanczos iteration.
//
// Why a second kernel for the same branch: for n = 150 (the README example) a whole 100 x 100 grid is 3.6e10 flop,
// 1 ms at the FP64 rate, but the tick engine needs max(iterations) ~ 45 ticks, and a tick cannot be shorter than the
// dependent chain of diagonal-block solves in the DMMA kernel (~110 us at n = 150).  Here nothing is batched across
// points, so nothing waits for the slowest point:
//   * the packed upper triangle of T (n(n+1)/2 complex128, 181 KB at n = 150) is staged ONCE per CTA into shared
//     memory and shared by all of its warps (one CTA per SM, up to 16 warps = 16 points in flight per SM);
//   * a warp keeps the right-hand side of the sweeps in registers (element i in lane i % 32, register i / 32) and solves in
//     column-sweep (axpy) form: the owner lane of x_j multiplies by the pre-inverted pivot 1 / (t_jj - z), one shuffle
//     broadcasts it, every lane updates its part of the right-hand side.  The dependent chain per column is
//     2 FMA + 2 FMA + SHFL; the shared-memory reads are address-independent of it.  Both access patterns are
//     bank-conflict free in the packed layout (a column is contiguous; along a row, 8 consecutive triangular
//     numbers are distinct mod 8);
//   * points are pulled from an atomic queue in longest-job-first order (predict_kernel), the cancel flag of
//     compute.jl:202-204 is polled through a mapped host word whenever a warp fetches its next point.
// The arithmetic per point is the reference's; only the summation order of the solves differs from the DMMA kernels
// (results agree to rounding, not bit for bit -- which kernel runs depends on n and maxit only, never on scheduling).
#pragma once
#include "psa_square.cuh"

namespace psa {

constexpr int PT_MAX_NPL = 6;     // vector elements per lane: n <= 192 by registers (shared memory binds earlier)
constexpr int PT_MAX_WARPS = 16;  // points in flight per SM
constexpr int PT_MIN_WARPS = 8;   // below this the tick engine is the better machine
constexpr int PT_SMEM_BUDGET = 227 * 1024 - 512;

struct PointParams {
  int n;
  const double2* T;  // n x n column-major, ld = ldt; only i <= j is read
  long long ldt;
  int P, rows_local, row0, row_step;  // local points p = k * rows_local + jr, grid row j = row0 + jr * row_step
  const double* x;
  const double* y;
  const int* order;      // nullable: queue position -> point
  const int* row_batch;  // nullable: grid row -> start-vector batch
  const double2* q0;     // [nbatch][n_pad]
  int n_pad;
  double tol;
  int maxit;
  double bigsigma;
  double* Z;    // local point order
  int* iters;   // nullable
  unsigned long long* counts;  // PSA_COUNT_*
  int* queue;                  // atomic queue head (zeroed before the launch)
  const volatile unsigned long long* stop;  // nullable, mapped host word: non-zero = stop fetching points
  double* hist_alpha;  // [gridDim.x * warps][hist]  unscaled Lanczos coefficients
  double* hist_beta;
  double2* qold;       // [gridDim.x * warps][2][PT_MAX_NPL * 32]  q and qold of the warp's point (touched once per iteration)
  int hist;            // stride of the above, and of the per-warp scaled copies in shared memory
};

__host__ __device__ inline size_t pt_tri_elems(int n) { return (size_t)n * (n + 1) / 2; }
// shared memory: the packed triangle, ONE zero element (what masked lanes read), then per warp 2 x hist doubles
__host__ __device__ inline size_t pt_mat_elems(int n) { return pt_tri_elems(n) + 1; }
inline size_t pt_smem(int n, int hist, int warps) { return pt_mat_elems(n) * sizeof(double2) + (size_t)warps * 2 * hist * sizeof(double); }
// warps per CTA that fit beside the matrix; 0 = this shape stays on the tick engine
inline int pt_warps(int n, int hist) {
  if ((n + 31) / 32 > PT_MAX_NPL) return 0;
  const long long left = (long long)PT_SMEM_BUDGET - (long long)(pt_mat_elems(n) * sizeof(double2));
  const long long w = left / (2LL * hist * (long long)sizeof(double));
  if (w < PT_MIN_WARPS) return 0;
  return (int)(w < PT_MAX_WARPS ? w : PT_MAX_WARPS);
}

// 16-byte shared-memory load by shared-window address (the matrix is read-only after staging: no memory clobber, the
// compiler is free to schedule these loads ahead of the dependent chain)
__device__ __forceinline__ double2 pt_lds(uint32_t addr) {
  double2 v;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

template <int NPL>
__global__ void __launch_bounds__(PT_MAX_WARPS * 32, 1) point_lanczos_kernel(PointParams S) {
  extern __shared__ __align__(16) unsigned char pt_smem_raw[];
  double2* Ts = reinterpret_cast<double2*>(pt_smem_raw);
  const int n = S.n;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* sd = reinterpret_cast<double*>(pt_smem_raw + pt_mat_elems(n) * sizeof(double2)) + (size_t)wid * 2 * S.hist;
  double* se2 = sd + S.hist;
  double* ha = S.hist_alpha + ((size_t)blockIdx.x * nw + wid) * S.hist;
  double* hb = S.hist_beta + ((size_t)blockIdx.x * nw + wid) * S.hist;
  // q and qold live in an L2-resident scratch (two buffers per warp that swap roles every iteration; each lane only ever
  // touches its own elements: element r at [32 r]): they are needed once per iteration, the registers belong to the sweeps
  constexpr int VEC = PT_MAX_NPL * 32;
  double2* const vecs = S.qold + ((size_t)blockIdx.x * nw + wid) * (2 * VEC) + lane;
  constexpr unsigned FULL = 0xffffffffu;

  // packed upper triangle: element (i, j), i <= j, at j (j + 1) / 2 + i
  for (int j = wid; j < n; j += nw) {
    const double2* col = S.T + (long long)j * S.ldt;
    double2* dst = Ts + (size_t)j * (j + 1) / 2;
    for (int i = lane; i <= j; i += 32) dst[i] = col[i];
  }
  const int zero_idx = (int)pt_tri_elems(n);
  if (threadIdx.x == 0) Ts[zero_idx] = make_double2(0.0, 0.0);
  __syncthreads();

  // Shared-window byte addresses.  Row j of T as this lane sees it: element (j, i), i = 32 r + lane, sits at
  // row0[r] + 16 j.  Lanes beyond the matrix (i >= n) and masked lanes read the zero element instead, so their vector
  // entries stay exactly 0 without any select on the FP64 side.
  const uint32_t ts0 = smem_u32(Ts), zaddr = ts0 + 16u * (uint32_t)zero_idx;

  for (;;) {
    int qpos = 0;
    if (lane == 0) {
      qpos = atomicAdd(S.queue, 1);
      if (S.stop && *S.stop != 0ull) qpos = S.P;
    }
    qpos = __shfl_sync(FULL, qpos, 0);
    if (qpos >= S.P) break;
    const int p = S.order ? S.order[qpos] : qpos;
    const int jr = p % S.rows_local, k = p / S.rows_local;
    const int jrow = S.row0 + jr * S.row_step;
    const double zx = S.x[k], zy = S.y[jrow];
    const double2* q0 = S.q0 + (size_t)(S.row_batch ? S.row_batch[jrow] : 0) * S.n_pad;

    double2 b[NPL];
    int cur = 0;  // buffer that holds q; the other one holds qold
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
      const int i = r * 32 + lane;
      __stcg(vecs + 32 * r, i < n ? q0[i] : make_double2(0.0, 0.0));  // q = q0[1:n], compute.jl:622
    }
    // 1 / (t_ii - z) for this lane's element of register block r: diag(T1) = diag(T) - z, compute.jl:595
    auto pivot_inv = [&](int r) {
      const int i = r * 32 + lane;
      double2 d = make_double2(1.0, 0.0);
      if (i < n) {
        d = Ts[i * (i + 1) / 2 + i];
        d.x -= zx;
        d.y -= zy;
      }
      return cinv_fixed(d);
    };
    double beta = 0.0, sigold = 0.0, sigma = 0.0;
    bool conv = false, failed = false;
    int l = 0;
    for (;;) {
      ++l;
      double2* const qv = vecs + cur * VEC;         // q
      double2* const qov = vecs + (cur ^ 1) * VEC;  // qold, overwritten by the next q at the end of the iteration
#pragma unroll
      for (int r = 0; r < NPL; ++r) b[r] = __ldcg(qv + 32 * r);

      // w = F1' \ q : forward sweep over the columns of L = F1^H, L[i][j] = conj(T[j][i]) (i > j), pivot conj(t_jj - z).
      // Within the diagonal register block an entry is final once its own column has been processed (later columns
      // are masked for it), so the pivot scaling of the stored entries is applied once per block, after its columns.
#pragma unroll
      for (int r = 0; r < NPL; ++r) {
        const int j0 = r * 32;
        const int cnt = min(32, n - j0);
        const double2 rdr = pivot_inv(r);
        const double2 rdc = make_double2(rdr.x, -rdr.y);
        uint32_t row[NPL], rstep[NPL];  // address of element (j, i) for the current column j, and its step per column
#pragma unroll
        for (int rr = r; rr < NPL; ++rr) {
          const int i = rr * 32 + lane;
          row[rr] = i < n ? ts0 + 16u * (uint32_t)(i * (i + 1) / 2 + j0) : zaddr;
          rstep[rr] = i < n ? 16u : 0u;
        }
        // not unrolled on purpose: 16 warps sit in 16 different column loops, and their bodies have to share the 32 KB
        // instruction cache (unrolled by two: README 400 x 400 grid 62 -> 84 ms)
#pragma unroll 1
        for (int ll = 0; ll < cnt; ++ll) {
          const double2 c = cmul_fixed(b[r], rdc);
          const double wx = __shfl_sync(FULL, c.x, ll), wy = __shfl_sync(FULL, c.y, ll);
          const double nwx = -wx, nwy = -wy;
#pragma unroll
          for (int rr = r; rr < NPL; ++rr) {
            const double2 t = pt_lds((rr == r && lane <= ll) ? zaddr : row[rr]);  // T[j][i]; only i > j is updated
            row[rr] += rstep[rr];
            b[rr].x = fma(t.x, nwx, b[rr].x);  // b_i -= conj(t) * w_j
            b[rr].x = fma(t.y, nwy, b[rr].x);
            b[rr].y = fma(t.x, nwy, b[rr].y);
            b[rr].y = fma(t.y, wx, b[rr].y);
          }
        }
        b[r] = cmul_fixed(b[r], rdc);
      }
      // v = F1 \ w : backward sweep over the columns of T
#pragma unroll
      for (int r = NPL - 1; r >= 0; --r) {
        const int j0 = r * 32;
        const int cnt = min(32, n - j0);
        const double2 rdr = pivot_inv(r);
        int j = j0 + cnt - 1;
        uint32_t col = ts0 + 16u * (uint32_t)(j * (j + 1) / 2 + lane);  // address of element (lane, j)
#pragma unroll 1
        for (int ll = cnt - 1; ll >= 0; --ll, --j) {
          const double2 c = cmul_fixed(b[r], rdr);
          const double vx = __shfl_sync(FULL, c.x, ll), vy = __shfl_sync(FULL, c.y, ll);
          const double nvx = -vx, nvy = -vy;
#pragma unroll
          for (int rr = 0; rr <= r; ++rr) {
            const double2 t = pt_lds((rr == r && lane >= ll) ? zaddr : col + 512u * rr);  // T[i][j]; only i < j is updated
            b[rr].x = fma(t.x, nvx, b[rr].x);  // b_i -= t * v_j
            b[rr].x = fma(t.y, vy, b[rr].x);
            b[rr].y = fma(t.x, nvy, b[rr].y);
            b[rr].y = fma(t.y, nvx, b[rr].y);
          }
          col -= 16u * (uint32_t)j;  // column j - 1 starts j elements earlier
        }
        b[r] = cmul_fixed(b[r], rdr);
      }

      // v -= beta * qold ; alpha = Re(q . v)          (compute.jl:630-631; same order as lanczos_warp_kernel)
      double part = 0.0;
      double2 q[NPL];
#pragma unroll
      for (int r = 0; r < NPL; ++r) {
        q[r] = __ldcg(qv + 32 * r);
        if (l > 1) {
          const double2 o = __ldcg(qov + 32 * r);
          b[r].x -= beta * o.x;
          b[r].y -= beta * o.y;
        }
        part = fma(q[r].x, b[r].x, part);
        part = fma(q[r].y, b[r].y, part);
      }
      const double alpha = warp_sum(part);
      // v -= alpha q ; beta = |v|                      (compute.jl:632-633)
      part = 0.0;
#pragma unroll
      for (int r = 0; r < NPL; ++r) {
        b[r].x -= alpha * q[r].x;
        b[r].y -= alpha * q[r].y;
        part = fma(b[r].x, b[r].x, part);
        part = fma(b[r].y, b[r].y, part);
      }
      const double ss = warp_sum(part);
      if (isfinite(ss) && ss > 1e-280) {
        beta = sqrt(ss);
      } else {  // scaled 2-norm, as dznrm2
        double mxv = 0.0;
#pragma unroll
        for (int r = 0; r < NPL; ++r) mxv = fmax(mxv, fmax(fabs(b[r].x), fabs(b[r].y)));
        mxv = warp_max(mxv);
        if (mxv == 0.0 || !isfinite(mxv)) {
          beta = mxv == 0.0 ? 0.0 : ss;
        } else {
          part = 0.0;
#pragma unroll
          for (int r = 0; r < NPL; ++r) {
            const double a_ = b[r].x / mxv, b_ = b[r].y / mxv;
            part = fma(a_, a_, part);
            part = fma(b_, b_, part);
          }
          beta = mxv * sqrt(warp_sum(part));
        }
      }
      // H[l,l] = alpha, H[l+1,l] = H[l,l+1] = beta     (compute.jl:636-638)
      const bool finite_ab = isfinite(alpha) && isfinite(beta);
      if (lane == 0) {
        __stcg(ha + l - 1, alpha);
        __stcg(hb + l - 1, beta);
      }
      __syncwarp();
      double mx = 0.0;
      if (finite_ab) {
        for (int i = lane; i < l; i += 32) {
          mx = fmax(mx, fabs(__ldcg(ha + i)));
          if (i < l - 1) mx = fmax(mx, fabs(__ldcg(hb + i)));
        }
      }
      mx = warp_max(mx);
      if (finite_ab && mx > 0.0) {
        for (int i = lane; i < l; i += 32) {
          sd[i] = __ldcg(ha + i) / mx;
          if (i < l - 1) {
            const double b_ = __ldcg(hb + i) / mx;
            se2[i] = b_ * b_;
          }
        }
      }
      __syncwarp();
      // sigma = maximum(eigvals(H[1:l,1:l]))           (compute.jl:640-642; non-finite -> bigsigma as 651-656)
      if (!finite_ab) {
        sigma = S.bigsigma;
        failed = true;
      } else if (mx == 0.0) {
        sigma = 0.0;
      } else {
        sigma = mx * tridiag_max_eig_warp(l, sd, se2, lane);
      }
      __syncwarp();
      conv = !failed && (fabs(sigold / sigma - 1.0) < S.tol || beta == 0.0);  // compute.jl:658-661
      if (conv || failed || l >= S.maxit) break;
      sigold = sigma;
#pragma unroll
      for (int r = 0; r < NPL; ++r)  // qold = q ; q = v / beta   (compute.jl:634-635): the buffers swap roles
        __stcg(qov + 32 * r, make_double2(b[r].x / beta, b[r].y / beta));
      cur ^= 1;
    }
    if (lane == 0) {
      S.Z[p] = 1.0 / sqrt(sigma);
      if (S.iters) S.iters[p] = l;
      if (!conv) atomicAdd(&S.counts[0], 1ull);  // incl. fallback points (compute.jl:654-656 leaves lancz_converged false)
      if (failed) atomicAdd(&S.counts[1], 1ull);
      atomicAdd(&S.counts[2], (unsigned long long)l);
      atomicAdd(&S.counts[3], 1ull);
    }
  }
}

inline int point_set_attrs() {
  const int bytes = PT_SMEM_BUDGET;
  cudaError_t e = cudaFuncSetAttribute(point_lanczos_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(point_lanczos_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(point_lanczos_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(point_lanczos_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(point_lanczos_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  return e == cudaSuccess ? 0 : -1;
}

// grid: one CTA per SM at most, never more warps than points
inline void launch_point_lanczos(const PointParams& pp, int ctas, int warps, cudaStream_t stream) {
  const size_t smem = pt_smem(pp.n, pp.hist, warps);
  const int npl = (pp.n + 31) / 32;
  const dim3 grid(ctas), block(warps * 32);
  switch (npl <= 2 ? 2 : npl) {
    case 2: point_lanczos_kernel<2><<<grid, block, smem, stream>>>(pp); break;
    case 3: point_lanczos_kernel<3><<<grid, block, smem, stream>>>(pp); break;
    case 4: point_lanczos_kernel<4><<<grid, block, smem, stream>>>(pp); break;
    case 5: point_lanczos_kernel<5><<<grid, block, smem, stream>>>(pp); break;
    default: point_lanczos_kernel<6><<<grid, block, smem, stream>>>(pp); break;
  }
}

}  // namespace psa
