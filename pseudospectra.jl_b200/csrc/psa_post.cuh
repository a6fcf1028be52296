// Kernels either side of the per-point work: input structure checks, assembly of the row-sharded result on device 0,
// device-side post-processing of the grid (the reference's un-mirror / clamp / level statistics,
// /root/reference/src/compute.jl:222-238 and the reductions of recalc_levels, src/utils.jl:24-31), and the small
// vector kernels of the pseudomode entry point (src/modes.jl:185-274).
#pragma once
#include "psa_common.cuh"

namespace psa {

// ---- structure of the input matrix ------------------------------------------------------------------
// Counts non-zero entries with i >= j + k (k = 1: strictly lower part, istriu; k = 2: below the first
// sub-diagonal, upper Hessenberg).  The reference tests `istriu(T1)` before treating T1 as UpperTriangular
// (compute.jl:597-601) and routes only Hessenberg input to the Givens / qr branches (compute.jl:579-590).
__global__ void __launch_bounds__(256) below_band_count_kernel(const double2* __restrict__ T, long long ldt, int m, int n,
                                                               int k, unsigned long long* __restrict__ nnz) {
  unsigned long long cnt = 0;
  for (int j = blockIdx.x; j < n; j += gridDim.x) {
    const double2* col = T + (long long)j * ldt;
    for (int i = j + k + threadIdx.x; i < m; i += blockDim.x) {
      const double2 v = col[i];
      cnt += (v.x != 0.0 || v.y != 0.0);
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, (unsigned)cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(nnz, cnt);
}

// ---- assembly of the cyclically sharded result ---------------------------------------------------------
// tile: rows_i x lx column-major (device i's rows i, i+ng, ...) -> full: ly x lx column-major.
template <typename T>
__global__ void interleave_rows_kernel(const T* __restrict__ tile, int rows_i, int lx, int i, int ng, int ly,
                                       T* __restrict__ full) {
  const long long total = (long long)rows_i * lx;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int jr = (int)(e % rows_i), k = (int)(e / rows_i);
    full[(long long)k * ly + i + (long long)jr * ng] = tile[e];
  }
}

// ---- post-processing on the device ------------------------------------------------------------------------
// Un-mirror (compute.jl:222-235) + clamp!(Z, ps_tiny, Inf) (236-238) + the three reductions recalc_levels needs
// (utils.jl:24-31): min, max and the smallest value >= small_sigma.  Z: ly x lx column-major (computed rows),
// out: lyf x lx column-major with lyf = ly + |n_mirror| (or ly + n_mirror rows taken from 1.. when y[0] == 0).
// Positive doubles order like their bit patterns, so the reductions use 64-bit integer atomics.
struct PostParams {
  const double* Z;
  int ly, lx;
  int n_mirror;   // as returned by shift_axes (utils.jl:112-141); 0 = no symmetry used
  int y0_is_zero; // y[1] == 0 (the mirrored rows then start at the second row, compute.jl:230-233)
  double ps_tiny, small_sigma;
  double* out;
  int lyf;
  unsigned long long* stats;  // [0] min bits, [1] max bits, [2] min over values >= small_sigma (bits)
};

__device__ __forceinline__ int post_source_row(const PostParams& P, int r) {
  // row r of the un-mirrored array comes from which computed row?
  if (P.n_mirror < 0) {  // bottom half is master: [Z; reverse(Z[end+nm+1:end, :])]
    return r < P.ly ? r : P.ly - 1 - (r - P.ly);
  }
  const int nm = P.n_mirror;
  if (r >= nm) return r - nm;
  return P.y0_is_zero ? nm - r : nm - 1 - r;  // reverse(Z[2:nm+1]) resp. reverse(Z[1:nm])
}

__global__ void __launch_bounds__(256) postprocess_kernel(PostParams P) {
  unsigned long long lmin = ~0ull, lmax = 0ull, lmin2 = ~0ull;
  const long long total = (long long)P.lyf * P.lx;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e % P.lyf), k = (int)(e / P.lyf);
    double v = P.Z[(long long)k * P.ly + post_source_row(P, r)];
    if (v < P.ps_tiny) v = P.ps_tiny;  // clamp!(Z, ps_tiny, Inf); NaN stays NaN like Julia's clamp
    P.out[e] = v;
    if (v == v) {
      const unsigned long long b = (unsigned long long)__double_as_longlong(v);
      lmin = min(lmin, b);
      lmax = max(lmax, b);
      if (v >= P.small_sigma) lmin2 = min(lmin2, b);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    lmin2 = min(lmin2, __shfl_xor_sync(0xffffffffu, lmin2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&P.stats[0], lmin);
    atomicMax(&P.stats[1], lmax);
    atomicMin(&P.stats[2], lmin2);
  }
}

// ---- pseudomode helpers --------------------------------------------------------------------------------------
// basis[:, col] <- the slot's current Lanczos vector (the start vector while the slot is fresh)
__global__ void basis_store_kernel(int n, const double2* __restrict__ q, double2* __restrict__ basis_col) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) basis_col[i] = q[i];
}

// mode = basis[:, 0:l] * yv   (`Q[:,1:end-1] * E.vectors[:,end]`, modes.jl:257); basis is n x l column-major
__global__ void basis_combine_kernel(int n, int l, const double2* __restrict__ basis, const double* __restrict__ yv,
                                     double2* __restrict__ mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double sr = 0.0, si = 0.0;
  for (int c = 0; c < l; ++c) {
    const double2 b = basis[(size_t)c * n + i];
    const double w = yv[c];
    sr = fma(w, b.x, sr);
    si = fma(w, b.y, si);
  }
  mode[i] = make_double2(sr, si);
}

}  // namespace psa
