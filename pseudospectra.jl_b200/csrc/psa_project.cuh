// Device-side Schur reordering for the Trefethen/Wright projection (the heavy part of `_maybe_project`,
// /root/reference/src/compute.jl:339-357):
//     for i = 1:m,  for k = selection[i]-1:-1:i
//         G = givens(conj(T[k,k+1]), conj(T[k,k] - T[k+1,k+1]), k+1, k);  rmul!(T, G');  lmul!(G, T)
// Every (i, k) is a Givens similarity that swaps the diagonal entries k, k+1: eigenvalue selection[i] "bubbles" up to
// position i.  The reference runs these O(np * n) swaps one after the other, each touching two rows and two columns
// (O(n) work): tens of seconds of host time at n = 4000.  Swaps whose index pairs are at least two apart act on
// disjoint rows and disjoint columns and their rotations are computed from disjoint diagonal blocks, so they commute.
// Bubble i+1 can therefore follow bubble i at a distance of two positions: a systolic wavefront with up to np/2 swaps
// in flight.  One wavefront step = three small launches: rotation parameters, all column rotations, all row rotations
// (left and right multiplications commute, so "all columns, then all rows" equals the per-swap order in exact
// arithmetic; in floating point the result differs from the sequential order by rounding only).
#pragma once
#include "psa_common.cuh"

namespace psa {

struct ProjRot {
  int k;       // swap position (0-based: diagonal entries k, k+1) or -1 when the bubble is idle in this step
  double c;
  double2 s;   // the `sn` of compute.py:_maybe_project: G = [c sn; -conj(sn) c] on rows k, k+1
};

// zlartg-convention (c, s) of Julia's givensAlgorithm(f, g): [c s; -conj(s) c] [f; g] = [r; 0], c real
__device__ __forceinline__ void proj_givens(double2 f, double2 g, double& c, double2& s) {
  if (g.x == 0.0 && g.y == 0.0) {
    c = 1.0;
    s = make_double2(0.0, 0.0);
    return;
  }
  const double ag = hypot(g.x, g.y);
  if (f.x == 0.0 && f.y == 0.0) {
    c = 0.0;
    s = make_double2(g.x / ag, -g.y / ag);
    return;
  }
  const double af = hypot(f.x, f.y);
  const double d = hypot(af, ag);
  c = af / d;
  const double2 ph = make_double2(f.x / af, f.y / af);
  s = cmul(ph, make_double2(g.x / d, -g.y / d));
}

// one thread per bubble: where is it in wavefront step t, and with which rotation?
__global__ void proj_params_kernel(int t, int np, const int* __restrict__ sel, const int* __restrict__ start,
                                   const double2* __restrict__ T, long long ld, ProjRot* __restrict__ rot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const int j = t - start[i], len = sel[i] - i;
  ProjRot r;
  r.k = -1;
  r.c = 1.0;
  r.s = make_double2(0.0, 0.0);
  if (j >= 0 && j < len) {
    const int k = sel[i] - 1 - j;
    const double2 tkk = T[k + k * ld], tk1 = T[k + (k + 1) * ld], t11 = T[(k + 1) + (k + 1) * ld];
    // givens(conj(T[k,k+1]), conj(T[k,k] - T[k+1,k+1]), k+1, k): i1 > i2 flips the rotation (s -> -conj(s))
    double c;
    double2 s;
    proj_givens(make_double2(tk1.x, -tk1.y), make_double2(tkk.x - t11.x, -(tkk.y - t11.y)), c, s);
    r.k = k;
    r.c = c;
    r.s = make_double2(-s.x, s.y);  // -conj(s)
  }
  rot[i] = r;
}

// rmul!(T, G'): columns k, k+1, rows 0..k+1 (the matrix is upper triangular up to the entry the row rotation removes)
__global__ void __launch_bounds__(256) proj_cols_kernel(int np, const ProjRot* __restrict__ rot, double2* __restrict__ T,
                                                        long long ld) {
  const ProjRot r = rot[blockIdx.x];
  if (r.k < 0) return;
  double2* c0 = T + (long long)r.k * ld;
  double2* c1 = c0 + ld;
  const double2 scj = make_double2(r.s.x, -r.s.y), sneg = make_double2(-r.s.x, -r.s.y);
  for (int row = blockIdx.y * blockDim.x + threadIdx.x; row <= r.k + 1; row += gridDim.y * blockDim.x) {
    const double2 a1 = c0[row], a2 = c1[row];
    double2 n0 = cmul(a2, scj);  // a1*c + a2*conj(sn)
    n0.x = fma(a1.x, r.c, n0.x);
    n0.y = fma(a1.y, r.c, n0.y);
    double2 n1 = cmul(a1, sneg);  // -a1*sn + a2*c
    n1.x = fma(a2.x, r.c, n1.x);
    n1.y = fma(a2.y, r.c, n1.y);
    c0[row] = n0;
    c1[row] = n1;
  }
}

// lmul!(G, T): rows k, k+1, columns k..n-1
__global__ void __launch_bounds__(256) proj_rows_kernel(int np, int n, const ProjRot* __restrict__ rot,
                                                        double2* __restrict__ T, long long ld) {
  const ProjRot r = rot[blockIdx.x];
  if (r.k < 0) return;
  const double2 sncj = make_double2(-r.s.x, r.s.y);  // -conj(sn)
  for (int col = r.k + blockIdx.y * blockDim.x + threadIdx.x; col < n; col += gridDim.y * blockDim.x) {
    double2* p = T + r.k + (long long)col * ld;
    const double2 r1 = p[0], r2 = p[1];
    double2 n0 = cmul(r.s, r2);  // c*r1 + sn*r2
    n0.x = fma(r.c, r1.x, n0.x);
    n0.y = fma(r.c, r1.y, n0.y);
    double2 n1 = cmul(sncj, r1);  // -conj(sn)*r1 + c*r2
    n1.x = fma(r.c, r2.x, n1.x);
    n1.y = fma(r.c, r2.y, n1.y);
    if (col == r.k) n1 = make_double2(0.0, 0.0);  // the entry the swap annihilates (the reference's triu() drops it)
    p[0] = n0;
    p[1] = n1;
  }
}

// work <- triu(src) (the reference starts from `copy(Matrix(Targ))` of an UpperTriangular)
__global__ void proj_copy_triu_kernel(const double2* __restrict__ src, long long lds, int n, double2* __restrict__ dst,
                                      long long ldd) {
  const int j = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i + (long long)j * ldd] = i <= j ? src[i + (long long)j * lds] : make_double2(0.0, 0.0);
}

// out <- triu(work[0:np, 0:np])   (`UpperTriangular(triu(Tproj[1:m,1:m]))`, compute.jl:357)
__global__ void proj_extract_kernel(const double2* __restrict__ work, long long ldw, int np, double2* __restrict__ out) {
  const int j = blockIdx.x;
  for (int i = threadIdx.x; i < np; i += blockDim.x) out[i + (long long)j * np] = i <= j ? work[i + (long long)j * ldw] : make_double2(0.0, 0.0);
}

}  // namespace psa
