// EXPERIMENT (not built into libpsagpu): 512-thread CTA per SM, 32-column steps.  Measured on B200 at n = 4000:
// 76.8 ms per full-occupancy launch vs 77.0 ms for the shipped 256-thread kernel, with a worse latency floor
// (narrowest panel 32 shifts/SM: 23.6 ms vs 16.1 ms), so it was not adopted.  Kept for the record.
//
// Square (:sq_lanc) path, second-generation solve kernel: one 512-thread CTA per SM, 32-column steps.
//
// Same algorithm and data flow as psa_square.cuh's solve_kernel (left-looking block forward substitution on the
// pre-packed lower-triangular operators L_A = T^H and L_B = reversed T, FP64 DMMA tiles, TMA-engine bulk copies
// into an mbarrier ring) but with twice the work per pipeline step (KC2 = 32 columns) and 16 warps per CTA, so
// the per-step synchronisation cost is amortised over 4x more tensor work per SM:
//   16 warps = 4 (M) x 4 (N); warp tile 16 rows x NSV/4 shifts; NSV in {128, 96, 64, 32} shifts per CTA.
// Replaces the two ztrsv calls of /root/reference/src/compute.jl:630 for NSV shifts per CTA.
#pragma once
#include "psa_square.cuh"

namespace psa {

constexpr int KC2 = 32;                      // contraction chunk per step
constexpr int CPB2 = MB / KC2;               // 2 chunks per 64-row block
constexpr int TILE2_ELEMS = MB * KC2;        // complex elements per packed tile
constexpr int TILE2_BYTES = TILE2_ELEMS * 16;  // 32 KiB
constexpr int XROW2_BYTES = KC2 * 16 + 64;   // 576: 64 B pad keeps B-fragment LDS.128 conflict-free
constexpr int SOLVE2_THREADS = 512;

__host__ __device__ constexpr int stages2(int nsv) { return nsv > 64 ? 2 : (nsv > 32 ? 3 : 4); }
__host__ __device__ constexpr int stage2_bytes(int nsv) { return TILE2_BYTES + nsv * XROW2_BYTES; }
__host__ __device__ constexpr int solve2_smem(int nsv) {
  return stages2(nsv) * stage2_bytes(nsv) + 64 + nsv * 16 + nsv * 8 * 2;
}
__host__ __device__ inline long long tiles2_per_phase(int nblk) { return (long long)nblk * (nblk + 1); }
__host__ __device__ inline long long tile2_index(int I, int c) { return (long long)I * (I + 1) + c; }

// Packing for 32-column tiles; same operator definitions and fragment order as pack_kernel (psa_square.cuh):
// element (m, kk), kk in [0,32) -> double2 index ((kk>>2)*8 + (m>>3))*32 + (m&7)*4 + (kk&3).
__global__ void pack2_kernel(const double2* __restrict__ T, long long ldt, int n, int n_pad, int nblk,
                             double2* __restrict__ packA, double2* __restrict__ packB) {
  const long long ntiles = tiles2_per_phase(nblk);
  for (long long tile = blockIdx.x; tile < 2 * ntiles; tile += gridDim.x) {
    const int phase = tile >= ntiles;
    const long long t = phase ? tile - ntiles : tile;
    int I = (int)((sqrt(1.0 + 4.0 * (double)t) - 1.0) * 0.5);
    while (tile2_index(I + 1, 0) <= t) ++I;
    while (tile2_index(I, 0) > t) --I;
    const int c = (int)(t - tile2_index(I, 0));
    double2* out = (phase ? packB : packA) + t * TILE2_ELEMS;
    for (int e = threadIdx.x; e < TILE2_ELEMS; e += blockDim.x) {
      int m, kk;
      if (phase == 0) {
        kk = e & 31;
        m = e >> 5;
      } else {
        m = e & 63;
        kk = e >> 6;
      }
      const int p = MB * I + m, pp = KC2 * c + kk;
      double2 v = make_double2(0.0, 0.0);
      if (pp <= p) {
        const int op = phase ? n_pad - 1 - p : p, opp = phase ? n_pad - 1 - pp : pp;
        if (op < n && opp < n) {
          if (phase == 0) {
            v = T[(long long)opp + (long long)op * ldt];
            v.y = -v.y;
          } else {
            v = T[(long long)op + (long long)opp * ldt];
          }
        } else if (p == pp) {
          v.x = 1.0;
        }
      }
      out[((kk >> 2) * 8 + (m >> 3)) * 32 + (m & 7) * 4 + (kk & 3)] = v;
    }
  }
}

struct Step2 {
  int phase, I, c;
  __device__ __forceinline__ void advance(int nblk) {
    if (++c == CPB2 * I + CPB2) {
      c = 0;
      if (++I == nblk) {
        I = 0;
        ++phase;
      }
    }
  }
};

template <int NSV>
__global__ void __launch_bounds__(SOLVE2_THREADS, 1) solve2_kernel(SolveParams P) {
  constexpr int SW = NSV / 4;  // shifts per warp
  constexpr int NT = SW / 8;   // DMMA column tiles per warp
  constexpr int STG = stages2(NSV);
  constexpr int STAGE_BYTES = stage2_bytes(NSV);
  extern __shared__ __align__(128) unsigned char smem[];
  const int n_active = *P.n_active;
  const int panel = blockIdx.x;
  if (panel * NSV >= n_active) return;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STG * STAGE_BYTES);  // 64 B reserved
  double2* zsh = reinterpret_cast<double2*>(smem + STG * STAGE_BYTES + 64);
  double2** wptr = reinterpret_cast<double2**>(zsh + NSV);
  const double2** qptr = (const double2**)(wptr + NSV);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2, wn = warp & 3;  // the 4 warps sharing a row group sit on the 4 SM sub-partitions

  if (tid < NSV) {
    const int slot = P.active[panel * NSV + tid];
    wptr[tid] = P.Wv + (size_t)slot * P.n_pad;
    qptr[tid] = (P.sel[slot] ? P.Q1 : P.Q0) + (size_t)slot * P.n_pad;
    zsh[tid] = P.z[slot];
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STG; ++s) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int nblk = P.nblk, n_pad = P.n_pad, n = P.n;
  const int steps_per_phase = (int)tiles2_per_phase(nblk);
  const int total_steps = 2 * steps_per_phase;

  // Called by ALL warps (see psa_square.cuh: UBLKCP is serialised per lane, so the row copies are spread over the
  // 16 warps instead of queueing on one).
  auto issue = [&](const Step2& pc, int stage) {
    constexpr int RPW = NSV / (SOLVE2_THREADS / 32);
    unsigned char* sb = smem + stage * STAGE_BYTES;
    const bool needX = pc.c < CPB2 * pc.I;
    if (tid == 0) {
      mbar_arrive_expect_tx(&full_bar[stage], TILE2_BYTES + (needX ? NSV * KC2 * 16 : 0));
      const double2* src = (pc.phase ? P.packB : P.packA) + tile2_index(pc.I, pc.c) * TILE2_ELEMS;
      bulk_g2s(sb, src, TILE2_BYTES, &full_bar[stage]);
    }
    if (needX && lane < RPW) {
      const int mem0 = pc.phase ? n_pad - KC2 * pc.c - KC2 : KC2 * pc.c;
      const int s = warp * RPW + lane;
      bulk_g2s(sb + TILE2_BYTES + s * XROW2_BYTES, wptr[s] + mem0, KC2 * 16, &full_bar[stage]);
    }
  };

  double cre[2][NT][2], cim[2][NT][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b) cre[a][b][0] = cre[a][b][1] = cim[a][b][0] = cim[a][b][1] = 0.0;

  // acc += L[rows of this warp][k4 steps j0..j1) * X
  auto gemm_part = [&](const unsigned char* sb, int flip, int j0, int j1) {
    const double2* At = reinterpret_cast<const double2*>(sb);
    const unsigned char* Xt = sb + TILE2_BYTES + (SW * wn + (lane >> 2)) * XROW2_BYTES;
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
      double2 a[2], b[NT];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) a[mt] = At[(j * 8 + 2 * wm + mt) * 32 + lane];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        b[nt] = *reinterpret_cast<const double2*>(Xt + nt * 8 * XROW2_BYTES + (((4 * j + (lane & 3)) ^ flip) << 4));
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cre[mt][nt][0], cre[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cim[mt][nt][0], cim[mt][nt][1], a[mt].x, b[nt].y);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const double nai = -a[mt].y;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cre[mt][nt][0], cre[mt][nt][1], nai, b[nt].y);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cim[mt][nt][0], cim[mt][nt][1], a[mt].y, b[nt].x);
    }
  };

  // The warps owning rows [16*wm, 16*wm+16) of the block finish them: transpose acc through the stage's X buffer,
  // then one thread per shift solves the 16x16 triangular system in registers.  `sub` = 0/1: which half of the
  // 32-column chunk holds this sub-block's columns.
  auto diag_solve = [&](unsigned char* sb, const Step2& st, int sub) {
    const int flip = st.phase ? (KC2 - 1) : 0;
    const bool solver = lane < SW;
    const int s = SW * wn + (solver ? lane : 0);
    const int mem0 = st.phase ? n_pad - KC2 * st.c - KC2 : KC2 * st.c;
    const int kk0 = 16 * sub;  // first column of the sub-block inside the chunk (processing order)
    const double2* rp = (st.phase ? (const double2*)wptr[s] : qptr[s]) + mem0;
    double2 b[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) b[r] = rp[(kk0 + r) ^ flip];
    unsigned char* Xt = sb + TILE2_BYTES;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = 8 * mt + (lane >> 2), sl = SW * wn + 8 * nt + 2 * (lane & 3) + e;
          *reinterpret_cast<double2*>(Xt + sl * XROW2_BYTES + (((kk0 + r) ^ flip) << 4)) =
              make_double2(cre[mt][nt][e], cim[mt][nt][e]);
          cre[mt][nt][e] = 0.0;
          cim[mt][nt][e] = 0.0;
        }
    __syncwarp();
    double2* xrow = reinterpret_cast<double2*>(Xt + s * XROW2_BYTES);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const double2 a = xrow[(kk0 + r) ^ flip];
      b[r].x -= a.x;
      b[r].y -= a.y;
    }
    const double2* At = reinterpret_cast<const double2*>(sb);
    double2 zt = zsh[s];
    if (st.phase == 0) zt.y = -zt.y;
    const int mrow0 = 16 * wm;                 // first row of the sub-block inside the 64-row tile
    const int p0 = MB * st.I + mrow0;          // processing index of that row
#if !(defined(PSA_EXP) && (PSA_EXP & 4))
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int m = mrow0 + r, kk = kk0 + r;
      double2 d = At[((kk >> 2) * 8 + (m >> 3)) * 32 + (m & 7) * 4 + (kk & 3)];
      d.x -= zt.x;
      d.y -= zt.y;
      double2 x = cmul(b[r], crecip(d));
      const int op = st.phase ? n_pad - 1 - (p0 + r) : p0 + r;
      if (op >= n) x = make_double2(0.0, 0.0);
      b[r] = x;
#pragma unroll
      for (int r2 = r + 1; r2 < 16; ++r2) {
        const int m2 = mrow0 + r2;
        const double2 l = At[((kk >> 2) * 8 + (m2 >> 3)) * 32 + (m2 & 7) * 4 + (kk & 3)];
        cmsub(b[r2], l, x);
      }
    }
#endif
    double2* op_ = wptr[s] + mem0;
    if (solver) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        xrow[(kk0 + r) ^ flip] = b[r];
        op_[(kk0 + r) ^ flip] = b[r];
      }
    }
    fence_proxy_async();
  };

  // linear step index of the step that PRODUCES chunk c of a phase: the diagonal step (c / CPB2, c)
  auto produced_at = [&](int phase, int c) { return phase * steps_per_phase + (int)tile2_index(c / CPB2, c); };

  Step2 cons{0, 0, 0}, prod{0, 0, 0};
  int prod_idx = 0;  // linear index of the next step to issue
  // issue every step that fits the ring (distance < STG) and whose X chunk has already been produced
  auto pump = [&](int done_upto /* last completed step, -1 at start */) {
    while (prod_idx < total_steps && prod_idx <= done_upto + STG) {
      const bool needX = prod.c < CPB2 * prod.I;
      if (needX && produced_at(prod.phase, prod.c) > done_upto) break;
      issue(prod, prod_idx % STG);
      prod.advance(nblk);
      ++prod_idx;
    }
  };
  pump(-1);

  for (int t = 0; t < total_steps; ++t) {
    const int stage = t % STG;
    const uint32_t parity = (uint32_t)(t / STG) & 1u;
    unsigned char* sb = smem + stage * STAGE_BYTES;
    const int flip = cons.phase ? (KC2 - 1) : 0;
    mbar_wait(&full_bar[stage], parity);

    if (cons.c < CPB2 * cons.I) {
      gemm_part(sb, flip, 0, 8);
    } else {
      const int d = cons.c - CPB2 * cons.I;  // diagonal chunk 0/1: rows 32d..32d+31 of the block
#if defined(PSA_EXP) && (PSA_EXP & 1)
      if (d < 0)  // timing experiment only: skip the diagonal work (wrong results)
#endif
      if (wm == 2 * d) diag_solve(sb, cons, 0);
      __syncthreads();
#if defined(PSA_EXP) && (PSA_EXP & 1)
      if (d < 0)
#endif
      if (wm == 2 * d + 1) {
        gemm_part(sb, flip, 0, 4);  // couple in the first half of the chunk
        diag_solve(sb, cons, 1);
      }
      __syncthreads();
#if defined(PSA_EXP) && (PSA_EXP & 1)
      if (d < 0)
#endif
      if (wm > 2 * d + 1) gemm_part(sb, flip, 0, 8);
    }
#if defined(PSA_EXP) && (PSA_EXP & 8)
    __syncwarp();  // timing experiment only (racy)
#else
    __syncthreads();  // stage free
#endif
    pump(t);
    cons.advance(nblk);
  }
}

}  // namespace psa
