// Cluster-pipelined solve kernel for SPARSE launches (fewer panels than SMs / P): grid tails, strong scaling of one
// grid over many GPUs, the pseudomode -- the regime where the run time is (longest point's iteration count) x (latency
// of ONE panel), and one panel on one SM cannot go faster than its DMMA issue rate.
//
// A thread-block cluster of P CTAs (P SMs) shares ONE panel.  The left-looking accumulation of block row I,
//     acc = sum_{c < 4I} L[I][c] X[c],
// is cut into P contiguous ranges of c.  CTA r adds its range to the accumulators it RECEIVES from CTA r-1 and hands
// the result to CTA r+1 through distributed shared memory; the last CTA ("main") finishes the sum and does the
// diagonal steps.  CTA r works one block row ahead of CTA r+1 -- a P-stage pipeline over block rows:
//   * the summation order is unchanged (c ascending, one accumulation chain), so the result is BIT-IDENTICAL to the
//     single-CTA kernels, whatever P the scheduler picks for a tick;
//   * the ranges of the helper CTAs only contain chunks of x that main finished at least one block row earlier;
//   * the diagonal steps (serial, per shift) of row I overlap with the helpers' GEMM work on rows I+1 .. I+P-1.
// Inside each CTA the choreography is the deep-ring one (producer warp, full/empty mbarriers, 8 consumer warps).
// Cross-CTA protocol (all mbarriers / flags live in shared memory, remote ones reached with mapa):
//   acc_ready[slot]  in the receiver : CW remote arrivals (release.cluster) after the sender's warps stored their
//                                      accumulator fragments into the receiver's hand-off buffer slot;
//   acc_free[slot]   in the sender   : CW remote arrivals after the receiver's warps loaded the slot;
//   rows_done        in every CTA    : number of block rows (both phases, linear) whose x main has written to global
//                                      memory (st.release.cluster after a gpu-scope fence); a producer warp only issues
//                                      the load of chunk c once the row that produced it is done.
#pragma once
#include "psa_square.cuh"

namespace psa {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_arrive_wait() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f64x2(uint32_t raddr, double a, double b) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(raddr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait_acq_cluster(bar, parity)) {
  }
}
__device__ __forceinline__ void st_release_cluster_u32(uint32_t raddr, uint32_t v) {
  asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(raddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cluster_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

// Range of chunks [begin, end) of block row I that pipeline stage r (of P) accumulates.  DC = cost of main's four
// diagonal steps in units of one GEMM step (it gets that much less GEMM work); a helper that runs d = P-1-r rows ahead
// of main only takes chunks of rows <= I-d-1 (they are final by then).
template <int P, int DC>
__device__ __forceinline__ void cluster_range(int I, int r, int& begin, int& end) {
  const int total = CPB * I;
  const int share = (total + DC + P - 1) / P;
  auto bound = [&](int k) -> int {  // upper end of stage k-1's range == lower end of stage k's
    if (k <= 0) return 0;
    if (k >= P) return total;
    const int avail = CPB * max(0, I - (P - k));  // stage k-1 runs P-1-(k-1) = P-k rows ahead
    return min(min(k * share, total), avail);
  };
  begin = bound(r);
  end = bound(r + 1);
}

template <int NSV, int S, int P>
__host__ __device__ constexpr int cluster_smem() {
  return S * deep_stage_bytes(NSV) + 256 + 2 * (MB * NSV * 16) + NSV * 16 + NSV * 8 * 2;
}

template <int NSV, int S, int P, int DC>
__global__ void __cluster_dims__(P, 1, 1) __launch_bounds__(deep_threads(8), 1) solve_cluster_kernel(SolveParams Pm) {
  constexpr int CW = 8, NT = NSV / 8, CTHREADS = CW * 32;
  constexpr int STAGE_BYTES = deep_stage_bytes(NSV);
  constexpr int HAND_BYTES = MB * NSV * 16;  // one accumulator set: 64 rows x NSV shifts, complex
  extern __shared__ __align__(128) unsigned char smem[];
  const int n_active = *Pm.n_active;
  const int panel = blockIdx.x / P;
  if (panel * NSV >= n_active) return;  // the whole cluster leaves together
  const int rank = (int)cluster_ctarank();
  const bool is_main = rank == P - 1;

  unsigned char* ctl = smem + S * STAGE_BYTES;  // 256 B of control words
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctl);  // [S <= 8]
  uint64_t* empty_bar = full_bar + 8;                     // [8]
  uint64_t* acc_ready = empty_bar + 8;                    // [2] filled by the previous CTA
  uint64_t* acc_free = acc_ready + 2;                     // [2] released by the next CTA
  uint32_t* rows_done = reinterpret_cast<uint32_t*>(acc_free + 2);
  unsigned char* hand = ctl + 256;                        // [2][HAND_BYTES]
  double2* zsh = reinterpret_cast<double2*>(hand + 2 * HAND_BYTES);
  double2** wptr = reinterpret_cast<double2**>(zsh + NSV);
  const double2** qptr = (const double2**)(wptr + NSV);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp;

  if (tid < NSV) {
    const int slot = Pm.active[panel * NSV + tid];
    wptr[tid] = Pm.Wv + (size_t)slot * Pm.n_pad;
    qptr[tid] = Pm.iter[slot] == 0 ? Pm.q0 + (size_t)Pm.batch[slot] * Pm.n_pad
                                   : (Pm.sel[slot] ? Pm.Q1 : Pm.Q0) + (size_t)slot * Pm.n_pad;
    zsh[tid] = Pm.z[slot];
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1 + 32);
      mbar_init(&empty_bar[s], CW);
    }
    mbar_init(&acc_ready[0], CW);
    mbar_init(&acc_ready[1], CW);
    mbar_init(&acc_free[0], CW);
    mbar_init(&acc_free[1], CW);
    *rows_done = 0;
    mbar_fence_init();
  }
  __syncthreads();
  cluster_arrive_wait();  // every CTA's barriers exist before anybody signals across CTAs

  const int nblk = Pm.nblk, n_pad = Pm.n_pad, n = Pm.n;
  // leading chunks of the plain solve that are pure padding (x = 0): their GEMM steps and diagonal steps are left out
  // of every CTA's step sequence (see solve_deep_body)
  const int c0 = (n_pad - n) / KC;

  if (warp == CW) {
    // ================= producer warp: this CTA's share of every block row =================
    const uint64_t pol_T = l2_policy_evict_last(), pol_X = l2_policy_evict_first();
    uint32_t known_done = 0;
    int q = 0;
    for (int phase = 0; phase < 2; ++phase) {
      for (int I = 0; I < nblk; ++I) {
        int cb, ce;
        cluster_range<P, DC>(I, rank, cb, ce);
        if (phase == 1) cb = max(cb, min(c0, ce));
        const int cp0 = (phase == 1 && I == 0) ? min(c0, CPB) : 0;  // first diagonal chunk that is not pure padding
        const int nsteps = (ce - cb) + (is_main ? CPB - cp0 : 0);
        for (int k = 0; k < nsteps; ++k, ++q) {
          const bool gemm = k < ce - cb;
          const int c = gemm ? cb + k : CPB * I + cp0 + (k - (ce - cb));
          const int stage = q % S;
          if (q >= S) mbar_wait(&empty_bar[stage], (uint32_t)(q / S - 1) & 1u);
          // What this step loads must be final: chunk c of x once main has finished block row c / 4 of this phase; the
          // right-hand side of a diagonal step of the plain solve (w) once the adjoint phase finished the mirror row.
          uint32_t need = 0;
          if (gemm) need = (uint32_t)(phase * nblk + (c >> 2) + 1);
          else if (phase == 1) need = (uint32_t)(((CPB * nblk - 1 - c) >> 2) + 1);
          if (known_done < need) {
            if (lane == 0)
              while ((known_done = ld_acquire_cluster_u32(rows_done)) < need) {
              }
            known_done = __shfl_sync(0xffffffffu, known_done, 0);
            __syncwarp();
          }
          unsigned char* sb = smem + stage * STAGE_BYTES;
          if (lane == 0) {
            mbar_arrive_expect_tx(&full_bar[stage], TILE_BYTES);
            const double2* src = (phase ? Pm.packB : Pm.packA) + tile_index(I, c) * TILE_ELEMS;
            bulk_g2s_hint(sb, src, TILE_BYTES, &full_bar[stage], pol_T);
          }
          const bool from_q = !gemm && phase == 0;
          const int mem0 = phase ? n_pad - KC * c - KC : KC * c;
#pragma unroll
          for (int seg = lane; seg < NSV * KC; seg += 32) {
            const int s = seg / KC, part = seg % KC;
            const double2* src = (from_q ? qptr[s] : (const double2*)wptr[s]) + mem0 + part;
            cp_async16_hint(sb + TILE_BYTES + s * XROW_D + ((part ^ ((s & 1) << 2)) << 4), src, pol_X);
          }
          cp_async_mbar_arrive_noinc(&full_bar[stage]);
        }
      }
    }
    cluster_arrive_wait();  // nobody leaves while a neighbour may still write into this CTA's shared memory
    return;
  }

  // ================= consumer warps =================
  double cre[NT][2], cim[NT][2];  // one 8-row DMMA tile per warp (CW = 8)

  auto gemm_step = [&](const unsigned char* sb, int flip) {
    const double2* At = reinterpret_cast<const double2*>(sb);
    const unsigned char* Xt = sb + TILE_BYTES + (lane >> 2) * XROW_D;
    const int fx = flip ^ (((lane >> 2) & 1) << 2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2 b[NT];
      const double2 a = At[(j * 8 + wm) * 32 + lane];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        b[nt] = *reinterpret_cast<const double2*>(Xt + nt * 8 * XROW_D + (((4 * j + (lane & 3)) ^ fx) << 4));
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma884(cre[nt][0], cre[nt][1], a.x, b[nt].x);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma884(cim[nt][0], cim[nt][1], a.x, b[nt].y);
      const double nai = -a.y;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma884(cre[nt][0], cre[nt][1], nai, b[nt].y);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma884(cim[nt][0], cim[nt][1], a.y, b[nt].x);
    }
  };

  // this lane's accumulator fragments inside a hand-off slot: [warp][nt][e][lane] -> 16 B each (coalesced per warp)
  auto hand_off = [&](int nt, int e) -> int { return (((wm * NT + nt) * 2 + e) * 32 + lane) * 16; };
  const uint32_t hand_local = smem_u32(hand);
  const uint32_t next_hand = rank + 1 < P ? mapa_u32(hand_local, rank + 1) : 0;
  const uint32_t next_ready = rank + 1 < P ? mapa_u32(smem_u32(acc_ready), rank + 1) : 0;
  const uint32_t prev_free = rank > 0 ? mapa_u32(smem_u32(acc_free), rank - 1) : 0;

  int q = 0, row = 0;  // local step counter (ring), local block-row counter (hand-off slots)
  for (int phase = 0; phase < 2; ++phase) {
    const int flip = phase ? (KC - 1) : 0;
    for (int I = 0; I < nblk; ++I, ++row) {
      int cb, ce;
      cluster_range<P, DC>(I, rank, cb, ce);
      if (phase == 1) cb = max(cb, min(c0, ce));
      const int cp0 = (phase == 1 && I == 0) ? min(c0, CPB) : 0;
      const int slot = row & 1;
      const uint32_t slot_parity = (uint32_t)(row >> 1) & 1u;
      // ---- accumulators: zero (first stage) or received from the previous stage ----
      if (rank == 0) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) cre[nt][0] = cre[nt][1] = cim[nt][0] = cim[nt][1] = 0.0;
      } else {
        mbar_wait_acq_cluster(&acc_ready[slot], slot_parity);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double2 v = *reinterpret_cast<const double2*>(hand + slot * HAND_BYTES + hand_off(nt, e));
            cre[nt][e] = v.x;
            cim[nt][e] = v.y;
          }
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_release(prev_free + slot * 8);  // the sender may reuse the slot
      }
      // ---- this stage's share of the GEMM steps ----
      for (int c = cb; c < ce; ++c, ++q) {
        const int stage = q % S;
        mbar_wait(&full_bar[stage], (uint32_t)(q / S) & 1u);
        gemm_step(smem + stage * STAGE_BYTES, flip);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
      }
      if (!is_main) {
        // ---- hand the partial sums to the next stage ----
        if (row >= 2) mbar_wait_acq_cluster(&acc_free[slot], (uint32_t)((row >> 1) - 1) & 1u);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) st_cluster_f64x2(next_hand + slot * HAND_BYTES + hand_off(nt, e), cre[nt][e], cim[nt][e]);
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_release(next_ready + slot * 8);
        continue;
      }
      // ---- main: the four diagonal steps of this block row ----
      for (int cp = cp0; cp < CPB; ++cp, ++q) {
        const int c = CPB * I + cp;
        const int stage = q % S;
        unsigned char* sb = smem + stage * STAGE_BYTES;
        mbar_wait(&full_bar[stage], (uint32_t)(q / S) & 1u);
        unsigned char* Xt = sb + TILE_BYTES;
        if (wm >> 1 == cp) {  // the two warps owning rows 16cp..16cp+15: rhs <- rhs - acc, in place
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int r = 8 * (wm & 1) + (lane >> 2), sl = 8 * nt + 2 * (lane & 3) + e;
              double2* pb = reinterpret_cast<double2*>(Xt + sl * XROW_D + ((r ^ flip ^ (e << 2)) << 4));
              const double2 g = *pb;
              *pb = make_double2(g.x - cre[nt][e], g.y - cim[nt][e]);
            }
        }
        consumer_sync<CTHREADS>();
        if (warp < SOLVE_THREADS / 32) {
          const int grp = tid >> 2, pl = lane & 3;
          const bool solver = grp < NSV;
          const int s = solver ? grp : 0;
          const int mem0 = phase ? n_pad - KC * c - KC : KC * c;
          double2* xrow = reinterpret_cast<double2*>(Xt + s * XROW_D);
          const int fs = flip ^ ((s & 1) << 2);
          const double2* At = reinterpret_cast<const double2*>(sb);
          double2 zt = zsh[s];
          if (phase == 0) zt.y = -zt.y;
          const int p0 = MB * I + KC * cp;
          double2 b[4], dinv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = 4 * i + pl, m = KC * cp + r;
            b[i] = xrow[r ^ fs];
            double2 d = At[((r >> 2) * 8 + (m >> 3)) * 32 + (m & 7) * 4 + (r & 3)];
            d.x -= zt.x;
            d.y -= zt.y;
            dinv[i] = cinv_fixed(d);
          }
#pragma unroll
          for (int r = 0; r < KC; ++r) {
            const int ir = r >> 2, pr = r & 3;
            double2 x = cmul_fixed(b[ir], dinv[ir]);
            const int op = phase ? n_pad - 1 - (p0 + r) : p0 + r;
            if (op >= n) x = make_double2(0.0, 0.0);
            if (pl == pr) b[ir] = x;
            const int src = (lane & ~3) | pr;
            x.x = __shfl_sync(0xffffffffu, x.x, src);
            x.y = __shfl_sync(0xffffffffu, x.y, src);
#pragma unroll
            for (int i = ir; i < 4; ++i) {
              const int r2 = 4 * i + pl;
              if (r2 > r) {
                const int m2 = KC * cp + r2;
                const double2 l = At[((r >> 2) * 8 + (m2 >> 3)) * 32 + (m2 & 7) * 4 + (r & 3)];
                cmsub(b[i], l, x);
              }
            }
          }
          if (solver) {
            double2* op_ = wptr[s] + mem0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = 4 * i + pl;
              xrow[r ^ fs] = b[i];
              op_[r ^ flip] = b[i];
            }
          }
          if (cp == CPB - 1) __threadfence();  // the row's x must be visible device-wide before rows_done moves
        }
        consumer_sync<CTHREADS>();
        if (wm >> 1 == cp) {  // this chunk's rows are solved: their accumulators start the next block row from zero
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) cre[nt][0] = cre[nt][1] = cim[nt][0] = cim[nt][1] = 0.0;
        }
        if (wm >= 2 * (cp + 1)) gemm_step(sb, flip);  // rows below the chunk
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
      }
      // ---- main: publish "block row done" to every CTA of the cluster (its own producer warp included) ----
      if (tid == 0) {
        const uint32_t done = (uint32_t)(phase * nblk + I + 1);
        const uint32_t local = smem_u32(rows_done);
#pragma unroll
        for (int r = 0; r < P; ++r) st_release_cluster_u32(mapa_u32(local, r), done);
      }
    }
  }
  cluster_arrive_wait();
}

// ---- host side ------------------------------------------------------------------------------------------------------
// ring depth per width, cost of main's diagonal steps in GEMM-step units (r02 phase timings: td / tg of the sparse kernel)
constexpr int CL_S8 = 8, CL_S16 = 6, CL_S32 = 4;
constexpr int CL_DC8 = 18, CL_DC16 = 13, CL_DC32 = 9;

// how many clusters of P CTAs the device can run at once (clusters need P free SMs inside ONE GPC, so this is less than
// num_sms / P); queried once per process with the occupancy API: g_max_clusters[width index][P index]
static int g_max_clusters[3][2] = {{0, 0}, {0, 0}, {0, 0}};
static bool g_cluster_broken = false;  // set when a cluster launch was refused: fall back to the single-CTA kernels

template <typename K>
inline int query_max_clusters(K* kernel, int P, int smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(P * 64, 1, 1);
  cfg.blockDim = dim3(deep_threads(8), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = P;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

inline int cluster_set_attrs() {
  cudaError_t e = cudaSuccess;
  auto set = [&](auto* k, int bytes, int wi, int pi, int P) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) g_max_clusters[wi][pi] = query_max_clusters(k, P, bytes);
  };
  set(solve_cluster_kernel<8, CL_S8, 2, CL_DC8>, cluster_smem<8, CL_S8, 2>(), 0, 0, 2);
  set(solve_cluster_kernel<8, CL_S8, 4, CL_DC8>, cluster_smem<8, CL_S8, 4>(), 0, 1, 4);
  set(solve_cluster_kernel<16, CL_S16, 2, CL_DC16>, cluster_smem<16, CL_S16, 2>(), 1, 0, 2);
  set(solve_cluster_kernel<16, CL_S16, 4, CL_DC16>, cluster_smem<16, CL_S16, 4>(), 1, 1, 4);
  set(solve_cluster_kernel<32, CL_S32, 2, CL_DC32>, cluster_smem<32, CL_S32, 2>(), 2, 0, 2);
  set(solve_cluster_kernel<32, CL_S32, 4, CL_DC32>, cluster_smem<32, CL_S32, 4>(), 2, 1, 4);
  return e == cudaSuccess ? 0 : -1;
}

template <int P>
inline void launch_cluster(const SolveParams& sp, int width, int panels, cudaStream_t stream) {
  switch (width) {
    case 32: solve_cluster_kernel<32, CL_S32, P, CL_DC32><<<panels * P, deep_threads(8), cluster_smem<32, CL_S32, P>(), stream>>>(sp); break;
    case 16: solve_cluster_kernel<16, CL_S16, P, CL_DC16><<<panels * P, deep_threads(8), cluster_smem<16, CL_S16, P>(), stream>>>(sp); break;
    default: solve_cluster_kernel<8, CL_S8, P, CL_DC8><<<panels * P, deep_threads(8), cluster_smem<8, CL_S8, P>(), stream>>>(sp); break;
  }
}

// Solve launch of a tick: n_active is an upper bound of the device-side count (panels past it exit at once).
//   many panels                       -> deep-ring kernel, 2 CTAs per SM
//   <= one panel per SM               -> sparse deep-ring kernel (8 consumer warps)
//   few enough panels for P SMs each  -> cluster pipeline over P = 2 / 4 SMs per panel (bit-identical results)
// Measured at n = 4000 (profiles/r02_cluster_check.log): a lone 8-wide panel takes 6.13 / 3.37 / 2.02 ms with
// P = 1 / 2 / 4; the pipeline needs enough block rows to fill (n = 1000: 0.56 / 0.40 / 0.38 ms, n <= 333: no gain).
// PSA_CLUSTER=1|2|4 (experiments) caps the cluster size.
inline void launch_solve_auto(const SolveParams& sp, int n_active, int num_sms, cudaStream_t stream) {
  if (solve_flavor() != 0 && sp.nblk >= 12) {
    static const int widths[3] = {8, 16, 32};
    // launch time relative to the sparse single-CTA kernel of the same width
    const double f2 = 0.55, f4 = sp.nblk >= 32 ? 0.32 : 0.9;
    static const double t1[3] = {6.13, 10.4, 18.9};  // ms at n = 4000: only the ratios between widths matter here
    const char* ec = getenv("PSA_CLUSTER");
    const int cap = ec ? atoi(ec) : 4;
    const char* ew = getenv("PSA_PANEL_WIDTH");
    const int forced = ew ? atoi(ew) : 0;
    double best = 1e300;
    int bw = 0, bp = 1;
    for (int i = 0; i < 3; ++i) {
      if (forced && widths[i] != forced) continue;
      const int panels = (n_active + widths[i] - 1) / widths[i];
      if (panels > num_sms) continue;  // not a sparse launch at this width
      double c = t1[i];
      int P = 1;
      if (cap >= 4 && panels <= g_max_clusters[i][1] && t1[i] * f4 < c) c = t1[i] * f4, P = 4;
      if (cap >= 2 && panels <= g_max_clusters[i][0] && t1[i] * f2 < c) c = t1[i] * f2, P = 2;
      if (c < best) best = c, bw = widths[i], bp = P;
    }
    if (bw && bp > 1 && !g_cluster_broken) {
      const int panels = (n_active + bw - 1) / bw;
      if (bp == 4) launch_cluster<4>(sp, bw, panels, stream);
      else launch_cluster<2>(sp, bw, panels, stream);
      if (cudaPeekAtLastError() == cudaSuccess) return;
      // a cluster launch can be refused (e.g. a partitioned device): same result without it, just slower
      cudaGetLastError();
      g_cluster_broken = true;
    }
  }
  launch_solve(sp, n_active, num_sms, stream);
}

}  // namespace psa
