// Square (:sq_lanc) path kernels: multi-shift blocked triangular solves on FP64 tensor cores.
//
// Replaces, for hundreds of shifts at once, the two BLAS ztrsv calls per Lanczos iteration at
// /root/reference/src/compute.jl:630  (`v = F1 \ (F1' \ q)`), where F1 = T - z I differs between grid points
// only on its diagonal (compute.jl:405,595).  T is shared by all shifts, so the off-diagonal block updates
// are a dense contraction (complex GEMM as 4 real DMMA per 8x8x4 block); only the 16x16 diagonal blocks are
// per-shift.  See DESIGN.md for the layout and the step schedule.
#pragma once
#include <cstring>

#include "psa_common.cuh"

namespace psa {

// ---- tiling constants -------------------------------------------------------------------------------
constexpr int MB = 64;                        // rows per block row (M tile)
constexpr int KC = 16;                        // contraction chunk (one pipeline step)
constexpr int NS = 32;                        // widest panel (shifts per CTA)
constexpr int APAD = 32;                      // active lists are padded to a multiple of every panel width (8, 16, 32)
constexpr int CPB = MB / KC;                  // chunks per block = 4
constexpr int TILE_ELEMS = MB * KC;           // complex elements per packed T tile
constexpr int TILE_BYTES = TILE_ELEMS * 16;   // 16 KiB
constexpr int XROW_BYTES = KC * 16 + 64;      // 320: X smem row stride (64 B pad -> conflict-free B-fragment loads)
constexpr int STAGES = 2;                     // <= 4 (producer/consumer distance rule)
constexpr int SOLVE_THREADS = 128;            // 4 warps: one per SM sub-partition, each owning 16 rows x all shifts
// per panel width NSV (32, 16 or 8 shifts per CTA): X tile NSV rows, ring of STAGES (T tile + X tile) stages,
// then mbarriers (64 B reserved) + shifts + 2 pointer tables
__host__ __device__ constexpr int stage_bytes(int nsv) { return TILE_BYTES + nsv * XROW_BYTES; }
__host__ __device__ constexpr int solve_smem(int nsv) { return STAGES * stage_bytes(nsv) + 64 + nsv * 16 + nsv * 8 * 2; }
static_assert(STAGES * 8 <= 64, "mbarrier block");
constexpr int HIST_MIN = 128;                 // smallest per-slot alpha/beta history stride (maxit <= stride - 2)
constexpr int HIST_MAX = 4096;                // largest supported stride (maxit <= 4094)

__host__ __device__ inline long long tiles_per_phase(int nblk) { return 2LL * nblk * (nblk + 1); }
// tile (I, c), c in [0, 4I+4): offset in tiles within one phase stream
__host__ __device__ inline long long tile_index(int I, int c) { return 2LL * I * (I + 1) + c; }

// ---- one-time packing of T into DMMA A-fragment order -----------------------------------------------
// Processing index p in [0, n_pad).  Phase A (adjoint, L_A = T^H, forward order): orig(p) = p.
// Phase B (T upper, processed bottom-up): orig(p) = n_pad-1-p.  Both become "lower triangular, forward".
//   L_A[p][p'] = conj(T[orig(p')][orig(p)]),   L_B[p][p'] = T[orig(p)][orig(p')],   p' <= p, else 0.
// Padding (orig >= n): identity.  Tile (I,c) holds rows 64I.., columns 16c.. as fragments:
//   element (m, kk) -> double2 index ((kk>>2)*8 + (m>>3))*32 + (m&7)*4 + (kk&3).
__global__ void pack_kernel(const double2* __restrict__ T, long long ldt, int n, int n_pad, int nblk,
                            double2* __restrict__ packA, double2* __restrict__ packB) {
  const long long ntiles = tiles_per_phase(nblk);
  for (long long tile = blockIdx.x; tile < 2 * ntiles; tile += gridDim.x) {
    const int phase = tile >= ntiles;
    const long long t = phase ? tile - ntiles : tile;
    // invert tile_index: I = largest with 2I(I+1) <= t
    int I = (int)((sqrt(1.0 + 2.0 * (double)t) - 1.0) * 0.5);
    while (tile_index(I + 1, 0) <= t) ++I;
    while (tile_index(I, 0) > t) --I;
    const int c = (int)(t - tile_index(I, 0));
    double2* out = (phase ? packB : packA) + t * TILE_ELEMS;
    for (int e = threadIdx.x; e < TILE_ELEMS; e += blockDim.x) {
      int m, kk;
      if (phase == 0) {  // reads run down a column of T: kk fastest
        kk = e & 15;
        m = e >> 4;
      } else {  // reads run down a column of T: m fastest (orig row = n_pad-1-p)
        m = e & 63;
        kk = e >> 6;
      }
      const int p = MB * I + m, pp = KC * c + kk;
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

__global__ void diag_kernel(const double2* __restrict__ T, long long ldt, int n, double2* __restrict__ diag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) diag[i] = T[(long long)i * ldt + i];
}

// Cost predictor for longest-job-first scheduling: inverse Lanczos converges fast where the resolvent norm is
// large (near the spectrum) and slowly far from it, so points are queued by decreasing distance to the nearest
// eigenvalue (diag T).  Only the ORDER of the work changes; every point's result is independent of it.
__global__ void __launch_bounds__(256) predict_kernel(int P, int rows_local, int row0, int row_step,
                                                       const double* __restrict__ x, const double* __restrict__ y,
                                                       const double2* __restrict__ diag, int n,
                                                       float* __restrict__ key, int* __restrict__ val,
                                                       unsigned* __restrict__ key_max_bits) {
  __shared__ double2 ev[256];
  const int p = blockIdx.x * 256 + threadIdx.x;
  double zx = 0.0, zy = 0.0;
  if (p < P) {
    const int jr = p % rows_local, k = p / rows_local;
    zx = x[k];
    zy = y[row0 + jr * row_step];
  }
  double best = 1e300;
  for (int base = 0; base < n; base += 256) {
    __syncthreads();
    if (base + threadIdx.x < n) ev[threadIdx.x] = diag[base + threadIdx.x];
    __syncthreads();
    const int cnt = min(256, n - base);
    for (int i = 0; i < cnt; ++i) {
      const double dx = ev[i].x - zx, dy = ev[i].y - zy;
      best = fmin(best, dx * dx + dy * dy);
    }
  }
  if (p < P) {
    key[p] = (float)best;
    val[p] = p;
    if (key_max_bits) atomicMax(key_max_bits, __float_as_uint((float)best));  // non-negative floats order like their bits
  }
}

// ---- dynamic queue order -----------------------------------------------------------------------------------
// The static predictor (distance to the spectrum) is only a heuristic: for strongly non-normal matrices the iteration
// count follows the resolvent norm, not the distance.  But iteration counts are spatially smooth, and the grid itself
// tells us what they are as it is computed: a coarse LATTICE of points (stride s in both directions) is queued first,
// and every tick the not-yet-started points are re-keyed with the progress of the lattice points around them --
// iterations so far for a running one, the final count for a finished one -- so the neighbourhood of slow points is
// scheduled early and the grid does not end in a tail of sparse ticks.  Ties are broken by the static key.
// Only the ORDER of the work changes; every point's result is independent of it.
struct RekeyParams {
  int P, rows_local, lx, s;      // local points, local rows, columns, lattice stride
  const int* prog;               // per point: 0 = not started, else iterations so far (final count when finished)
  const float* key_static;
  const unsigned* key_max_bits;
  float* key;                    // out
  int* val;                      // out: point index
  int* state;                    // [0] started so far, [4] <- [0]: queue positions restart at 0 after the re-sort
};

__global__ void __launch_bounds__(256) rekey_kernel(RekeyParams R) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p == 0) R.state[4] = R.state[0];
  if (p >= R.P) return;
  R.val[p] = p;
  if (R.prog[p] > 0) {  // started (or finished): sorts behind every waiting point
    R.key[p] = -1.0f;
    return;
  }
  const int jr = p % R.rows_local, k = p / R.rows_local, s = R.s;
  const float stat = 0.999f * R.key_static[p] / fmaxf(__uint_as_float(*R.key_max_bits), 1e-30f);
  if (jr % s == 0 && k % s == 0) {  // a lattice point itself: first in line
    R.key[p] = 1.0e9f + stat;
    return;
  }
  const int j0 = jr / s * s, k0 = k / s * s;
  const int j1 = min(j0 + s, (R.rows_local - 1) / s * s), k1 = min(k0 + s, (R.lx - 1) / s * s);
  const int m = max(max(R.prog[k0 * R.rows_local + j0], R.prog[k0 * R.rows_local + j1]),
                    max(R.prog[k1 * R.rows_local + j0], R.prog[k1 * R.rows_local + j1]));
  R.key[p] = (float)m + stat;
}

// ---- the solve kernel ---------------------------------------------------------------------------------
struct SolveParams {
  const double2* packA;
  const double2* packB;
  int n, n_pad, nblk;
  const int* active;   // slot ids, padded to a multiple of APAD with the scratch slot
  const int* n_active; // device scalar
  double2* Q0;         // [W+1][n_pad]
  double2* Q1;
  double2* Wv;
  const int* sel;      // which of Q0/Q1 currently holds q
  const double2* z;    // per-slot shift
  const int* iter;     // per-slot completed Lanczos iterations: 0 = fresh slot, q is still the start vector
  const int* batch;    // per-slot start-vector index (row batch, compute.jl:199-208)
  const double2* q0;   // [nbatch][n_pad] start vectors, zero padded
  int pair;            // fuse pairs of GEMM steps whose stages have both landed (PSA_PAIR=0 switches it off: A/B timing)
};

struct StepCursor {
  int phase, I, c;
  __device__ __forceinline__ void advance() {
    if (++c == CPB * I + CPB) {
      c = 0;
      ++I;
    }
  }
};

// One CTA = one panel of NSV shifts (32 in steady state; 16 / 8 when few shifts are active, so that the panels
// still cover every SM); 4 warps, one per SM sub-partition, warp w owning rows 16w..16w+15 of the block row for all
// NSV shifts; up to 4 CTAs per SM, i.e. four independent barrier domains whose pipeline bubbles overlap
// (measured: 33.0 TFLOP/s against 31.9 for the earlier 8-warp / 64-shift CTAs, profiles/r01_panel_sweep.txt).
// Left-looking block forward substitution (both phases, see pack_kernel): for block row I
//   acc = sum_{c < 4I} L[I][c] * X[c]              (GEMM steps: T tile + X tile through the bulk-copy ring)
//   for cp = 0..3: rows 16cp..16cp+15:  x = (L_dd - z)^{-1} (rhs - acc)   (diag step, per-shift, in registers)
//                  acc += L[I][4I+cp] * x           (rows below, same DMMA code)
template <int NSV>
__global__ void __launch_bounds__(SOLVE_THREADS, 4) solve_kernel(SolveParams P) {
  constexpr int SW = NSV;      // shifts per warp: every warp sees the whole panel
  constexpr int NT = SW / 8;   // 8-wide DMMA column tiles per warp
  constexpr int STAGE_BYTES = stage_bytes(NSV);
  extern __shared__ __align__(128) unsigned char smem[];
  const int n_active = *P.n_active;
  const int panel = blockIdx.x;
  if (panel * NSV >= n_active) return;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);  // 64 B reserved
  double2* zsh = reinterpret_cast<double2*>(smem + STAGES * STAGE_BYTES + 64);    // 16 B aligned
  double2** wptr = reinterpret_cast<double2**>(zsh + NSV);
  const double2** qptr = (const double2**)(wptr + NSV);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp, wn = 0;  // warp = row group = SM sub-partition

  if (tid < NSV) {
    const int slot = P.active[panel * NSV + tid];
    wptr[tid] = P.Wv + (size_t)slot * P.n_pad;
    // a fresh slot's q IS the start vector of its row batch: read it in place (lanczos_kernel materialises it)
    qptr[tid] = P.iter[slot] == 0 ? P.q0 + (size_t)P.batch[slot] * P.n_pad
                                  : (P.sel[slot] ? P.Q1 : P.Q0) + (size_t)slot * P.n_pad;
    zsh[tid] = P.z[slot];
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1 + SOLVE_THREADS);  // thread 0 expect_tx + one cp.async arrival per thread
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int nblk = P.nblk, n_pad = P.n_pad, n = P.n;
  const int steps_per_phase = (int)tiles_per_phase(nblk);
  const int total_steps = 2 * steps_per_phase;

  // producer side: issue the loads of one step into its ring stage; called by ALL threads.
  //  * T tile: one 16 KiB bulk copy through the TMA engine (thread 0), completing by bytes on the stage mbarrier;
  //  * X tile: NSV rows of 256 B, one row per shift at a scattered address.  A bulk copy is a uniform-datapath
  //    instruction (UBLKCP), so per-lane bulk copies are serialised (~80 clk each); instead every thread issues
  //    16-byte cp.async copies (LDGSTS) for its share of the rows and hooks their completion onto the same
  //    mbarrier (one pre-counted arrival per thread).
  const uint64_t pol_T = l2_policy_evict_last(), pol_X = l2_policy_evict_first();
  auto issue = [&](const StepCursor& pc, int stage) {
    unsigned char* sb = smem + stage * STAGE_BYTES;
    const bool needX = pc.c < CPB * pc.I;
    if (tid == 0) {
      mbar_arrive_expect_tx(&full_bar[stage], TILE_BYTES);
      const double2* src = (pc.phase ? P.packB : P.packA) + tile_index(pc.I, pc.c) * TILE_ELEMS;
      bulk_g2s_hint(sb, src, TILE_BYTES, &full_bar[stage], pol_T);
    }
    if (needX) {
      const int mem0 = pc.phase ? n_pad - KC * pc.c - KC : KC * pc.c;
#pragma unroll
      for (int seg = tid; seg < NSV * KC; seg += SOLVE_THREADS) {  // 16-byte segments: row = seg / KC
        const int s = seg / KC, part = seg % KC;
        cp_async16_hint(sb + TILE_BYTES + s * XROW_BYTES + part * 16, wptr[s] + mem0 + part, pol_X);
      }
    }
    cp_async_mbar_arrive_noinc(&full_bar[stage]);  // arrives once this thread's copies (if any) have landed
  };

  double cre[2][NT][2], cim[2][NT][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b) cre[a][b][0] = cre[a][b][1] = cim[a][b][0] = cim[a][b][1] = 0.0;

  auto gemm_step = [&](const unsigned char* sb, int flip) {
    const double2* At = reinterpret_cast<const double2*>(sb);
    const unsigned char* Xt = sb + TILE_BYTES + (SW * wn + (lane >> 2)) * XROW_BYTES;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2 a[2], b[NT];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) a[mt] = At[(j * 8 + 2 * wm + mt) * 32 + lane];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        b[nt] = *reinterpret_cast<const double2*>(Xt + nt * 8 * XROW_BYTES + (((4 * j + (lane & 3)) ^ flip) << 4));
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

  StepCursor cons{0, 0, 0}, prod{0, 0, 0};
  auto advance_prod = [&]() {
    prod.advance();
    if (prod.I == nblk) prod = StepCursor{1, 0, 0};
  };
  for (int s = 0; s < STAGES && s < total_steps; ++s) {
    issue(prod, s);
    advance_prod();
  }

  // optional per-phase cycle counters (build with -DPSA_TIMING; prints from the first and last CTA)
#ifdef PSA_TIMING
  long long tm_wait = 0, tm_gemm = 0, tm_diag = 0, tm_sync = 0, tm0;
#define PSA_T0() tm0 = clock64()
#define PSA_T1(acc) acc += clock64() - tm0
#else
#define PSA_T0()
#define PSA_T1(acc)
#endif
  for (int t = 0; t < total_steps; ++t) {
    const int stage = t % STAGES;
    const uint32_t parity = (uint32_t)(t / STAGES) & 1u;
    unsigned char* sb = smem + stage * STAGE_BYTES;
    const int flip = cons.phase ? (KC - 1) : 0;
    PSA_T0();
    mbar_wait(&full_bar[stage], parity);
    PSA_T1(tm_wait);

    if (cons.c < CPB * cons.I) {
      PSA_T0();
      gemm_step(sb, flip);
      PSA_T1(tm_gemm);
    } else {
      PSA_T0();
      const int cp = cons.c - CPB * cons.I;  // diagonal sub-chunk 0..3
      unsigned char* Xt = sb + TILE_BYTES;
      if (wm == cp) {
        // this warp owns rows 16cp..16cp+15 of the block row: hand its accumulators over through the stage's
        // (unused) X buffer, transposed: element (row r, shift sl) -> Xt[sl][r ^ flip]
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int r = 8 * mt + (lane >> 2), sl = 8 * nt + 2 * (lane & 3) + e;
              *reinterpret_cast<double2*>(Xt + sl * XROW_BYTES + ((r ^ flip) << 4)) =
                  make_double2(cre[mt][nt][e], cim[mt][nt][e]);
              cre[mt][nt][e] = 0.0;
              cim[mt][nt][e] = 0.0;
            }
      }
      __syncthreads();
      // Per-shift 16x16 triangular solve (L_dd - z~ I) x = rhs - acc, spread over the whole CTA: four lanes per
      // shift (lane p owns rows p, p+4, p+8, p+12), pivots broadcast with shuffles.  The FP64 pipe is saturated by
      // the co-resident CTAs' DMMAs, so what matters here is the number of FP64 instructions per thread.
      {
        constexpr int GROUPS = SOLVE_THREADS / 4;          // shifts solved concurrently (32)
        const int grp = tid >> 2, pl = lane & 3;            // shift group and lane-in-group
        const bool solver = grp < NSV;
        const int s = solver ? grp : 0;
        static_assert(NSV <= GROUPS, "one 4-lane group per shift");
        const int mem0 = cons.phase ? n_pad - KC * cons.c - KC : KC * cons.c;
        const double2* rp = (cons.phase ? (const double2*)wptr[s] : qptr[s]) + mem0;
        double2* xrow = reinterpret_cast<double2*>(Xt + s * XROW_BYTES);
        const double2* At = reinterpret_cast<const double2*>(sb);
        double2 zt = zsh[s];
        if (cons.phase == 0) zt.y = -zt.y;  // adjoint system: diagonal conj(t_ii - z)
        const int p0 = MB * cons.I + KC * cp;
        double2 b[4], dinv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 4 * i + pl, m = KC * cp + r;
          const double2 g = rp[r ^ flip], a = xrow[r ^ flip];
          b[i] = make_double2(g.x - a.x, g.y - a.y);
          double2 d = At[((r >> 2) * 8 + (m >> 3)) * 32 + (m & 7) * 4 + (r & 3)];
          d.x -= zt.x;
          d.y -= zt.y;
          dinv[i] = cinv_fixed(d);
        }
#pragma unroll
        for (int r = 0; r < KC; ++r) {  // forward substitution in processing order; pivot r lives on lane r & 3
          const int ir = r >> 2, pr = r & 3;
          double2 x = cmul_fixed(b[ir], dinv[ir]);
          const int op = cons.phase ? n_pad - 1 - (p0 + r) : p0 + r;
          if (op >= n) x = make_double2(0.0, 0.0);
          if (pl == pr) b[ir] = x;
          const int src = (lane & ~3) | pr;
          x.x = __shfl_sync(0xffffffffu, x.x, src);
          x.y = __shfl_sync(0xffffffffu, x.y, src);
#pragma unroll
          for (int i = ir; i < 4; ++i) {
            const int r2 = 4 * i + pl;  // this lane's row in group i
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
            xrow[r ^ flip] = b[i];
            op_[r ^ flip] = b[i];
          }
        }
        fence_proxy_async();  // belt and braces: later block rows re-read these x with asynchronous copies
      }
      __syncthreads();
      if (wm > cp) gemm_step(sb, flip);
      PSA_T1(tm_diag);
    }
    PSA_T0();
    __syncthreads();  // every warp is done with this stage: refill it
    if (t + STAGES < total_steps) {
      issue(prod, stage);
      advance_prod();
    }
    PSA_T1(tm_sync);  // BAR.SYNC is deferred-blocking: the wait shows up at the first dependent instruction
    cons.advance();
    if (cons.I == nblk) cons = StepCursor{1, 0, 0};
  }
#ifdef PSA_TIMING
  if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
    printf("[timing] cta %d warp %d: load-wait %lld gemm %lld diag %lld barrier+issue %lld clk, %d steps\n", (int)blockIdx.x, warp,
           tm_wait, tm_gemm, tm_diag, tm_sync, total_steps);
#endif
}

// ---- the decoupled ("deep ring") solve kernel ----------------------------------------------------------------
// Same arithmetic, same step order and therefore bit-identical results as solve_kernel, different choreography:
//   * at most 2 CTAs per SM, so each CTA affords a ring of S = 4..6 stages (solve_kernel: 2) and ~200 registers;
//   * a fifth warp is the PRODUCER: it alone issues the loads (one bulk copy for the T tile, its lanes' share of
//     16-byte cp.async copies for the X rows), so the four consumer warps spend their issue slots on DMMAs only;
//   * no CTA-wide barrier per step: a consumer warp that has finished a step releases the stage on a per-stage
//     "empty" mbarrier (4 arrivals) and moves on; the producer refills a stage once all four warps released it, up to
//     S steps ahead, so consumer warps may drift apart by a whole ring;
//   * the right-hand side rows of a diagonal step arrive with the stage (they replace the unused X tile) instead of
//     being fetched from global memory inside the serial part of the step;
//   * a load is only issued once the step that PRODUCES its data has been released by all consumers (dep_step): with
//     a deep ring the first block rows of a phase, and the hand-over of w from the adjoint to the plain solve, are
//     closer than S steps.
// Only the diagonal steps (3 % of the steps at n = 4000) still synchronise the four consumer warps (named barrier 1).
// CW consumer warps (4: each owns 16 rows of the block row; 8: 8 rows each, i.e. four DMMA-issuing warps per SM
// sub-partition at 2 CTAs per SM) + 1 producer warp.
__host__ __device__ constexpr int deep_threads(int cw) { return cw * 32 + 32; }

// X tile rows are 256 B (16 complex), UNPADDED, with a one-bit swizzle instead: the 16-byte element e of row s sits at
// position e ^ (4 * (s & 1)).  A quarter-warp of a B-fragment LDS.128 covers rows (s, s+1) x four consecutive elements:
// the two rows then use complementary halves of the 32 banks -- conflict-free like the 320-byte padded rows of
// solve_kernel, with 8 % smaller stages (deeper rings in the same shared memory).
// (PAD = true keeps the 320-byte padded rows instead: measured 1 % faster for the full 32-wide launch -- 72.1 against
// 73.0 ms, L2 hit rate 63.4 against 61.1 % -- so the full variant pads and the sparse / cluster variants, which want the
// deepest possible ring, swizzle.)
constexpr int XROW_D = KC * 16;
__host__ __device__ constexpr int deep_xrow(bool pad) { return pad ? XROW_BYTES : XROW_D; }
__host__ __device__ constexpr int deep_stage_bytes(int nsv, bool pad = false) { return TILE_BYTES + nsv * deep_xrow(pad); }
// NP panels per CTA share the T tile of a stage; every panel has its own X tile, shifts and pointer tables
template <int NSV, int S, bool PAD = false, int NP = 1>
__host__ __device__ constexpr int deep_smem() {
  return S * (TILE_BYTES + NP * NSV * deep_xrow(PAD)) + 256 + NP * (NSV * 16 + NSV * 8 * 2);
}

template <int THREADS>
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }

template <int NSV, int S, int CW, bool PAD, int NP = 1>
__device__ __forceinline__ void solve_deep_body(const SolveParams& P) {
  constexpr int XROW = deep_xrow(PAD);     // X tile row stride
  constexpr int SWZ = PAD ? 0 : 4;         // XOR applied to the element index of odd rows
  static_assert(S >= 3 && S <= 8 && (CW == 4 || CW == 8), "ring geometry");
  static_assert(NP == 1 || CW == 4, "several panels per CTA: lane-solve variant only");
  static_assert((2 + NP) * S * 8 <= 256, "mbarrier block");
  constexpr int NT = NSV / 8;
  constexpr int MT = 8 / CW;              // 8-row DMMA tiles per consumer warp
  constexpr int CTHREADS = CW * 32;       // consumer threads of one panel
  constexpr int NCW = CW * NP;            // consumer warps of the CTA; the producer is warp NCW
  constexpr int XTILE = NSV * XROW;       // one panel's X tile
  constexpr int STAGE_BYTES = TILE_BYTES + NP * XTILE;
  // CW = 4: "lane-solve" diagonal steps (see the consumer loop).  Their right-hand side tile is laid out row-major by
  // ROW, D[r][s ^ (r & 1)] (r = row of the chunk in processing order, s = shift of the panel, 16 B each), so that a
  // lane that owns one shift reads / writes its 16 rows without bank conflicts and the tile serves as the B operand of
  // the update of the rows below as it stands.
  constexpr bool LANE_SOLVE = (CW == 4);
  constexpr bool PAIR_STEPS = true;
  constexpr int DROW = NSV * 16;
  extern __shared__ __align__(128) unsigned char smem[];
  const int n_active = *P.n_active;
  const int panel = blockIdx.x * NP;  // first panel of this CTA
  if (panel * NSV >= n_active) return;
  const int np_here = min(NP, (n_active - panel * NSV + NSV - 1) / NSV);  // panels of this CTA that hold active shifts

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);  // 256 B reserved: full[S], empty[S], xready[S][NP]
  uint64_t* empty_bar = full_bar + S;
  uint64_t* xready_bar = empty_bar + S;
  double2* zsh = reinterpret_cast<double2*>(smem + S * STAGE_BYTES + 256);
  double2** wptr = reinterpret_cast<double2**>(zsh + NP * NSV);
  const double2** qptr = (const double2**)(wptr + NP * NSV);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pi = NP == 1 ? 0 : warp / CW;   // consumer warps: panel of the CTA this warp works for
  const int wq = NP == 1 ? warp : warp % CW;
  int wm = wq;  // row group of the block row this warp owns (lane-solve: rotates with the block row)
  // byte offset of this warp's X tile within a stage.  Spelled with selects: for the product pi * XTILE ptxas 12.9 emits
  // IDP.2A.LO.U16.U8, which a B200 rejects at run time ("an illegal instruction was encountered").
  const int pxo = (pi >= 1 ? XTILE : 0) + (pi >= 2 ? XTILE : 0) + (pi >= 3 ? XTILE : 0);
  static_assert(NP <= 4, "pxo");

  if (tid < np_here * NSV) {
    const int slot = P.active[panel * NSV + tid];
    wptr[tid] = P.Wv + (size_t)slot * P.n_pad;
    qptr[tid] = P.iter[slot] == 0 ? P.q0 + (size_t)P.batch[slot] * P.n_pad
                                  : (P.sel[slot] ? P.Q1 : P.Q0) + (size_t)slot * P.n_pad;
    zsh[tid] = P.z[slot];
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1 + 32);               // producer lane 0 expect_tx + one cp.async arrival per producer lane
      mbar_init(&empty_bar[s], CW * np_here);        // one arrival per consumer warp
#pragma unroll
      for (int k = 0; k < NP; ++k) mbar_init(&xready_bar[s * NP + k], 1);  // lane-solve diagonal steps: the owner warp publishes x
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp < NCW && pi >= np_here) return;  // consumer warps of a panel without active shifts

  const int nblk = P.nblk, n_pad = P.n_pad, n = P.n;
  const int steps_per_phase = (int)tiles_per_phase(nblk);
  const int total_steps = 2 * steps_per_phase;
  // The plain solve runs bottom-up, so the rows padded up to n_pad come FIRST there: its leading c0 chunks are pure
  // padding (x = 0 by construction).  Steps that would only multiply by those zeros, or solve for them, are skipped:
  // they keep their slot in the ring (the barriers are exercised, nothing is loaded or computed).  n = 4000: 126 of
  // 16 128 steps.  Adding exact zeros changes nothing, so variants that skip and variants that do not agree bit for bit.
  const int c0 = (n_pad - n) / KC;

  if (warp == NCW) {
    // ================= producer warp =================
    const uint64_t pol_T = l2_policy_evict_last(), pol_X = l2_policy_evict_first();
    StepCursor pc{0, 0, 0};
    for (int q = 0; q < total_steps; ++q) {
      const int stage = q % S;
      const bool gemm = pc.c < CPB * pc.I;
      // the stage is free once every consumer warp released step q - S
      if (q >= S) mbar_wait(&empty_bar[stage], (uint32_t)(q / S - 1) & 1u);
      if (pc.phase == 1 && pc.c < c0) {  // skipped step: complete the stage's full barrier without loading anything
        mbar_arrive(&full_bar[stage]);
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        pc.advance();
        if (pc.I == nblk) pc = StepCursor{1, 0, 0};
        continue;
      }
      // the data this step loads is final once the step that produced it has been released by every consumer warp:
      //   GEMM step: X chunk c, produced by the diagonal step (c/4, c) of the same phase;
      //   diagonal step of the plain solve: its right-hand side w, produced by the adjoint phase (mirror chunk);
      //   diagonal step of the adjoint solve: q, an input.
      int dep = -1;
      if (gemm) dep = pc.phase * steps_per_phase + (int)tile_index(pc.c >> 2, pc.c);
      else if (pc.phase == 1) {
        const int cA = CPB * nblk - 1 - pc.c;
        dep = (int)tile_index(cA >> 2, cA);
      }
      if (dep >= 0 && dep > q - S) mbar_wait(&empty_bar[dep % S], (uint32_t)(dep / S) & 1u);  // (older steps: implied by the wait above)
      unsigned char* sb = smem + stage * STAGE_BYTES;
      if (lane == 0) {
        mbar_arrive_expect_tx(&full_bar[stage], TILE_BYTES);
        const double2* src = (pc.phase ? P.packB : P.packA) + tile_index(pc.I, pc.c) * TILE_ELEMS;
        bulk_g2s_hint(sb, src, TILE_BYTES, &full_bar[stage], pol_T);
      }
      const bool from_q = !gemm && pc.phase == 0;
      const int mem0 = pc.phase ? n_pad - KC * pc.c - KC : KC * pc.c;
      const bool dlay = LANE_SOLVE && !gemm;       // right-hand side of a lane-solve diagonal step: D layout
      const int flipP = pc.phase ? (KC - 1) : 0;   // memory order -> processing order of the plain solve
      for (int k = 0; k < np_here; ++k) {
        // The tile base goes through an opaque register move: left warp-uniform, ptxas 12.9 addresses the copies as
        // LDGSTS [R+UR+imm], which a B200 rejects at run time ("an illegal instruction was encountered").
        uint32_t xb = smem_u32(sb + TILE_BYTES) + k * XTILE;
        asm volatile("" : "+r"(xb));
#pragma unroll
        for (int seg = lane; seg < NSV * KC; seg += 32) {  // 16-byte segments: row = seg / KC
          const int s = seg / KC, part = seg % KC;
          const double2* src = (from_q ? qptr[k * NSV + s] : (const double2*)wptr[k * NSV + s]) + mem0 + part;
          const int pr = part ^ flipP;
          const uint32_t dst = dlay ? xb + pr * DROW + ((s ^ (pr & 1)) << 4) : xb + s * XROW + ((part ^ ((s & 1) * SWZ)) << 4);
          cp_async16_hint_s(dst, src, pol_X);
        }
      }
      cp_async_mbar_arrive_noinc(&full_bar[stage]);
      pc.advance();
      if (pc.I == nblk) pc = StepCursor{1, 0, 0};
    }
    return;
  }

  // ================= consumer warps =================
  double cre[MT][NT][2], cim[MT][NT][2];
#pragma unroll
  for (int a = 0; a < MT; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b) cre[a][b][0] = cre[a][b][1] = cim[a][b][0] = cim[a][b][1] = 0.0;

  // Early probe of the NEXT step's full barrier, issued in the middle of a GEMM step's DMMA stream: by the time the
  // step ends the answer is there, and the (normal) case "already loaded" costs no barrier round trip at the step boundary.
  bool next_ready = false;
  uint64_t* next_bar = nullptr;
  uint32_t next_par = 0;
  auto gemm_step = [&](const unsigned char* sb, int flip) {
    const double2* At = reinterpret_cast<const double2*>(sb);
    const unsigned char* Xt = sb + TILE_BYTES + pxo + (lane >> 2) * XROW;
    const int fx = flip ^ (((lane >> 2) & 1) * SWZ);  // memory-order flip of the plain solve + the row swizzle
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2 a[MT], b[NT];
      if (j == 2 && next_bar) next_ready = mbar_test_wait(next_bar, next_par);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = At[(j * 8 + MT * wm + mt) * 32 + lane];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        b[nt] = *reinterpret_cast<const double2*>(Xt + nt * 8 * XROW + (((4 * j + (lane & 3)) ^ fx) << 4));
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cre[mt][nt][0], cre[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cim[mt][nt][0], cim[mt][nt][1], a[mt].x, b[nt].y);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const double nai = -a[mt].y;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cre[mt][nt][0], cre[mt][nt][1], nai, b[nt].y);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cim[mt][nt][0], cim[mt][nt][1], a[mt].y, b[nt].x);
    }
  };
  // the same update with the B operand in the D layout of a lane-solve diagonal step (rows in processing order)
  auto gemm_step_d = [&](const unsigned char* sb) {
    const double2* At = reinterpret_cast<const double2*>(sb);
    const unsigned char* Dt = sb + TILE_BYTES + pxo + (lane & 3) * DROW + (((lane >> 2) ^ (lane & 1)) << 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2 a[MT], b[NT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = At[(j * 8 + MT * wm + mt) * 32 + lane];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) b[nt] = *reinterpret_cast<const double2*>(Dt + 4 * j * DROW + nt * 128);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cre[mt][nt][0], cre[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cim[mt][nt][0], cim[mt][nt][1], a[mt].x, b[nt].y);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const double nai = -a[mt].y;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cre[mt][nt][0], cre[mt][nt][1], nai, b[nt].y);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cim[mt][nt][0], cim[mt][nt][1], a[mt].y, b[nt].x);
    }
  };

  // two consecutive GEMM steps (stages sbA, sbB) as one unrolled stream of 8 k4 groups; stage A is released after B's
  // first fragment loads have been issued (program order: the compiler may hoist those loads above A's last DMMAs)
  auto gemm_pair = [&](const unsigned char* sbA, const unsigned char* sbB, int flip, uint64_t* emptyA) {
    const int fx = flip ^ (((lane >> 2) & 1) * SWZ);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const unsigned char* sbx = jj < 4 ? sbA : sbB;
      const int j = jj & 3;
      const double2* At = reinterpret_cast<const double2*>(sbx);
      const unsigned char* Xt = sbx + TILE_BYTES + pxo + (lane >> 2) * XROW;
      double2 a[MT], b[NT];
      if (jj == 6 && next_bar) next_ready = mbar_test_wait(next_bar, next_par);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = At[(j * 8 + MT * wm + mt) * 32 + lane];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        b[nt] = *reinterpret_cast<const double2*>(Xt + nt * 8 * XROW + (((4 * j + (lane & 3)) ^ fx) << 4));
      if (jj == 4 && lane == 0) mbar_arrive(emptyA);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cre[mt][nt][0], cre[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cim[mt][nt][0], cim[mt][nt][1], a[mt].x, b[nt].y);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const double nai = -a[mt].y;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cre[mt][nt][0], cre[mt][nt][1], nai, b[nt].y);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(cim[mt][nt][0], cim[mt][nt][1], a[mt].y, b[nt].x);
    }
  };

  StepCursor cons{0, 0, 0};
  uint32_t xr_par = 0;  // lane-solve: per-stage parity of the next completion of xready_bar[stage]
#ifdef PSA_TIMING
  long long tm_wait = 0, tm_gemm = 0, tm_diag = 0, tm_sync = 0, tm0;
  unsigned long long gt0, gt1;
  unsigned smid;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
#endif
  for (int t = 0; t < total_steps; ++t) {
    const int stage = t % S;
    unsigned char* sb = smem + stage * STAGE_BYTES;
    const int flip = cons.phase ? (KC - 1) : 0;
    if constexpr (LANE_SOLVE) wm = (wq + cons.I + pi) & 3;  // ownership rotates: every warp solves one chunk per block row
    {
      // A run of GEMM steps (the rest of this block row's off-diagonal tiles) in a tight loop: the ring position advances
      // by increments, nothing else is recomputed per step.  No __syncwarp before the release: mma.sync is a warp-wide
      // instruction, so once the step's last DMMA has issued every lane's fragment loads from the stage have completed.
      const int first = (cons.phase == 1) ? max(cons.c, c0) : cons.c;   // (skipped padding steps take the generic path)
      if (first == cons.c && cons.c < CPB * cons.I) {
        const int cnt = CPB * cons.I - cons.c;
        int st = stage;
        uint32_t par = (uint32_t)(t / S) & 1u;
        unsigned char* sbr = sb;
        PSA_T0();
        const bool pairing = PAIR_STEPS && P.pair != 0;
        for (int k = 0; k < cnt;) {
          if (!next_ready) mbar_wait(&full_bar[st], par);
          next_ready = false;
          const int st1 = st + 1 == S ? 0 : st + 1;
          const uint32_t par1 = st1 == 0 ? par ^ 1u : par;
          unsigned char* sb1 = st1 == 0 ? smem : sbr + STAGE_BYTES;
          // two steps as ONE instruction stream when the next stage has landed too: no barrier round trip and no drained
          // DMMA pipe between them (the second step's first fragment loads overlap the first step's last DMMAs)
          if (pairing && k + 1 < cnt && mbar_test_wait(&full_bar[st1], par1)) {
            const int st2 = st1 + 1 == S ? 0 : st1 + 1;
            next_bar = (t + k + 2 < total_steps) ? &full_bar[st2] : nullptr;
            next_par = st2 == 0 ? par1 ^ 1u : par1;
            gemm_pair(sbr, sb1, flip, &empty_bar[st]);
            if (lane == 0) mbar_arrive(&empty_bar[st1]);
            sbr = st2 == 0 ? smem : sb1 + STAGE_BYTES;
            par = next_par;
            st = st2;
            k += 2;
          } else {
            next_bar = (t + k + 1 < total_steps) ? &full_bar[st1] : nullptr;
            next_par = par1;
            gemm_step(sbr, flip);
            if (lane == 0) mbar_arrive(&empty_bar[st]);
            sbr = sb1;
            par = par1;
            st = st1;
            k += 1;
          }
        }
        next_bar = nullptr;
        PSA_T1(tm_gemm);
        t += cnt - 1;
        cons.c = CPB * cons.I - 1;
        cons.advance();
        if (cons.I == nblk) cons = StepCursor{1, 0, 0};
        continue;
      }
    }
    PSA_T0();
    if (!next_ready) mbar_wait(&full_bar[stage], (uint32_t)(t / S) & 1u);
    next_ready = false;
    PSA_T1(tm_wait);

    if (cons.phase == 1 && cons.c < c0) {
      // skipped step (pure padding of the plain solve)
    } else if (cons.c < CPB * cons.I) {
      PSA_T0();
      {
        const int t1 = t + 1, st1 = t1 % S;
        next_bar = t1 < total_steps ? &full_bar[st1] : nullptr;
        next_par = (uint32_t)(t1 / S) & 1u;
      }
      gemm_step(sb, flip);
      next_bar = nullptr;
      PSA_T1(tm_gemm);
    } else {
      PSA_T0();
      const int cp = cons.c - CPB * cons.I;  // diagonal sub-chunk 0..3
      if constexpr (LANE_SOLVE) {
        // Lane-solve diagonal step.  No CTA-wide synchronisation: the warp that owns rows 16cp..16cp+15 solves the
        // chunk on its own, one LANE per shift (the whole 16x16 forward substitution in registers, no shuffles), and
        // publishes x on the stage's xready barrier; warps that own rows below wait for it and update; warps that own
        // rows above have nothing to do and move on.  With the ownership rotating from block row to block row every
        // warp pays one solve and, on average, 1.5 updates per block row, and the four warps pass through the four
        // diagonal steps skewed by up to a ring, overlapping each other's solves with GEMM steps.
        // Same arithmetic per shift and row, in the same order, as the four-lanes-per-shift version: bit-identical.
        unsigned char* Dt = sb + TILE_BYTES + pxo;
        uint64_t* xbar = &xready_bar[stage * NP + pi];
        const uint32_t xpar = (xr_par >> stage) & 1u;
        xr_par ^= 1u << stage;
        if (wm == cp) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int r = 8 * mt + (lane >> 2), sl = 8 * nt + 2 * (lane & 3) + e;
                double2* pb = reinterpret_cast<double2*>(Dt + r * DROW + ((sl ^ (r & 1)) << 4));
                const double2 g = *pb;
                *pb = make_double2(g.x - cre[mt][nt][e], g.y - cim[mt][nt][e]);
                cre[mt][nt][e] = 0.0;
                cim[mt][nt][e] = 0.0;
              }
          __syncwarp();
          if (lane < NSV) {
            const int s = lane;
            const double2* At = reinterpret_cast<const double2*>(sb);
            double2 zt = zsh[pi * NSV + s];
            if (cons.phase == 0) zt.y = -zt.y;  // adjoint system: diagonal conj(t_ii - z)
            const int p0 = MB * cons.I + KC * cp;
            const int mem0 = cons.phase ? n_pad - KC * cons.c - KC : KC * cons.c;
            double2 b[KC];
#pragma unroll
            for (int r = 0; r < KC; ++r) b[r] = *reinterpret_cast<const double2*>(Dt + r * DROW + ((s ^ (r & 1)) << 4));
#pragma unroll
            for (int r = 0; r < KC; ++r) {
              const int m = KC * cp + r;
              double2 d = At[((r >> 2) * 8 + (m >> 3)) * 32 + (m & 7) * 4 + (r & 3)];
              d.x -= zt.x;
              d.y -= zt.y;
              const double2 dinv = cinv_fixed(d);
              double2 x = cmul_fixed(b[r], dinv);
              const int op = cons.phase ? n_pad - 1 - (p0 + r) : p0 + r;
              if (op >= n) x = make_double2(0.0, 0.0);
              b[r] = x;
#pragma unroll
              for (int i = r + 1; i < KC; ++i) {
                const int m2 = KC * cp + i;
                const double2 l = At[((r >> 2) * 8 + (m2 >> 3)) * 32 + (m2 & 7) * 4 + (r & 3)];
                cmsub(b[i], l, x);
              }
            }
            double2* op_ = wptr[pi * NSV + s] + mem0;
#pragma unroll
            for (int r = 0; r < KC; ++r) {
              *reinterpret_cast<double2*>(Dt + r * DROW + ((s ^ (r & 1)) << 4)) = b[r];
              op_[r ^ flip] = b[r];
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(xbar);  // x is in the stage's D tile (and on its way to global memory)
        } else if (wm > cp) {
          mbar_wait(xbar, xpar);
          gemm_step_d(sb);  // rows below the chunk
        }
      } else {
      unsigned char* Xt = sb + TILE_BYTES;   // holds the right-hand side rows: element (row r, shift sl) at Xt[sl][r ^ flip]
      if ((MT * wm) >> 1 == cp) {
        // this warp owns (some of the) rows 16cp..16cp+15 of the block row: turn the right-hand side into rhs - acc
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int r = 8 * ((MT * wm + mt) & 1) + (lane >> 2), sl = 8 * nt + 2 * (lane & 3) + e;
              double2* pb = reinterpret_cast<double2*>(Xt + sl * XROW + ((r ^ flip ^ (e * SWZ)) << 4));  // sl & 1 == e
              const double2 g = *pb;
              *pb = make_double2(g.x - cre[mt][nt][e], g.y - cim[mt][nt][e]);
              cre[mt][nt][e] = 0.0;
              cim[mt][nt][e] = 0.0;
            }
      }
      consumer_sync<CTHREADS>();
      if (warp < SOLVE_THREADS / 32) {  // the first four warps solve (4 lanes per shift); with CW = 8 the others wait
        constexpr int GROUPS = SOLVE_THREADS / 4;
        const int grp = tid >> 2, pl = lane & 3;
        const bool solver = grp < NSV;
        const int s = solver ? grp : 0;
        static_assert(NSV <= GROUPS, "one 4-lane group per shift");
        const int mem0 = cons.phase ? n_pad - KC * cons.c - KC : KC * cons.c;
        double2* xrow = reinterpret_cast<double2*>(Xt + s * XROW);
        const int fs = flip ^ ((s & 1) * SWZ);  // flip + row swizzle of this shift's row
        const double2* At = reinterpret_cast<const double2*>(sb);
        double2 zt = zsh[s];
        if (cons.phase == 0) zt.y = -zt.y;  // adjoint system: diagonal conj(t_ii - z)
        const int p0 = MB * cons.I + KC * cp;
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
        for (int r = 0; r < KC; ++r) {  // forward substitution in processing order; pivot r lives on lane r & 3
          const int ir = r >> 2, pr = r & 3;
          double2 x = cmul_fixed(b[ir], dinv[ir]);
          const int op = cons.phase ? n_pad - 1 - (p0 + r) : p0 + r;
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
      }
      consumer_sync<CTHREADS>();  // x is in shared memory (this step's X tile) and on its way to global memory
      if (MT * wm >= 2 * (cp + 1)) gemm_step(sb, flip);  // rows below the chunk
      }
      PSA_T1(tm_diag);
    }
    PSA_T0();
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);  // release: this warp is done with the stage (and its x is written)
    PSA_T1(tm_sync);
    cons.advance();
    if (cons.I == nblk) cons = StepCursor{1, 0, 0};
  }
#ifdef PSA_TIMING
  if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
    printf("[timing deep] cta %d warp %d: load-wait %lld gemm %lld diag %lld release %lld clk, %d steps\n", (int)blockIdx.x,
           warp, tm_wait, tm_gemm, tm_diag, tm_sync, total_steps);
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
  if (lane == 0 && warp == 0)
    printf("[cta] %d sm %u start %llu end %llu wait %lld gemm %lld diag %lld\n", (int)blockIdx.x, smid, gt0, gt1, tm_wait, tm_gemm,
           tm_diag);
#endif
}

// one CTA (sparse launches) or two CTAs per SM: the register file is not the limit
template <int NSV, int S, int CW, int OCC>
__global__ void __launch_bounds__(deep_threads(CW), OCC) solve_deep_kernel(SolveParams P) {
  solve_deep_body<NSV, S, CW, OCC == 2>(P);  // full variant (2 CTAs per SM): padded rows; sparse variant: swizzled
}

// NP panels per CTA, one CTA per SM: 4 NP consumer warps (NP DMMA-issuing warps per SM sub-partition) on ONE ring whose
// T tile serves all NP panels -- a lone DMMA-issuing warp reaches only ~85 % of a sub-partition's pipe, so the more
// warps remain when one of them is busy with a diagonal step or waits for a load, the better; and the T tile is
// fetched once per SM and step instead of once per panel.
template <int NSV, int S, int NP, bool PAD>
__global__ void __launch_bounds__(deep_threads(4 * NP), 1) solve_multi_kernel(SolveParams P) {
  solve_deep_body<NSV, S, 4, PAD, NP>(P);
}

// ---- scheduler: refill idle slots from the point queue, compact the active list -------------------------
struct SchedParams {
  int W;              // slot capacity (scratch slot has id W)
  int P;              // points owned by this device
  int rows_local;     // rows of the grid owned by this device
  int row0, row_step; // global row of local row r: row0 + r*row_step
  const double* x;    // lx
  const double* y;    // ly (global)
  const int* order;   // nullable: queue position -> local point index (longest-job-first order)
  const int* row_batch;  // nullable: global row -> start-vector index (per-batch q, compute.jl:199-208)
  int* slot_point;    // local point index or -1
  double2* slot_z;
  int* slot_iter;
  double* slot_beta;
  double* slot_sigold;
  int* slot_sel;
  int* slot_batch;
  int* active;        // out, padded with W (the scratch slot) to a multiple of APAD
  int* fresh;         // out, list of newly filled slots (the HessQR path factorises those)
  int* state;         // [0]=points started so far [1]=n_active [2]=n_fresh [3]=live ticks [4]=value of [0] at the last
                      // re-sort of the queue (queue position of the next point = [0] - [4])
  int* prog;          // nullable: per point progress, set to 1 when the point gets a slot (dynamic queue order)
  unsigned long long* host_state;  // mapped pinned mirror: [0] = (tick << 32) | n_active, [1] = (tick << 32) | n_fresh
  unsigned tick;      // sequence number of this tick (1, 2, ...)
};

// exclusive scan of one int per thread over a 1024-thread block (warp shuffles + one shared-memory hop);
// returns the exclusive prefix, leaves the block total in *total_s
__device__ __forceinline__ int block_exscan_1024(int v, int* wsum /* 32 ints */, int* total_s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protects wsum / total_s from the previous use
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    wsum[lane] = winc - w;
    if (lane == 31) *total_s = winc;
  }
  __syncthreads();
  return wsum[warp] + inc - v;
}

__global__ void __launch_bounds__(1024) sched_kernel(SchedParams S) {
  __shared__ int wsum[32];
  __shared__ int total_s;
  const int tid = threadIdx.x;
  const int C = (S.W + 1023) / 1024;
  const int lo = min(S.W, tid * C), hi = min(S.W, lo + C);
  const int started = S.state[0], head = started - S.state[4];
  int idle = 0;
  for (int i = lo; i < hi; ++i) idle += (S.slot_point[i] < 0);
  int rank = block_exscan_1024(idle, wsum, &total_s);
  const int total_idle = total_s;
  const int n_fresh = max(0, min(total_idle, S.P - started));
  int act = 0;
  for (int i = lo; i < hi; ++i) {
    if (S.slot_point[i] < 0) {
      const int qpos = head + rank;
      if (rank < n_fresh) {
        const int p = S.order ? S.order[qpos] : qpos;
        if (S.prog) S.prog[p] = 1;
        const int jr = p % S.rows_local, k = p / S.rows_local;
        const int j = S.row0 + jr * S.row_step;
        S.slot_point[i] = p;
        S.slot_z[i] = make_double2(S.x[k], S.y[j]);
        S.slot_iter[i] = 0;
        S.slot_beta[i] = 0.0;
        S.slot_sigold[i] = 0.0;
        S.slot_sel[i] = 0;
        S.slot_batch[i] = S.row_batch ? S.row_batch[j] : 0;
        S.fresh[rank] = i;
        ++act;
      }
      ++rank;
    } else {
      ++act;
    }
  }
  int arank = block_exscan_1024(act, wsum, &total_s);
  const int n_active = total_s;
  for (int i = lo; i < hi; ++i)
    if (S.slot_point[i] >= 0) S.active[arank++] = i;
  const int padded = (n_active + APAD - 1) / APAD * APAD;
  for (int i = n_active + tid; i < padded; i += 1024) S.active[i] = S.W;
  if (tid == 0) {
    S.state[0] = started + n_fresh;
    S.state[1] = n_active;
    S.state[2] = n_fresh;
    if (n_active > 0) S.state[3] += 1;  // live ticks so far
    volatile unsigned long long* hs = S.host_state;
    hs[1] = ((unsigned long long)S.tick << 32) | (unsigned)n_fresh;
    __threadfence_system();
    hs[0] = ((unsigned long long)S.tick << 32) | (unsigned)n_active;  // the word the host polls, written last
    __threadfence_system();
  }
}

// ---- largest eigenvalue of the l x l Lanczos tridiagonal, one warp, 32-way multisection Sturm counts ---
// Stands in for `maximum(eigvals(H[1:l,1:l]))` (compute.jl:640-642).  d[], e2[] are in shared memory and
// scaled to unit max-norm (e2 = squared off-diagonals), so nothing can overflow.
__device__ __forceinline__ int sturm_count(int l, const double* d, const double* e2, double x) {
  double q = d[0] - x;
  if (q == 0.0) q = -1e-300;
  int cnt = q < 0.0;
  for (int i = 1; i < l; ++i) {
    q = d[i] - x - e2[i - 1] / q;
    if (q == 0.0) q = -1e-300;
    cnt += q < 0.0;
  }
  return cnt;  // eigenvalues < x
}

__device__ double tridiag_max_eig_warp(int l, const double* d, const double* e2, int lane) {
  if (l == 1) return d[0];
  double lo = -1e300, hi = -1e300;
  for (int i = 0; i < l; ++i) {
    const double el = i > 0 ? sqrt(e2[i - 1]) : 0.0, er = i + 1 < l ? sqrt(e2[i]) : 0.0;
    lo = fmax(lo, d[i]);            // lambda_max >= max diagonal  => count(lo) < l
    hi = fmax(hi, d[i] + el + er);  // Gershgorin
  }
  hi += 4e-16 * fmax(fabs(hi), 1.0);  // count(hi) == l
  for (int round = 0; round < 16; ++round) {
    const double w = hi - lo;
    const double x = lo + w * ((double)(lane + 1) / 33.0);
    const bool ok = sturm_count(l, d, e2, x) >= l;
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    double nlo, nhi;
    if (mask == 0u) {
      nlo = __shfl_sync(0xffffffffu, x, 31);
      nhi = hi;
    } else {
      const int f = __ffs(mask) - 1;
      nhi = __shfl_sync(0xffffffffu, x, f);
      const double below = __shfl_sync(0xffffffffu, x, f > 0 ? f - 1 : 0);
      nlo = f > 0 ? below : lo;
    }
    lo = nlo;
    hi = nhi;
    if (hi - lo <= 4.5e-16 * fmax(fabs(lo), fabs(hi))) break;
  }
  return 0.5 * (lo + hi);
}

// ---- fused Lanczos recurrence + convergence test ------------------------------------------------------
// One CTA per active shift.  compute.jl:630-661:  v -= beta*qold; alpha = Re(q.v); v -= alpha q; beta = |v|;
// qold = q; q = v/beta; sigma = max eig(H_l); stop on |sigold/sigma - 1| < tol or beta == 0.
struct LanczosParams {
  int n, n_pad, W;
  const int* active;
  const int* n_active;
  double2* Q0;
  double2* Q1;
  double2* Wv;
  int* slot_point;
  int* slot_iter;
  double* slot_beta;
  double* slot_sigold;
  int* slot_sel;
  const int* slot_batch;
  const double2* q0;   // [nbatch][n_pad] start vectors
  double* hist_alpha;  // [W+1][hist]
  double* hist_beta;
  int hist;            // history stride (>= maxit + 2); also sizes the dynamic shared memory (2 * hist doubles)
  double tol;
  int maxit;
  double bigsigma;
  double* Z;          // local point order
  int* iters;         // nullable
  unsigned long long* counts;  // PSA_COUNT_*
  int* prog;          // nullable: per point iterations so far (dynamic queue order)
};

constexpr int LZ_THREADS = 256;

__global__ void __launch_bounds__(LZ_THREADS) lanczos_kernel(LanczosParams L) {
  __shared__ double red[LZ_THREADS / 32];
  extern __shared__ double lz_dyn[];
  double* sd = lz_dyn;
  double* se2 = lz_dyn + L.hist;
  if ((int)blockIdx.x >= *L.n_active) return;
  const int slot = L.active[blockIdx.x];
  const int tid = threadIdx.x;
  const int n = L.n;
  const int sel = L.slot_sel[slot];
  double2* q = (sel ? L.Q1 : L.Q0) + (size_t)slot * L.n_pad;
  double2* qold = (sel ? L.Q0 : L.Q1) + (size_t)slot * L.n_pad;
  double2* w = L.Wv + (size_t)slot * L.n_pad;
  const int l = L.slot_iter[slot] + 1;
  const double beta_old = L.slot_beta[slot];
  // first iteration of a point: q is the start vector of its row batch (`q = q0[1:n]`, compute.jl:622); it is
  // materialised into the slot here (the solve kernel read it in place), so no separate init pass is needed
  const double2* qsrc = (l == 1) ? L.q0 + (size_t)L.slot_batch[slot] * L.n_pad : q;

  // v = w - beta*qold ; alpha = Re(q . v)
  double part = 0.0;
  for (int i = tid; i < n; i += LZ_THREADS) {
    double2 v = w[i];
    const double2 qq = qsrc[i];
    if (l > 1) {
      const double2 o = qold[i];
      v.x -= beta_old * o.x;
      v.y -= beta_old * o.y;
      w[i] = v;
    } else {
      q[i] = qq;
    }
    part = fma(qq.x, v.x, part);
    part = fma(qq.y, v.y, part);
  }
  const double alpha = block_sum<LZ_THREADS>(part, red);
  // v -= alpha q ; beta = |v|
  part = 0.0;
  for (int i = tid; i < n; i += LZ_THREADS) {
    double2 v = w[i];
    const double2 qq = q[i];
    v.x -= alpha * qq.x;
    v.y -= alpha * qq.y;
    w[i] = v;
    part = fma(v.x, v.x, part);
    part = fma(v.y, v.y, part);
  }
  double ss = block_sum<LZ_THREADS>(part, red);
  double beta;
  if (isfinite(ss) && ss > 1e-280) {
    beta = sqrt(ss);
  } else {  // scaled 2-norm (the reference's norm() is BLAS dznrm2, which cannot overflow prematurely)
    double mx = 0.0;
    for (int i = tid; i < n; i += LZ_THREADS) {
      const double2 v = w[i];
      mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
    }
    mx = block_max<LZ_THREADS>(mx, red);
    if (mx == 0.0 || !isfinite(mx)) {
      beta = mx == 0.0 ? 0.0 : ss;  // 0, or inf/nan propagated
    } else {
      part = 0.0;
      for (int i = tid; i < n; i += LZ_THREADS) {
        const double2 v = w[i];
        const double a = v.x / mx, b = v.y / mx;
        part = fma(a, a, part);
        part = fma(b, b, part);
      }
      beta = mx * sqrt(block_sum<LZ_THREADS>(part, red));
    }
  }
  // qold <- q (buffer swap), q <- v / beta (written over the old qold buffer)
  for (int i = tid; i < n; i += LZ_THREADS) {
    const double2 v = w[i];
    qold[i] = make_double2(v.x / beta, v.y / beta);
  }

  double* ha = L.hist_alpha + (size_t)slot * L.hist;
  double* hb = L.hist_beta + (size_t)slot * L.hist;
  const bool finite_ab = isfinite(alpha) && isfinite(beta);
  if (tid == 0) {
    ha[l - 1] = alpha;
    hb[l - 1] = beta;
  }
  // scaled copy of H[1:l,1:l] into shared memory
  double mx = 0.0;
  if (finite_ab) {
    for (int i = tid; i < l; i += LZ_THREADS) {
      const double a = (i == l - 1) ? alpha : ha[i];
      mx = fmax(mx, fabs(a));
      if (i < l - 1) mx = fmax(mx, fabs(hb[i]));
    }
  }
  mx = block_max<LZ_THREADS>(mx, red);
  if (finite_ab && mx > 0.0) {
    for (int i = tid; i < l; i += LZ_THREADS) {
      const double a = (i == l - 1) ? alpha : ha[i];
      sd[i] = a / mx;
      if (i < l - 1) {
        const double b = hb[i] / mx;
        se2[i] = b * b;
      }
    }
  }
  __syncthreads();
  if (tid < 32) {
    double sigma;
    bool failed = false;
    if (!finite_ab) {
      sigma = L.bigsigma;
      failed = true;
    } else if (mx == 0.0) {
      sigma = 0.0;
    } else {
      sigma = mx * tridiag_max_eig_warp(l, sd, se2, tid);
    }
    if (tid == 0) {
      const double sigold = L.slot_sigold[slot];
      const bool conv = !failed && (fabs(sigold / sigma - 1.0) < L.tol || beta == 0.0);
      if (conv || failed || l >= L.maxit) {
        const int p = L.slot_point[slot];
        L.Z[p] = 1.0 / sqrt(sigma);
        if (L.iters) L.iters[p] = l;
        if (L.prog) L.prog[p] = l;
        L.slot_point[slot] = -1;
        if (!conv) atomicAdd(&L.counts[0], 1ull);  // incl. fallback points (compute.jl:654-656 leaves lancz_converged false)
        if (failed) atomicAdd(&L.counts[1], 1ull);
        atomicAdd(&L.counts[2], (unsigned long long)l);
        atomicAdd(&L.counts[3], 1ull);
      } else {
        L.slot_sigold[slot] = sigma;
        L.slot_iter[slot] = l;
        L.slot_beta[slot] = beta;
        L.slot_sel[slot] = sel ^ 1;
        if (L.prog) L.prog[L.slot_point[slot]] = l + 1;
      }
    }
  }
}

// Warp-per-slot variant for short vectors (n <= LZW_MAX_N): one warp owns a slot, all reductions are shuffle
// butterflies (deterministic), no block-level synchronisation, eight slots per CTA.  For the small-matrix configs
// (README 150 x 150, the 200 x 200 Hessenberg fast path) the CTA-per-slot kernel above spends its time in
// __syncthreads with one element per thread; this one is ~4x cheaper per tick there.  Which kernel runs depends on
// n only, never on scheduling, so results stay reproducible.
constexpr int LZW_MAX_N = 1024;
constexpr int LZW_WARPS = 8;
constexpr int LZW_MAX_HIST = 1024;  // 8 warps x 2 x 1024 doubles = 128 KiB of dynamic shared memory at most

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(LZW_WARPS * 32) lanczos_warp_kernel(LanczosParams L) {
  extern __shared__ double lz_dyn[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int a = blockIdx.x * LZW_WARPS + wid;
  if (a >= *L.n_active) return;
  double* sd = lz_dyn + (size_t)wid * 2 * L.hist;
  double* se2 = sd + L.hist;
  const int slot = L.active[a];
  const int n = L.n;
  const int sel = L.slot_sel[slot];
  double2* q = (sel ? L.Q1 : L.Q0) + (size_t)slot * L.n_pad;
  double2* qold = (sel ? L.Q0 : L.Q1) + (size_t)slot * L.n_pad;
  double2* w = L.Wv + (size_t)slot * L.n_pad;
  const int l = L.slot_iter[slot] + 1;
  const double beta_old = L.slot_beta[slot];
  const double2* qsrc = (l == 1) ? L.q0 + (size_t)L.slot_batch[slot] * L.n_pad : q;

  double part = 0.0;
  for (int i = lane; i < n; i += 32) {
    double2 v = w[i];
    const double2 qq = qsrc[i];
    if (l > 1) {
      const double2 o = qold[i];
      v.x -= beta_old * o.x;
      v.y -= beta_old * o.y;
      w[i] = v;
    } else {
      q[i] = qq;
    }
    part = fma(qq.x, v.x, part);
    part = fma(qq.y, v.y, part);
  }
  const double alpha = warp_sum(part);
  part = 0.0;
  for (int i = lane; i < n; i += 32) {
    double2 v = w[i];
    const double2 qq = qsrc[i];
    v.x -= alpha * qq.x;
    v.y -= alpha * qq.y;
    w[i] = v;
    part = fma(v.x, v.x, part);
    part = fma(v.y, v.y, part);
  }
  const double ss = warp_sum(part);
  double beta;
  if (isfinite(ss) && ss > 1e-280) {
    beta = sqrt(ss);
  } else {  // scaled 2-norm, as dznrm2
    double mx = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double2 v = w[i];
      mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
    }
    mx = warp_max(mx);
    if (mx == 0.0 || !isfinite(mx)) {
      beta = mx == 0.0 ? 0.0 : ss;
    } else {
      part = 0.0;
      for (int i = lane; i < n; i += 32) {
        const double2 v = w[i];
        const double a_ = v.x / mx, b_ = v.y / mx;
        part = fma(a_, a_, part);
        part = fma(b_, b_, part);
      }
      beta = mx * sqrt(warp_sum(part));
    }
  }
  for (int i = lane; i < n; i += 32) {
    const double2 v = w[i];
    qold[i] = make_double2(v.x / beta, v.y / beta);
  }
  double* ha = L.hist_alpha + (size_t)slot * L.hist;
  double* hb = L.hist_beta + (size_t)slot * L.hist;
  const bool finite_ab = isfinite(alpha) && isfinite(beta);
  if (lane == 0) {
    ha[l - 1] = alpha;
    hb[l - 1] = beta;
  }
  double mx = 0.0;
  if (finite_ab) {
    for (int i = lane; i < l; i += 32) {
      const double a_ = (i == l - 1) ? alpha : ha[i];
      mx = fmax(mx, fabs(a_));
      if (i < l - 1) mx = fmax(mx, fabs(hb[i]));
    }
  }
  mx = warp_max(mx);
  if (finite_ab && mx > 0.0) {
    for (int i = lane; i < l; i += 32) {
      const double a_ = (i == l - 1) ? alpha : ha[i];
      sd[i] = a_ / mx;
      if (i < l - 1) {
        const double b_ = hb[i] / mx;
        se2[i] = b_ * b_;
      }
    }
  }
  __syncwarp();
  double sigma;
  bool failed = false;
  if (!finite_ab) {
    sigma = L.bigsigma;
    failed = true;
  } else if (mx == 0.0) {
    sigma = 0.0;
  } else {
    sigma = mx * tridiag_max_eig_warp(l, sd, se2, lane);
  }
  if (lane == 0) {
    const double sigold = L.slot_sigold[slot];
    const bool conv = !failed && (fabs(sigold / sigma - 1.0) < L.tol || beta == 0.0);
    if (conv || failed || l >= L.maxit) {
      const int p = L.slot_point[slot];
      L.Z[p] = 1.0 / sqrt(sigma);
      if (L.iters) L.iters[p] = l;
      if (L.prog) L.prog[p] = l;
      L.slot_point[slot] = -1;
      if (!conv) atomicAdd(&L.counts[0], 1ull);
      if (failed) atomicAdd(&L.counts[1], 1ull);
      atomicAdd(&L.counts[2], (unsigned long long)l);
      atomicAdd(&L.counts[3], 1ull);
    } else {
      L.slot_sigold[slot] = sigma;
      L.slot_iter[slot] = l;
      L.slot_beta[slot] = beta;
      L.slot_sel[slot] = sel ^ 1;
      if (L.prog) L.prog[L.slot_point[slot]] = l + 1;
    }
  }
}

// n_active: upper bound of the device-side count
inline void launch_lanczos(const LanczosParams& lp, int n_active, cudaStream_t stream) {
  if (lp.n <= LZW_MAX_N && lp.hist <= LZW_MAX_HIST && !getenv("PSA_LANCZOS_CTA"))
    lanczos_warp_kernel<<<(n_active + LZW_WARPS - 1) / LZW_WARPS, LZW_WARPS * 32, (size_t)LZW_WARPS * 2 * lp.hist * sizeof(double),
                          stream>>>(lp);
  else
    lanczos_kernel<<<n_active, LZ_THREADS, 2 * lp.hist * sizeof(double), stream>>>(lp);
}

// ---- helpers for the apply_resolvent diagnostic hook ---------------------------------------------------
__global__ void load_slots_kernel(int ns, int n, int n_pad, const double2* __restrict__ Q, const double2* __restrict__ z,
                                  double2* __restrict__ Q0, double2* __restrict__ slot_z, int* __restrict__ slot_sel,
                                  int* __restrict__ slot_iter, int* __restrict__ active, int W, int* __restrict__ state) {
  const int s = blockIdx.x;
  if (s < ns) {
    double2* q = Q0 + (size_t)s * n_pad;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) q[i] = (i < n) ? Q[(size_t)s * n + i] : make_double2(0.0, 0.0);
    if (threadIdx.x == 0) {
      slot_z[s] = z[s];
      slot_sel[s] = 0;
      slot_iter[s] = 1;  // not a fresh slot: the solve kernel reads q from Q0
      active[s] = s;
    }
  } else if (threadIdx.x == 0) {
    active[s] = W;
  }
  if (s == 0 && threadIdx.x == 0) state[1] = ns;
}

__global__ void store_slots_kernel(int n, int n_pad, const double2* __restrict__ Wv, double2* __restrict__ V) {
  const int s = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) V[(size_t)s * n + i] = Wv[(size_t)s * n_pad + i];
}

// ---- host-side launch helpers --------------------------------------------------------------------------------
// Panel width for a tick.  The solve kernel runs 4-warp CTAs (one warp per SM sub-partition), up to 4 per SM, each
// owning a panel of 32, 16 or 8 shifts.  Wide panels reuse each T tile for more shifts; narrow ones spread the
// active shifts over every SM when few are active (grid tails, strong scaling over many GPUs).  The choice
// minimises a small cost model measured on B200 (profiles/r01_panel_sweep.txt): launch time, relative to a full
// launch, with k = 1..4 co-resident CTAs per SM; more panels than that run in rounds.
inline int pick_panel_width(int n_active, int num_sms) {
  const char* env = getenv("PSA_PANEL_WIDTH");  // experiments only
  const int forced = env ? atoi(env) : 0;
  if (forced == 32 || forced == 16 || forced == 8) return forced;
  static const int widths[3] = {8, 16, 32};
  static const double cost[3][5] = {{0.0, 0.131, 0.162, 0.223, 0.280},   // width 8:  k = 1..4 CTAs per SM
                                    {0.0, 0.207, 0.290, 0.406, 0.519},   // width 16
                                    {0.0, 0.349, 0.509, 0.772, 1.000}};  // width 32
  int best = 32;
  double best_cost = 1e300;
  for (int i = 0; i < 3; ++i) {
    const int panels = (n_active + widths[i] - 1) / widths[i];
    const int per_sm = (panels + num_sms - 1) / num_sms;  // CTAs on the busiest SM
    const double c = (per_sm / 4) * cost[i][4] + cost[i][per_sm % 4];
    if (c < best_cost - 1e-12) {
      best_cost = c;
      best = widths[i];
    }
  }
  return best;
}

// Instantiations of the deep-ring kernel.  "Full" launches (more panels than SMs): 2 CTAs per SM, 4 consumer warps
// each, the deepest ring that fits twice into shared memory.  (Measured and not adopted, profiles/r02_flavor_check_v4.log:
// 3 CTAs per SM with 3-stage rings and 136 registers -- 30.7 TFLOP/s against 33.6: ring depth matters more than a third
// DMMA-issuing warp per sub-partition; 8 consumer warps at 2 CTAs per SM -- 96 registers, spills, 32.4 TFLOP/s.)
// "Sparse" launches (at most one panel per SM: grid tails, strong scaling over many GPUs, the pseudomode): one CTA per
// SM with 8 consumer warps (four DMMA-issuing warps per sub-partition) and 6..8 stages.
constexpr int DEEP_OCC = 2;
constexpr int DEEP_S32 = 4, DEEP_S16 = 5, DEEP_S8 = 6;        // 2 CTAs per SM, CW = 4
constexpr int SPARSE_S32 = 6, SPARSE_S16 = 8, SPARSE_S8 = 8;  // one CTA per SM, CW = 8
constexpr int MULTI2_S = 6;                                   // one CTA per SM, 2 panels of 32 per CTA (experimental)

inline int solve_set_attrs() {
  cudaError_t e = cudaSuccess;
  auto set = [&](auto* k, int bytes) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  };
  set(solve_kernel<32>, solve_smem(32));
  set(solve_kernel<16>, solve_smem(16));
  set(solve_kernel<8>, solve_smem(8));
  set(solve_deep_kernel<32, DEEP_S32, 4, DEEP_OCC>, deep_smem<32, DEEP_S32, true>());
  set(solve_deep_kernel<16, DEEP_S16, 4, DEEP_OCC>, deep_smem<16, DEEP_S16, true>());
  set(solve_deep_kernel<8, DEEP_S8, 4, DEEP_OCC>, deep_smem<8, DEEP_S8, true>());
  set(solve_deep_kernel<32, SPARSE_S32, 8, 1>, deep_smem<32, SPARSE_S32>());
  set(solve_deep_kernel<16, SPARSE_S16, 8, 1>, deep_smem<16, SPARSE_S16>());
  set(solve_deep_kernel<8, SPARSE_S8, 8, 1>, deep_smem<8, SPARSE_S8>());
  set(solve_multi_kernel<32, MULTI2_S, 2, false>, deep_smem<32, MULTI2_S, false, 2>());
  return e == cudaSuccess ? 0 : -1;
}

inline int solve_flavor() {  // experiments: PSA_SOLVE_FLAVOR=shallow|deep|auto (read at every launch)
  const char* e = getenv("PSA_SOLVE_FLAVOR");
  if (!e) return 2;
  return !strcmp(e, "shallow") ? 0 : !strcmp(e, "deep") ? 1 : 2;
}

// Panel width for the deep-ring kernel.  A launch costs (GEMM steps) x tg + (diagonal steps) x td per panel round, with
// per-step times (clock cycles per CTA, wall clock) measured on B200 at n = 4000 (profiles/r02_flavor_check_v3.log, r02_flavor_check_v4.log,
// r02_phase_timing_deep_v2.log) for k = 1 (at most one CTA per SM: run with 8 consumer warps, four DMMA-issuing warps per
// SM sub-partition) and k = 2 (two CTAs per SM, 4 consumer warps each: register budget).  The diagonal steps cost
// about the same for every width, so small matrices (few block rows: diagonal-dominated) prefer wide panels and
// large ones prefer whatever fills the SMs evenly.
inline int pick_deep_width(int n_active, int num_sms, int nblk) {
  const char* env = getenv("PSA_PANEL_WIDTH");  // experiments only
  const int forced = env ? atoi(env) : 0;
  if (forced == 32 || forced == 16 || forced == 8) return forced;
  static const int widths[3] = {8, 16, 32};
  static const double tg[3][3] = {{0, 679, 1195}, {0, 1190, 2236}, {0, 2235, 4300}};
  static const double td[3][3] = {{0, 3000, 5000}, {0, 3800, 6000}, {0, 4800, 7000}};
  const double G = 2.0 * nblk * (nblk - 1), D = 4.0 * nblk;
  int best = 32;
  double best_cost = 1e300;
  for (int i = 0; i < 3; ++i) {
    const int panels = (n_active + widths[i] - 1) / widths[i];
    const int rounds = panels / (2 * num_sms), rest = panels % (2 * num_sms);
    const double t1 = G * tg[i][1] + D * td[i][1], t2 = G * tg[i][2] + D * td[i][2];
    const double c = rounds * t2 + (rest == 0 ? 0.0 : (rest <= num_sms ? t1 : t2));
    if (c < best_cost * (1.0 - 1e-9)) {
      best_cost = c;
      best = widths[i];
    }
  }
  return best;
}

inline void launch_deep(const SolveParams& sp, int width, int panels, bool sparse, cudaStream_t stream) {
  if (const char* em = getenv("PSA_MULTI")) {  // experiments only: two panels of 32 per CTA, one CTA per SM, 6-stage ring
    if (width == 32 && atoi(em) == 2) {
      solve_multi_kernel<32, MULTI2_S, 2, false><<<(panels + 1) / 2, deep_threads(8), deep_smem<32, MULTI2_S, false, 2>(), stream>>>(sp);
      return;
    }
  }
  if (sparse) {
    switch (width) {
      case 32: solve_deep_kernel<32, SPARSE_S32, 8, 1><<<panels, deep_threads(8), deep_smem<32, SPARSE_S32>(), stream>>>(sp); break;
      case 16: solve_deep_kernel<16, SPARSE_S16, 8, 1><<<panels, deep_threads(8), deep_smem<16, SPARSE_S16>(), stream>>>(sp); break;
      default: solve_deep_kernel<8, SPARSE_S8, 8, 1><<<panels, deep_threads(8), deep_smem<8, SPARSE_S8>(), stream>>>(sp); break;
    }
    return;
  }
  switch (width) {
    case 32: solve_deep_kernel<32, DEEP_S32, 4, DEEP_OCC><<<panels, deep_threads(4), deep_smem<32, DEEP_S32, true>(), stream>>>(sp); break;
    case 16: solve_deep_kernel<16, DEEP_S16, 4, DEEP_OCC><<<panels, deep_threads(4), deep_smem<16, DEEP_S16, true>(), stream>>>(sp); break;
    default: solve_deep_kernel<8, DEEP_S8, 4, DEEP_OCC><<<panels, deep_threads(4), deep_smem<8, DEEP_S8, true>(), stream>>>(sp); break;
  }
}

// n_active is an upper bound of the device-side count (panels past it exit at once).  The deep-ring kernel is the
// default; PSA_SOLVE_FLAVOR=shallow selects the round-1 choreography (same results bit for bit) for A/B timing.
inline void launch_solve(const SolveParams& sp, int n_active, int num_sms, cudaStream_t stream) {
  if (solve_flavor() != 0) {
    const int width = pick_deep_width(n_active, num_sms, sp.nblk);
    const int panels = (n_active + width - 1) / width;
    const char* ecw = getenv("PSA_DEEP_CW");  // experiments only: 8 = sparse instantiation, 4 = full one
    const bool sparse = ecw ? atoi(ecw) == 8 : panels <= num_sms;
    launch_deep(sp, width, panels, sparse, stream);
    return;
  }
  const int width = pick_panel_width(n_active, num_sms);
  const int panels = (n_active + width - 1) / width;
  switch (width) {
    case 32: solve_kernel<32><<<panels, SOLVE_THREADS, solve_smem(32), stream>>>(sp); break;
    case 16: solve_kernel<16><<<panels, SOLVE_THREADS, solve_smem(16), stream>>>(sp); break;
    default: solve_kernel<8><<<panels, SOLVE_THREADS, solve_smem(8), stream>>>(sp); break;
  }
}

}  // namespace psa
