// libpsagpu.so -- host side of the C ABI declared in include/psa_gpu.h.
// One host thread drives every device of the context through per-device streams; all launches are
// asynchronous and the only host<->device synchronisation is one tiny mapped-memory read per scheduler tick.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/psa_gpu.h"
#include "psa_hess.cuh"
#include "psa_nccl.h"
#include "psa_square.cuh"
#include "psa_svd.cuh"

namespace {

using namespace psa;

#define PSA_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      char buf_[512];                                                                          \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      ctx->err = buf_;                                                                         \
      return e_ == cudaErrorMemoryAllocation ? PSA_ERR_NOMEM : PSA_ERR_CUDA;                   \
    }                                                                                          \
  } while (0)

#define PSA_NCCL(call)                                                                               \
  do {                                                                                               \
    ncclResult_t r_ = (call);                                                                        \
    if (r_ != ncclSuccess) {                                                                         \
      ctx->err = std::string(#call) + " failed: " + ctx->nccl.GetErrorString(r_);                    \
      return PSA_ERR_NCCL;                                                                           \
    }                                                                                                \
  } while (0)

struct EventPair {
  cudaEvent_t a, b;
};

struct Device {
  int dev = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  // matrix
  double2* T = nullptr;  // raw copy (m x n, ld = m)
  double2* packA = nullptr;
  double2* packB = nullptr;
  size_t T_cap = 0, pack_cap = 0, diag_cap = 0;
  double2* diagT = nullptr;  // eigenvalues (diagonal of T), for the scheduling-order predictor
  // scheduling order
  float *key_in = nullptr, *key_out = nullptr;
  int *ord_in = nullptr, *ord_out = nullptr;
  void* sort_tmp = nullptr;
  size_t ord_cap = 0, sort_tmp_bytes = 0;
  // slots
  int W = 0;
  size_t vec_elems = 0;
  double2 *Q0 = nullptr, *Q1 = nullptr, *Wv = nullptr;
  int *slot_point = nullptr, *slot_iter = nullptr, *slot_sel = nullptr;
  double2* slot_z = nullptr;
  double *slot_beta = nullptr, *slot_sigold = nullptr, *hist_alpha = nullptr, *hist_beta = nullptr;
  int *active = nullptr, *fresh = nullptr, *state = nullptr;
  int* host_state = nullptr;      // pinned + mapped
  int* host_state_dev = nullptr;  // device alias
  unsigned long long* counts = nullptr;
  // HessQR workspace
  double2* R = nullptr;  // per-slot triangular factors
  size_t R_elems = 0;
  // grid buffers (owned unless device API)
  double *x = nullptr, *y = nullptr, *Z = nullptr;
  double2* q0 = nullptr;
  int* iters = nullptr;
  size_t x_cap = 0, y_cap = 0, q0_cap = 0, Z_cap = 0, it_cap = 0;
  // events
  std::vector<EventPair> ev;
  cudaEvent_t g0 = nullptr, g1 = nullptr;
};

struct Context {
  std::vector<Device> devs;
  std::string err;
  NcclApi nccl;                    // multi-device contexts only
  std::vector<ncclComm_t> comms;   // one communicator per device (ncclCommInitAll)
  std::vector<double*> gather_buf; // on device 0: receive buffers for the other devices' tiles
  std::vector<int*> gather_ibuf;
  std::vector<size_t> gather_cap;
  int m = 0, n = 0, n_pad = 0, nblk = 0, algo = PSA_ALGO_NONE;
  double stats[PSA_NSTATS] = {0};
  bool solve_attr_set = false;
};

template <typename T>
int dev_alloc(Context* ctx, T** p, size_t count);

// Communicators over the context's devices (created lazily: single-device contexts never touch NCCL).
int ensure_comms(Context* ctx) {
  if (!ctx->comms.empty()) return PSA_OK;
  if (!ctx->nccl.load(ctx->err)) return PSA_ERR_NCCL;
  const int ng = (int)ctx->devs.size();
  std::vector<int> ids(ng);
  for (int i = 0; i < ng; ++i) ids[i] = ctx->devs[i].dev;
  ctx->comms.assign(ng, nullptr);
  ncclResult_t r = ctx->nccl.CommInitAll(ctx->comms.data(), ng, ids.data());
  if (r != ncclSuccess) {
    ctx->err = std::string("ncclCommInitAll failed: ") + ctx->nccl.GetErrorString(r);
    ctx->comms.clear();
    return PSA_ERR_NCCL;
  }
  ctx->gather_buf.assign(ng, nullptr);
  ctx->gather_ibuf.assign(ng, nullptr);
  ctx->gather_cap.assign(ng, 0);
  return PSA_OK;
}

template <typename T>
int dev_alloc(Context* ctx, T** p, size_t count) {
  if (*p) {
    cudaFree(*p);
    *p = nullptr;
  }
  PSA_CUDA(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
  return PSA_OK;
}

void free_device(Device& d) {
  cudaSetDevice(d.dev);
  void* ptrs[] = {d.diagT, d.key_in, d.key_out, d.ord_in, d.ord_out, d.sort_tmp, d.T, d.packA, d.packB, d.Q0, d.Q1, d.Wv, d.slot_point, d.slot_iter, d.slot_sel, d.slot_z,
                  d.slot_beta, d.slot_sigold, d.hist_alpha, d.hist_beta, d.active, d.fresh, d.state, d.counts,
                  d.R, d.x, d.y, d.Z, d.q0, d.iters};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (d.host_state) cudaFreeHost(d.host_state);
  for (auto& e : d.ev) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  if (d.g0) cudaEventDestroy(d.g0);
  if (d.g1) cudaEventDestroy(d.g1);
  if (d.stream) cudaStreamDestroy(d.stream);
  d = Device();
}

int ensure_slots(Context* ctx, Device& d, int want_W) {
  const size_t vec_elems = (size_t)ctx->n_pad;
  const bool hess = ctx->algo == PSA_ALGO_HESSQR || ctx->algo == PSA_ALGO_RECT_QR;
  const size_t r_elems = hess ? hess_R_elems(ctx->n) : 0;
  if (d.W >= want_W && d.vec_elems == vec_elems && d.R_elems == r_elems) return PSA_OK;
  PSA_CUDA(cudaSetDevice(d.dev));
  const int W = want_W;
  int rc;
  if ((rc = dev_alloc(ctx, &d.Q0, (size_t)(W + 1) * vec_elems))) return rc;
  if ((rc = dev_alloc(ctx, &d.Q1, (size_t)(W + 1) * vec_elems))) return rc;
  if ((rc = dev_alloc(ctx, &d.Wv, (size_t)(W + 1) * vec_elems))) return rc;
  if ((rc = dev_alloc(ctx, &d.slot_point, W + 1))) return rc;
  if ((rc = dev_alloc(ctx, &d.slot_iter, W + 1))) return rc;
  if ((rc = dev_alloc(ctx, &d.slot_sel, W + 1))) return rc;
  if ((rc = dev_alloc(ctx, &d.slot_z, W + 1))) return rc;
  if ((rc = dev_alloc(ctx, &d.slot_beta, W + 1))) return rc;
  if ((rc = dev_alloc(ctx, &d.slot_sigold, W + 1))) return rc;
  if ((rc = dev_alloc(ctx, &d.hist_alpha, (size_t)(W + 1) * HIST))) return rc;
  if ((rc = dev_alloc(ctx, &d.hist_beta, (size_t)(W + 1) * HIST))) return rc;
  if ((rc = dev_alloc(ctx, &d.active, W + 2 * APAD))) return rc;
  if ((rc = dev_alloc(ctx, &d.fresh, W + 1))) return rc;
  if (hess && (rc = dev_alloc(ctx, &d.R, (size_t)(W + 1) * r_elems))) return rc;
  // scratch slot W: benign, finite contents
  PSA_CUDA(cudaMemsetAsync(d.Q0 + (size_t)W * vec_elems, 0, vec_elems * sizeof(double2), d.stream));
  PSA_CUDA(cudaMemsetAsync(d.Q1 + (size_t)W * vec_elems, 0, vec_elems * sizeof(double2), d.stream));
  PSA_CUDA(cudaMemsetAsync(d.Wv + (size_t)W * vec_elems, 0, vec_elems * sizeof(double2), d.stream));
  PSA_CUDA(cudaMemsetAsync(d.slot_sel, 0, (W + 1) * sizeof(int), d.stream));
  PSA_CUDA(cudaMemsetAsync(d.slot_z, 0, (W + 1) * sizeof(double2), d.stream));
  if (hess) PSA_CUDA(cudaMemsetAsync(d.R + (size_t)W * r_elems, 0, r_elems * sizeof(double2), d.stream));
  d.W = W;
  d.vec_elems = vec_elems;
  d.R_elems = r_elems;
  return PSA_OK;
}

int slot_capacity(const Context* ctx, const Device& d, long long points) {
  long long cap;
  if (ctx->algo == PSA_ALGO_HESSQR || ctx->algo == PSA_ALGO_RECT_QR) {
    cap = hess_slot_capacity(ctx->n, d.num_sms);
  } else {
    cap = 4LL * d.num_sms * NS;  // four resident solve CTAs per SM, NS shifts each
  }
  long long w = std::min(cap, (points + NS - 1) / NS * NS);
  return (int)std::max<long long>(w, NS);
}

EventPair& event_pair(Device& d, size_t i) {
  while (d.ev.size() <= i) {
    EventPair e;
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    d.ev.push_back(e);
  }
  return d.ev[i];
}

int set_solve_attr(Context* ctx) {
  if (ctx->solve_attr_set) return PSA_OK;
  for (auto& d : ctx->devs) {
    PSA_CUDA(cudaSetDevice(d.dev));
    PSA_CUDA(cudaFuncSetAttribute(solve_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, solve_smem(32)));
    PSA_CUDA(cudaFuncSetAttribute(solve_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, solve_smem(16)));
    PSA_CUDA(cudaFuncSetAttribute(solve_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, solve_smem(8)));
    PSA_CUDA(cudaFuncSetAttribute(solve_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    PSA_CUDA(cudaFuncSetAttribute(solve_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    PSA_CUDA(cudaFuncSetAttribute(solve_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int rc = hess_set_attrs();
    if (rc != 0) {
      ctx->err = "cudaFuncSetAttribute failed for the HessQR kernels";
      return PSA_ERR_CUDA;
    }
  }
  ctx->solve_attr_set = true;
  return PSA_OK;
}

SolveParams solve_params(const Context* ctx, const Device& d) {
  SolveParams sp;
  sp.packA = d.packA;
  sp.packB = d.packB;
  sp.n = ctx->n;
  sp.n_pad = ctx->n_pad;
  sp.nblk = ctx->nblk;
  sp.active = d.active;
  sp.n_active = d.state + 1;
  sp.Q0 = d.Q0;
  sp.Q1 = d.Q1;
  sp.Wv = d.Wv;
  sp.sel = d.slot_sel;
  sp.z = d.slot_z;
  return sp;
}

// Panel width for a tick.  The solve kernel runs 4-warp CTAs (one warp per SM sub-partition), up to 4 per SM, each
// owning a panel of 32, 16 or 8 shifts.  Wide panels reuse each T tile for more shifts; narrow ones spread the
// active shifts over every SM when few are active (grid tails, strong scaling over many GPUs).  The choice
// minimises a small cost model measured on B200 (profiles/r01_panel_sweep.txt): launch time, relative to a full
// launch, with k = 1..4 co-resident CTAs per SM; more panels than that run in rounds.
int pick_panel_width(int n_active, int num_sms) {
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

void launch_solve(const Context* ctx, const Device& d, int n_active, int width) {
  const SolveParams sp = solve_params(ctx, d);
  const int panels = (n_active + width - 1) / width;
  switch (width) {
    case 32: solve_kernel<32><<<panels, SOLVE_THREADS, solve_smem(32), d.stream>>>(sp); break;
    case 16: solve_kernel<16><<<panels, SOLVE_THREADS, solve_smem(16), d.stream>>>(sp); break;
    default: solve_kernel<8><<<panels, SOLVE_THREADS, solve_smem(8), d.stream>>>(sp); break;
  }
}

// pack / prepare the device-resident matrix on one device (d.T already holds the raw m x n copy)
int prepare_matrix(Context* ctx, Device& d) {
  PSA_CUDA(cudaSetDevice(d.dev));
  if (ctx->algo == PSA_ALGO_SQ_LANC) {
    const long long ntiles = tiles_per_phase(ctx->nblk);
    int rc;
    if (d.pack_cap < (size_t)ntiles * TILE_ELEMS) {  // buffers are reused across set_matrix calls of the same size
      if ((rc = dev_alloc(ctx, &d.packA, (size_t)ntiles * TILE_ELEMS))) return rc;
      if ((rc = dev_alloc(ctx, &d.packB, (size_t)ntiles * TILE_ELEMS))) return rc;
      d.pack_cap = (size_t)ntiles * TILE_ELEMS;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, d.stream);
    if (d.diag_cap < (size_t)ctx->n) {
      if ((rc = dev_alloc(ctx, &d.diagT, (size_t)ctx->n))) return rc;
      d.diag_cap = (size_t)ctx->n;
    }
    diag_kernel<<<(ctx->n + 255) / 256, 256, 0, d.stream>>>(d.T, ctx->m, ctx->n, d.diagT);
    const int grid = (int)std::min<long long>(2 * ntiles, 148LL * 32);
    pack_kernel<<<grid, 256, 0, d.stream>>>(d.T, ctx->m, ctx->n, ctx->n_pad, ctx->nblk, d.packA, d.packB);
    cudaEventRecord(e1, d.stream);
    PSA_CUDA(cudaGetLastError());
    PSA_CUDA(cudaStreamSynchronize(d.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (&d == &ctx->devs[0]) ctx->stats[PSA_STAT_PACK_MS] = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  return PSA_OK;  // slot buffers are re-validated by ensure_slots (they depend on n_pad / algo)
}

int classify(int m, int n, bool is_hessenberg) {
  if (n < 55) return (long long)m * n <= SVD_MAX_ELEMS ? PSA_ALGO_SVD : PSA_ALGO_NONE;  // compute.jl:478
  if (m == n + 1 && m <= 200 && is_hessenberg) return PSA_ALGO_RECT_QR;                 // compute.jl:588-590
  if (m == n && n >= 55) return PSA_ALGO_SQ_LANC;        // compute.jl:478,534,565,594
  if (m == n + 1 && m > 200 && n <= HQ_MAX_N) return PSA_ALGO_HESSQR;  // compute.jl:579 (bw == 2 is the caller's promise)
  return PSA_ALGO_NONE;
}

struct GridArgs {
  int lx, ly;
  double tol;
  int maxit;
  double bigsigma;
  volatile int* cancel;
};

// Run the tick loop on every device.  Per device: d.x, d.y, d.q0 are resident, d.Z / d.iters are the
// (local point order) outputs.  Rows are dealt cyclically: device i owns rows i, i+ng, ...
int run_grid(Context* ctx, const GridArgs& g, std::vector<double*>& Zout, std::vector<int*>& itout, int64_t* counts) {
  const int ng = (int)ctx->devs.size();
  int rc = set_solve_attr(ctx);
  if (rc) return rc;
  std::vector<int> rows(ng), P(ng), done(ng, 0);
  std::vector<const int*> order(ng, nullptr);
  std::vector<size_t> nev(ng, 0);
  double launches = 0, solve_launches = 0, ticks = 0;
  for (int i = 0; i < ng; ++i) {
    Device& d = ctx->devs[i];
    rows[i] = (g.ly - i + ng - 1) / ng;
    P[i] = rows[i] * g.lx;
    PSA_CUDA(cudaSetDevice(d.dev));
    if (ctx->algo != PSA_ALGO_SVD) {  // the SVD branch has no Lanczos state
      if ((rc = ensure_slots(ctx, d, slot_capacity(ctx, d, P[i])))) return rc;
      PSA_CUDA(cudaMemsetAsync(d.slot_point, 0xff, (d.W + 1) * sizeof(int), d.stream));
    }
    PSA_CUDA(cudaMemsetAsync(d.state, 0, 4 * sizeof(int), d.stream));
    PSA_CUDA(cudaMemsetAsync(d.counts, 0, PSA_NCOUNTS * sizeof(unsigned long long), d.stream));
    PSA_CUDA(cudaEventRecord(d.g0, d.stream));
    if (P[i] == 0) done[i] = 1;
    order[i] = nullptr;
    if (ctx->algo == PSA_ALGO_SQ_LANC && P[i] > d.W && !getenv("PSA_NATURAL_ORDER")) {
      // more points than slots: queue them longest-job-first (see predict_kernel)
      if (d.ord_cap < (size_t)P[i]) {
        if ((rc = dev_alloc(ctx, &d.key_in, (size_t)P[i]))) return rc;
        if ((rc = dev_alloc(ctx, &d.key_out, (size_t)P[i]))) return rc;
        if ((rc = dev_alloc(ctx, &d.ord_in, (size_t)P[i]))) return rc;
        if ((rc = dev_alloc(ctx, &d.ord_out, (size_t)P[i]))) return rc;
        d.ord_cap = (size_t)P[i];
      }
      predict_kernel<<<(P[i] + 255) / 256, 256, 0, d.stream>>>(P[i], rows[i], i, ng, d.x, d.y, d.diagT, ctx->n, d.key_in,
                                                               d.ord_in);
      size_t need = 0;
      cub::DeviceRadixSort::SortPairsDescending(nullptr, need, d.key_in, d.key_out, d.ord_in, d.ord_out, P[i], 0, 32,
                                                d.stream);
      if (need > d.sort_tmp_bytes) {
        if (d.sort_tmp) cudaFree(d.sort_tmp);
        d.sort_tmp = nullptr;
        PSA_CUDA(cudaMalloc(&d.sort_tmp, need));
        d.sort_tmp_bytes = need;
      }
      PSA_CUDA(cub::DeviceRadixSort::SortPairsDescending(d.sort_tmp, need, d.key_in, d.key_out, d.ord_in, d.ord_out, P[i],
                                                         0, 32, d.stream));
      order[i] = d.ord_out;
      launches += 2;
    }
  }
  if (ctx->algo == PSA_ALGO_SVD) {
    // no Lanczos, no slots: one launch per device over all of its points
    for (int i = 0; i < ng; ++i) {
      Device& d = ctx->devs[i];
      PSA_CUDA(cudaSetDevice(d.dev));
      PSA_CUDA(cudaFuncSetAttribute(svd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      if (P[i] == 0) continue;
      SvdParams sp;
      sp.m = ctx->m;
      sp.n = ctx->n;
      sp.T = d.T;
      sp.P = P[i];
      sp.rows_local = rows[i];
      sp.row0 = i;
      sp.row_step = ng;
      sp.x = d.x;
      sp.y = d.y;
      sp.Z = Zout[i];
      sp.iters = itout[i];
      sp.counts = d.counts;
      EventPair& ep = event_pair(d, nev[i]++);
      PSA_CUDA(cudaEventRecord(ep.a, d.stream));
      svd_kernel<<<std::min(P[i], 8 * d.num_sms), SVD_THREADS, svd_smem(ctx->m, ctx->n), d.stream>>>(sp);
      PSA_CUDA(cudaEventRecord(ep.b, d.stream));
      PSA_CUDA(cudaGetLastError());
      launches += 1;
      solve_launches += (i == 0);
      done[i] = 1;
    }
  }
  const int hist_maxit = std::min(g.maxit, HIST - 2);
  int remaining = 0;
  for (int i = 0; i < ng; ++i) remaining += !done[i];
  bool cancelled = false;
  while (remaining > 0) {
    if (g.cancel && *g.cancel == 1) {
      cancelled = true;
      break;
    }
    // 1. schedule
    for (int i = 0; i < ng; ++i) {
      if (done[i]) continue;
      Device& d = ctx->devs[i];
      PSA_CUDA(cudaSetDevice(d.dev));
      SchedParams sp;
      sp.W = d.W;
      sp.P = P[i];
      sp.rows_local = rows[i];
      sp.row0 = i;
      sp.row_step = ng;
      sp.x = d.x;
      sp.y = d.y;
      sp.order = order[i];
      sp.slot_point = d.slot_point;
      sp.slot_z = d.slot_z;
      sp.slot_iter = d.slot_iter;
      sp.slot_beta = d.slot_beta;
      sp.slot_sigold = d.slot_sigold;
      sp.slot_sel = d.slot_sel;
      sp.active = d.active;
      sp.fresh = d.fresh;
      sp.state = d.state;
      sp.host_state = d.host_state_dev;
      sched_kernel<<<1, 1024, 0, d.stream>>>(sp);
      launches += 1;
    }
    // 2. read back the tick's counts and launch the tick
    for (int i = 0; i < ng; ++i) {
      if (done[i]) continue;
      Device& d = ctx->devs[i];
      PSA_CUDA(cudaSetDevice(d.dev));
      PSA_CUDA(cudaStreamSynchronize(d.stream));
      const int n_active = d.host_state[1], n_fresh = d.host_state[2];
      if (n_active == 0) {
        done[i] = 1;
        --remaining;
        continue;
      }
      ticks += (i == 0);
      EventPair& ep = event_pair(d, nev[i]++);
      if (ctx->algo == PSA_ALGO_SQ_LANC) {
        if (n_fresh > 0) {
          init_fresh_kernel<<<n_fresh, 256, 0, d.stream>>>(d.fresh, ctx->n_pad, ctx->n, d.q0, d.Q0);
          launches += 1;
        }
        PSA_CUDA(cudaEventRecord(ep.a, d.stream));
        launch_solve(ctx, d, n_active, pick_panel_width(n_active, d.num_sms));
        PSA_CUDA(cudaEventRecord(ep.b, d.stream));
        launches += 1;
        solve_launches += (i == 0);
      } else {
        HessTick ht;
        ht.m = ctx->m;
        ht.n = ctx->n;
        ht.n_pad = ctx->n_pad;
        ht.H = d.T;
        ht.R = d.R;
        ht.q0 = d.q0;
        ht.fresh = d.fresh;
        ht.n_fresh = n_fresh;
        ht.active = d.active;
        ht.n_active = n_active;
        ht.n_active_dev = d.state + 1;
        ht.Q0 = d.Q0;
        ht.Q1 = d.Q1;
        ht.Wv = d.Wv;
        ht.sel = d.slot_sel;
        ht.z = d.slot_z;
        ht.full_sweep = ctx->algo == PSA_ALGO_RECT_QR;
        PSA_CUDA(cudaEventRecord(ep.a, d.stream));
        const int nl = hess_launch_tick(ht, d.stream, ep.a, ep.b);
        launches += nl;
        solve_launches += (i == 0);
      }
      LanczosParams lp;
      lp.n = ctx->n;
      lp.n_pad = ctx->n_pad;
      lp.W = d.W;
      lp.active = d.active;
      lp.n_active = d.state + 1;
      lp.Q0 = d.Q0;
      lp.Q1 = d.Q1;
      lp.Wv = d.Wv;
      lp.slot_point = d.slot_point;
      lp.slot_iter = d.slot_iter;
      lp.slot_beta = d.slot_beta;
      lp.slot_sigold = d.slot_sigold;
      lp.slot_sel = d.slot_sel;
      lp.hist_alpha = d.hist_alpha;
      lp.hist_beta = d.hist_beta;
      lp.tol = g.tol;
      lp.maxit = hist_maxit;
      lp.bigsigma = g.bigsigma;
      lp.Z = Zout[i];
      lp.iters = itout[i];
      lp.counts = d.counts;
      lanczos_kernel<<<n_active, LZ_THREADS, 0, d.stream>>>(lp);
      launches += 1;
      PSA_CUDA(cudaGetLastError());
    }
  }
  // drain, collect counts + timing
  unsigned long long tot[PSA_NCOUNTS] = {0, 0, 0, 0};
  for (int i = 0; i < ng; ++i) {
    Device& d = ctx->devs[i];
    PSA_CUDA(cudaSetDevice(d.dev));
    PSA_CUDA(cudaEventRecord(d.g1, d.stream));
    PSA_CUDA(cudaStreamSynchronize(d.stream));
    unsigned long long c[PSA_NCOUNTS];
    PSA_CUDA(cudaMemcpy(c, d.counts, sizeof c, cudaMemcpyDeviceToHost));
    for (int k = 0; k < PSA_NCOUNTS; ++k) tot[k] += c[k];
    if (i == 0) {
      double solve_ms = 0;
      for (size_t e = 0; e < nev[i]; ++e) {
        float ms = 0;
        cudaEventElapsedTime(&ms, d.ev[e].a, d.ev[e].b);
        solve_ms += ms;
      }
      float gms = 0;
      cudaEventElapsedTime(&gms, d.g0, d.g1);
      ctx->stats[PSA_STAT_SOLVE_MS] = solve_ms;
      ctx->stats[PSA_STAT_GRID_MS] = gms;
    }
  }
  ctx->stats[PSA_STAT_SOLVE_LAUNCHES] = solve_launches;
  ctx->stats[PSA_STAT_TOTAL_LAUNCHES] = launches;
  ctx->stats[PSA_STAT_TICKS] = ticks;
  const double nn = (double)ctx->n;
  if (ctx->algo == PSA_ALGO_SQ_LANC) {
    ctx->stats[PSA_STAT_FLOPS] = 8.0 * nn * nn * (double)tot[PSA_COUNT_TOTAL_ITERS];
    ctx->stats[PSA_STAT_BYTES] = 0.0;
  } else {
    ctx->stats[PSA_STAT_FLOPS] = 10.0 * nn * nn * (double)tot[PSA_COUNT_POINTS] + 8.0 * nn * nn * (double)tot[PSA_COUNT_TOTAL_ITERS];
    ctx->stats[PSA_STAT_BYTES] = 16.0 * ctx->m * nn * (double)tot[PSA_COUNT_POINTS] + 16.0 * nn * nn * (double)tot[PSA_COUNT_TOTAL_ITERS];
  }
  if (counts)
    for (int k = 0; k < PSA_NCOUNTS; ++k) counts[k] = (int64_t)tot[k];
  return cancelled ? PSA_ERR_CANCELLED : PSA_OK;
}

template <typename T>
int ensure_cap(Context* ctx, T** p, size_t* cap, size_t want) {
  if (*cap >= want && *p) return PSA_OK;
  int rc = dev_alloc(ctx, p, want);
  if (rc) return rc;
  *cap = want;
  return PSA_OK;
}

}  // namespace

extern "C" {

const char* psa_gpu_version(void) { return "psagpu 0.1.0 sm_100a"; }

const char* psa_gpu_strerror(int status) {
  switch (status) {
    case PSA_OK: return "ok";
    case PSA_ERR_ARG: return "invalid argument";
    case PSA_ERR_CUDA: return "CUDA error";
    case PSA_ERR_CANCELLED: return "computation cancelled";
    case PSA_ERR_UNSUPPORTED: return "matrix shape not covered by the GPU path (use the host algorithm)";
    case PSA_ERR_NODEVICE: return "no usable sm_100 CUDA device";
    case PSA_ERR_NOMATRIX: return "no matrix set";
    case PSA_ERR_NCCL: return "NCCL error";
    case PSA_ERR_NOMEM: return "device memory allocation failed";
    default: return "unknown status";
  }
}

const char* psa_gpu_last_error(void* vctx) {
  if (!vctx) return "";
  return static_cast<Context*>(vctx)->err.c_str();
}

int psa_gpu_init(int ngpus, const int* device_ids, void** out) {
  if (!out || ngpus < 1) return PSA_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) return PSA_ERR_NODEVICE;
  if (ngpus > count) return PSA_ERR_NODEVICE;
  Context* ctx = new Context();
  ctx->devs.resize(ngpus);
  for (int i = 0; i < ngpus; ++i) {
    Device& d = ctx->devs[i];
    d.dev = device_ids ? device_ids[i] : i;
    cudaDeviceProp prop;
    if (d.dev < 0 || d.dev >= count || cudaGetDeviceProperties(&prop, d.dev) != cudaSuccess || prop.major != 10) {
      for (auto& dd : ctx->devs) free_device(dd);
      delete ctx;
      return PSA_ERR_NODEVICE;
    }
    d.num_sms = prop.multiProcessorCount;
    bool ok = cudaSetDevice(d.dev) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&d.state, 4 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&d.counts, PSA_NCOUNTS * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaHostAlloc((void**)&d.host_state, 4 * sizeof(int), cudaHostAllocMapped) == cudaSuccess;
    ok = ok && cudaHostGetDevicePointer((void**)&d.host_state_dev, d.host_state, 0) == cudaSuccess;
    ok = ok && cudaEventCreate(&d.g0) == cudaSuccess && cudaEventCreate(&d.g1) == cudaSuccess;
    if (!ok) {
      for (auto& dd : ctx->devs) free_device(dd);
      delete ctx;
      return PSA_ERR_CUDA;
    }
  }
  *out = ctx;
  return PSA_OK;
}

int psa_gpu_free(void* vctx) {
  if (!vctx) return PSA_ERR_ARG;
  Context* ctx = static_cast<Context*>(vctx);
  if (!ctx->devs.empty()) cudaSetDevice(ctx->devs[0].dev);
  for (double* p : ctx->gather_buf)
    if (p) cudaFree(p);
  for (int* p : ctx->gather_ibuf)
    if (p) cudaFree(p);
  for (ncclComm_t c : ctx->comms)
    if (c) ctx->nccl.CommDestroy(c);
  for (auto& d : ctx->devs) free_device(d);
  delete ctx;
  return PSA_OK;
}

static int set_matrix_common(Context* ctx, const double* T, int64_t ldt, int m, int n, bool on_device) {
  if (!T || m < n || n < 1 || ldt < m) return PSA_ERR_ARG;
  bool is_hess = false;
  if (m == n + 1 && m <= 200 && n >= 55) {  // rect_qr candidates: is the input upper Hessenberg?
    std::vector<double> host;
    const double* Th = T;
    if (on_device) {
      host.resize((size_t)m * n * 2);
      PSA_CUDA(cudaSetDevice(ctx->devs[0].dev));
      PSA_CUDA(cudaMemcpy2D(host.data(), (size_t)m * 16, T, (size_t)ldt * 16, (size_t)m * 16, n, cudaMemcpyDeviceToHost));
      Th = host.data();
    }
    const int64_t ld = on_device ? m : ldt;
    is_hess = true;
    for (int j = 0; j < n && is_hess; ++j)
      for (int i = j + 2; i < m; ++i)
        if (Th[2 * (i + j * ld)] != 0.0 || Th[2 * (i + j * ld) + 1] != 0.0) {
          is_hess = false;
          break;
        }
  }
  const int algo = classify(m, n, is_hess);
  if (algo == PSA_ALGO_NONE) return PSA_ERR_UNSUPPORTED;
  ctx->algo = PSA_ALGO_NONE;
  ctx->m = m;
  ctx->n = n;
  ctx->n_pad = (n + MB - 1) / MB * MB;
  ctx->nblk = ctx->n_pad / MB;
  // raw copy to device 0, then replicate
  Device& d0 = ctx->devs[0];
  PSA_CUDA(cudaSetDevice(d0.dev));
  int rc;
  if (d0.T_cap < (size_t)m * n) {
    if ((rc = dev_alloc(ctx, &d0.T, (size_t)m * n))) return rc;
    d0.T_cap = (size_t)m * n;
  }
  PSA_CUDA(cudaMemcpy2DAsync(d0.T, (size_t)m * 16, T, (size_t)ldt * 16, (size_t)m * 16, n,
                             on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, d0.stream));
  PSA_CUDA(cudaStreamSynchronize(d0.stream));
  if (ctx->devs.size() > 1) {
    // replicate T with one NCCL broadcast over NVLink (root = device 0)
    if ((rc = ensure_comms(ctx))) return rc;
    for (size_t i = 1; i < ctx->devs.size(); ++i) {
      Device& d = ctx->devs[i];
      PSA_CUDA(cudaSetDevice(d.dev));
      if (d.T_cap < (size_t)m * n) {
        if ((rc = dev_alloc(ctx, &d.T, (size_t)m * n))) return rc;
        d.T_cap = (size_t)m * n;
      }
    }
    PSA_NCCL(ctx->nccl.GroupStart());
    for (size_t i = 0; i < ctx->devs.size(); ++i) {
      Device& d = ctx->devs[i];
      PSA_NCCL(ctx->nccl.Broadcast(d0.T, d.T, (size_t)m * n * 2, ncclDouble, 0, ctx->comms[i], d.stream));
    }
    PSA_NCCL(ctx->nccl.GroupEnd());
  }
  ctx->algo = algo;
  for (auto& d : ctx->devs)
    if ((rc = prepare_matrix(ctx, d))) {
      ctx->algo = PSA_ALGO_NONE;
      return rc;
    }
  return PSA_OK;
}

int psa_gpu_set_matrix(void* vctx, const double* T, int64_t ldt, int m, int n) {
  if (!vctx) return PSA_ERR_ARG;
  return set_matrix_common(static_cast<Context*>(vctx), T, ldt, m, n, false);
}

int psa_gpu_set_matrix_device(void* vctx, const double* d_T, int64_t ldt, int m, int n) {
  if (!vctx) return PSA_ERR_ARG;
  return set_matrix_common(static_cast<Context*>(vctx), d_T, ldt, m, n, true);
}

int psa_gpu_grid(void* vctx, const double* x, int lx, const double* y, int ly, const double* q0, double tol, int maxit,
                 double bigsigma, double* Z, int32_t* iters, int64_t* counts, int* algo, volatile int* cancel) {
  if (!vctx) return PSA_ERR_ARG;
  Context* ctx = static_cast<Context*>(vctx);
  if (ctx->algo == PSA_ALGO_NONE) return PSA_ERR_NOMATRIX;
  if (!x || !y || !q0 || !Z || lx < 1 || ly < 1 || maxit < 1) return PSA_ERR_ARG;
  if (algo) *algo = ctx->algo;
  const int ng = (int)ctx->devs.size();
  std::vector<double*> Zout(ng);
  std::vector<int*> itout(ng);
  int rc;
  for (int i = 0; i < ng; ++i) {
    Device& d = ctx->devs[i];
    PSA_CUDA(cudaSetDevice(d.dev));
    const int rows = (ly - i + ng - 1) / ng;
    const size_t P = (size_t)rows * lx;
    if ((rc = ensure_cap(ctx, &d.x, &d.x_cap, (size_t)lx))) return rc;
    if ((rc = ensure_cap(ctx, &d.y, &d.y_cap, (size_t)ly))) return rc;
    if ((rc = ensure_cap(ctx, &d.q0, &d.q0_cap, (size_t)ctx->n))) return rc;
    if ((rc = ensure_cap(ctx, &d.Z, &d.Z_cap, P))) return rc;
    if ((rc = ensure_cap(ctx, &d.iters, &d.it_cap, P))) return rc;
    PSA_CUDA(cudaMemcpyAsync(d.x, x, lx * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    PSA_CUDA(cudaMemcpyAsync(d.y, y, ly * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    PSA_CUDA(cudaMemcpyAsync(d.q0, q0, (size_t)ctx->n * 16, cudaMemcpyHostToDevice, d.stream));
    Zout[i] = d.Z;
    itout[i] = d.iters;
  }
  GridArgs g{lx, ly, tol, maxit, bigsigma, cancel};
  rc = run_grid(ctx, g, Zout, itout, counts);
  if (rc != PSA_OK) return rc;
  // gather: device i owns rows i, i+ng, ... ; local layout is rows_i x lx column-major
  Device& d0 = ctx->devs[0];
  if (ng == 1) {
    const size_t P = (size_t)ly * lx;
    PSA_CUDA(cudaSetDevice(d0.dev));
    PSA_CUDA(cudaMemcpy(Z, d0.Z, P * sizeof(double), cudaMemcpyDeviceToHost));
    if (iters) PSA_CUDA(cudaMemcpy(iters, d0.iters, P * sizeof(int), cudaMemcpyDeviceToHost));
    return PSA_OK;
  }
  // the sigma_min tiles travel to device 0 over NVLink (grouped ncclSend/ncclRecv), then one D2H per tile
  PSA_CUDA(cudaSetDevice(d0.dev));
  for (int i = 1; i < ng; ++i) {
    const size_t P = (size_t)((ly - i + ng - 1) / ng) * lx;
    if (ctx->gather_cap[i] < P) {
      if ((rc = dev_alloc(ctx, &ctx->gather_buf[i], P))) return rc;
      if ((rc = dev_alloc(ctx, &ctx->gather_ibuf[i], P))) return rc;
      ctx->gather_cap[i] = P;
    }
  }
  PSA_NCCL(ctx->nccl.GroupStart());
  for (int i = 1; i < ng; ++i) {
    Device& d = ctx->devs[i];
    const size_t P = (size_t)((ly - i + ng - 1) / ng) * lx;
    if (P == 0) continue;
    PSA_NCCL(ctx->nccl.Send(d.Z, P, ncclDouble, 0, ctx->comms[i], d.stream));
    PSA_NCCL(ctx->nccl.Recv(ctx->gather_buf[i], P, ncclDouble, i, ctx->comms[0], d0.stream));
    if (iters) {
      PSA_NCCL(ctx->nccl.Send(d.iters, P, ncclInt32, 0, ctx->comms[i], d.stream));
      PSA_NCCL(ctx->nccl.Recv(ctx->gather_ibuf[i], P, ncclInt32, i, ctx->comms[0], d0.stream));
    }
  }
  PSA_NCCL(ctx->nccl.GroupEnd());
  for (int i = 1; i < ng; ++i) {
    PSA_CUDA(cudaSetDevice(ctx->devs[i].dev));
    PSA_CUDA(cudaStreamSynchronize(ctx->devs[i].stream));
  }
  PSA_CUDA(cudaSetDevice(d0.dev));
  PSA_CUDA(cudaStreamSynchronize(d0.stream));
  std::vector<double> zb;
  std::vector<int> ib;
  for (int i = 0; i < ng; ++i) {
    const int rows = (ly - i + ng - 1) / ng;
    const size_t P = (size_t)rows * lx;
    if (P == 0) continue;
    zb.resize(P);
    PSA_CUDA(cudaMemcpy(zb.data(), i == 0 ? d0.Z : ctx->gather_buf[i], P * sizeof(double), cudaMemcpyDeviceToHost));
    for (int k = 0; k < lx; ++k)
      for (int jr = 0; jr < rows; ++jr) Z[(size_t)k * ly + i + (size_t)jr * ng] = zb[(size_t)k * rows + jr];
    if (iters) {
      ib.resize(P);
      PSA_CUDA(cudaMemcpy(ib.data(), i == 0 ? d0.iters : ctx->gather_ibuf[i], P * sizeof(int), cudaMemcpyDeviceToHost));
      for (int k = 0; k < lx; ++k)
        for (int jr = 0; jr < rows; ++jr) iters[(size_t)k * ly + i + (size_t)jr * ng] = ib[(size_t)k * rows + jr];
    }
  }
  return PSA_OK;
}

int psa_gpu_grid_device(void* vctx, const double* d_x, int lx, const double* d_y, int ly, const double* d_q0, double tol,
                        int maxit, double bigsigma, double* d_Z, int32_t* d_iters, int64_t* counts, int* algo,
                        volatile int* cancel) {
  if (!vctx) return PSA_ERR_ARG;
  Context* ctx = static_cast<Context*>(vctx);
  if (ctx->algo == PSA_ALGO_NONE) return PSA_ERR_NOMATRIX;
  if (ctx->devs.size() != 1) return PSA_ERR_UNSUPPORTED;
  if (!d_x || !d_y || !d_q0 || !d_Z || lx < 1 || ly < 1 || maxit < 1) return PSA_ERR_ARG;
  if (algo) *algo = ctx->algo;
  Device& d = ctx->devs[0];
  PSA_CUDA(cudaSetDevice(d.dev));
  // borrow the caller's buffers for this call
  double *sx = d.x, *sy = d.y;
  double2* sq = d.q0;
  d.x = const_cast<double*>(d_x);
  d.y = const_cast<double*>(d_y);
  d.q0 = reinterpret_cast<double2*>(const_cast<double*>(d_q0));
  std::vector<double*> Zout{d_Z};
  std::vector<int*> itout{d_iters};
  GridArgs g{lx, ly, tol, maxit, bigsigma, cancel};
  int rc = run_grid(ctx, g, Zout, itout, counts);
  d.x = sx;
  d.y = sy;
  d.q0 = sq;
  return rc;
}

int psa_gpu_apply_resolvent(void* vctx, const double* z, int ns, const double* Q, double* V) {
  if (!vctx) return PSA_ERR_ARG;
  Context* ctx = static_cast<Context*>(vctx);
  if (ctx->algo != PSA_ALGO_SQ_LANC) return ctx->algo == PSA_ALGO_NONE ? PSA_ERR_NOMATRIX : PSA_ERR_UNSUPPORTED;
  if (!z || !Q || !V || ns < 1) return PSA_ERR_ARG;
  Device& d = ctx->devs[0];
  PSA_CUDA(cudaSetDevice(d.dev));
  int rc = set_solve_attr(ctx);
  if (rc) return rc;
  const int W = (ns + APAD - 1) / APAD * APAD;  // padded active list
  if ((rc = ensure_slots(ctx, d, std::max(W, d.W)))) return rc;
  struct Scratch {  // temporary device buffers, released on every exit path
    double2 *Q = nullptr, *z = nullptr, *V = nullptr;
    ~Scratch() {
      cudaFree(Q);
      cudaFree(z);
      cudaFree(V);
    }
  } tmp;
  double2 *&dQ = tmp.Q, *&dz = tmp.z, *&dV = tmp.V;
  const size_t n = ctx->n;
  if ((rc = dev_alloc(ctx, &dQ, ns * n))) return rc;
  if ((rc = dev_alloc(ctx, &dz, (size_t)ns))) return rc;
  if ((rc = dev_alloc(ctx, &dV, ns * n))) return rc;
  PSA_CUDA(cudaMemcpyAsync(dQ, Q, ns * n * 16, cudaMemcpyHostToDevice, d.stream));
  PSA_CUDA(cudaMemcpyAsync(dz, z, (size_t)ns * 16, cudaMemcpyHostToDevice, d.stream));
  load_slots_kernel<<<W, 256, 0, d.stream>>>(ns, ctx->n, ctx->n_pad, dQ, dz, d.Q0, d.slot_z, d.slot_sel, d.active, d.W,
                                             d.state);
  EventPair& ep = event_pair(d, 0);
  PSA_CUDA(cudaEventRecord(ep.a, d.stream));
  launch_solve(ctx, d, ns, pick_panel_width(ns, d.num_sms));
  PSA_CUDA(cudaEventRecord(ep.b, d.stream));
  store_slots_kernel<<<ns, 256, 0, d.stream>>>(ctx->n, ctx->n_pad, d.Wv, dV);
  PSA_CUDA(cudaGetLastError());
  PSA_CUDA(cudaMemcpyAsync(V, dV, ns * n * 16, cudaMemcpyDeviceToHost, d.stream));
  PSA_CUDA(cudaStreamSynchronize(d.stream));
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, ep.a, ep.b);
    ctx->stats[PSA_STAT_SOLVE_MS] = ms;
    ctx->stats[PSA_STAT_SOLVE_LAUNCHES] = 1;
    ctx->stats[PSA_STAT_FLOPS] = 8.0 * (double)n * (double)n * ns;
  }
  return PSA_OK;
}

int psa_gpu_last_stats(void* vctx, double* stats) {
  if (!vctx || !stats) return PSA_ERR_ARG;
  Context* ctx = static_cast<Context*>(vctx);
  for (int i = 0; i < PSA_NSTATS; ++i) stats[i] = ctx->stats[i];
  return PSA_OK;
}

}  // extern "C"
