// libpsagpu.so -- host side of the C ABI declared in include/psa_gpu.h.
// One host thread drives every device of the context through per-device streams.  All launches are asynchronous;
// the host learns the progress of each device from one 8-byte word the scheduler kernel publishes in mapped pinned
// memory and never blocks on a stream inside the tick loop (see run_grid).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <time.h>

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/psa_gpu.h"
#include "psa_hess.cuh"
#include "psa_nccl.h"
#include "psa_post.cuh"
#include "psa_project.cuh"
#include "psa_square.cuh"
#include "psa_cluster.cuh"
#include "psa_point.cuh"
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
  double2* T = nullptr;     // raw copy of the matrix as set (m_full x n_full, ld = m_full)
  double2* Tp = nullptr;    // projected matrix (npick x npick) after psa_gpu_project
  double2* Tcur = nullptr;  // the matrix the kernels work on: T, or Tp while a projection is active
  double2* Twork = nullptr; // device 0: scratch copy for the Schur reordering
  int *proj_sel = nullptr, *proj_start = nullptr;
  ProjRot* proj_rot = nullptr;
  size_t Tp_cap = 0, Twork_cap = 0, proj_cap = 0;
  double2* packA = nullptr;
  double2* packB = nullptr;
  size_t T_cap = 0, pack_cap = 0, diag_cap = 0;
  double2* diagT = nullptr;  // eigenvalues (diagonal of T), for the scheduling-order predictor
  // scheduling order
  float *key_in = nullptr, *key_out = nullptr, *key_static = nullptr;
  int *ord_in = nullptr, *ord_out = nullptr, *prog = nullptr;
  void* sort_tmp = nullptr;
  size_t ord_cap = 0, sort_tmp_bytes = 0;
  // slots
  int W = 0;
  int hist = 0;
  size_t vec_elems = 0;
  double2 *Q0 = nullptr, *Q1 = nullptr, *Wv = nullptr;
  int *slot_point = nullptr, *slot_iter = nullptr, *slot_sel = nullptr, *slot_batch = nullptr;
  double2* slot_z = nullptr;
  double *slot_beta = nullptr, *slot_sigold = nullptr, *hist_alpha = nullptr, *hist_beta = nullptr;
  int *active = nullptr, *fresh = nullptr, *state = nullptr;
  unsigned long long* host_state = nullptr;      // pinned + mapped: [0] (tick << 32) | n_active, [1] (tick << 32) | n_fresh,
                                                 // [2] stop word of the point-resident kernel (cancel flag)
  unsigned long long* host_state_dev = nullptr;  // device alias
  unsigned long long* counts = nullptr;
  unsigned long long* scratch64 = nullptr;  // 4 words: structure-check counter, post-processing reductions
  // point-resident Lanczos kernel (small n): per-warp alpha / beta histories
  double* pt_hist = nullptr;
  size_t pt_hist_cap = 0;
  // HessQR workspace
  double2* R = nullptr;  // per-slot triangular factors
  size_t R_elems = 0;
  // grid buffers (owned unless device API)
  double *x = nullptr, *y = nullptr, *Z = nullptr;
  double2* q0 = nullptr;  // [nbatch][n_pad], zero padded
  int* row_batch = nullptr;
  int* iters = nullptr;
  size_t x_cap = 0, y_cap = 0, q0_cap = 0, Z_cap = 0, it_cap = 0, rb_cap = 0;
  // device 0 only: assembled result of a multi-device grid, post-processing output, pseudomode basis
  double* Zfull = nullptr;
  int* itfull = nullptr;
  double* post = nullptr;
  double2* basis = nullptr;
  double* yvec = nullptr;
  size_t Zfull_cap = 0, itfull_cap = 0, post_cap = 0, basis_cap = 0, yvec_cap = 0;
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
  int m = 0, n = 0, n_pad = 0, nblk = 0, algo = PSA_ALGO_NONE;  // current (possibly projected) matrix
  int m_full = 0, n_full = 0, algo_full = PSA_ALGO_NONE;         // as set by psa_gpu_set_matrix
  int minlancs = 55, maxstdqr = 200;  // ComputeThresholds, compute.jl:20-27
  double stats[PSA_NSTATS] = {0};
  bool solve_attr_set = false;
  // result of the last grid call, resident on device 0 (ly x lx column-major)
  const double* last_Z = nullptr;
  int last_lx = 0, last_ly = 0;
};

template <typename T>
int dev_alloc(Context* ctx, T** p, size_t count) {
  if (*p) {
    cudaFree(*p);
    *p = nullptr;
  }
  PSA_CUDA(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
  return PSA_OK;
}

template <typename T>
int ensure_cap(Context* ctx, T** p, size_t* cap, size_t want) {
  if (*cap >= want && *p) return PSA_OK;
  *cap = 0;
  int rc = dev_alloc(ctx, p, want);
  if (rc) return rc;
  *cap = want;
  return PSA_OK;
}

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
void free_ptr(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

// slot buffers: released as a unit, so that a failed (re)allocation can never leave a half-sized set behind
void free_slots(Device& d) {
  free_ptr(d.Q0);
  free_ptr(d.Q1);
  free_ptr(d.Wv);
  free_ptr(d.slot_point);
  free_ptr(d.slot_iter);
  free_ptr(d.slot_sel);
  free_ptr(d.slot_batch);
  free_ptr(d.slot_z);
  free_ptr(d.slot_beta);
  free_ptr(d.slot_sigold);
  free_ptr(d.hist_alpha);
  free_ptr(d.hist_beta);
  free_ptr(d.active);
  free_ptr(d.fresh);
  free_ptr(d.R);
  d.W = 0;
  d.hist = 0;
  d.vec_elems = 0;
  d.R_elems = 0;
}

void free_device(Device& d) {
  cudaSetDevice(d.dev);
  free_slots(d);
  void* ptrs[] = {d.key_static, d.prog, d.diagT, d.key_in, d.key_out, d.ord_in, d.ord_out, d.sort_tmp, d.T,     d.packA, d.packB,
                  d.state, d.counts, d.scratch64, d.x,    d.y,       d.Z,        d.q0,    d.row_batch, d.iters,
                  d.Zfull, d.itfull, d.post,      d.basis, d.yvec, d.Tp, d.Twork, d.proj_sel, d.proj_start, d.proj_rot, d.pt_hist};
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

bool is_hess_algo(int algo) { return algo == PSA_ALGO_HESSQR || algo == PSA_ALGO_RECT_QR; }

int ensure_slots(Context* ctx, Device& d, int want_W, int hist) {
  const size_t vec_elems = (size_t)ctx->n_pad;
  const bool hess = is_hess_algo(ctx->algo);
  const size_t r_elems = hess ? hess_R_elems(ctx->n) : 0;
  if (d.W >= want_W && d.hist >= hist && d.vec_elems == vec_elems && d.R_elems == r_elems) return PSA_OK;
  PSA_CUDA(cudaSetDevice(d.dev));
  const int W = std::max(want_W, d.vec_elems == vec_elems && d.R_elems == r_elems ? d.W : 0);
  hist = std::max(hist, d.hist);
  free_slots(d);  // also drops a HessQR factor store (up to 24 GB) when the new matrix does not need one
  int rc = PSA_OK;
  auto A = [&](auto** p, size_t count) {
    if (rc == PSA_OK) rc = dev_alloc(ctx, p, count);
  };
  A(&d.Q0, (size_t)(W + 1) * vec_elems);
  A(&d.Q1, (size_t)(W + 1) * vec_elems);
  A(&d.Wv, (size_t)(W + 1) * vec_elems);
  A(&d.slot_point, (size_t)W + 1);
  A(&d.slot_iter, (size_t)W + 1);
  A(&d.slot_sel, (size_t)W + 1);
  A(&d.slot_batch, (size_t)W + 1);
  A(&d.slot_z, (size_t)W + 1);
  A(&d.slot_beta, (size_t)W + 1);
  A(&d.slot_sigold, (size_t)W + 1);
  A(&d.hist_alpha, (size_t)(W + 1) * hist);
  A(&d.hist_beta, (size_t)(W + 1) * hist);
  A(&d.active, (size_t)W + 2 * APAD);
  A(&d.fresh, (size_t)W + 1);
  if (hess) A(&d.R, (size_t)(W + 1) * r_elems);
  if (rc != PSA_OK) {
    free_slots(d);
    return rc;
  }
  // scratch slot W: benign, finite contents
  cudaError_t e = cudaMemsetAsync(d.Q0 + (size_t)W * vec_elems, 0, vec_elems * sizeof(double2), d.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d.Q1 + (size_t)W * vec_elems, 0, vec_elems * sizeof(double2), d.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d.Wv + (size_t)W * vec_elems, 0, vec_elems * sizeof(double2), d.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d.slot_sel, 0, (W + 1) * sizeof(int), d.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d.slot_iter, 0xff, (W + 1) * sizeof(int), d.stream);  // scratch slot: never "fresh"
  if (e == cudaSuccess) e = cudaMemsetAsync(d.slot_batch, 0, (W + 1) * sizeof(int), d.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d.slot_z, 0, (W + 1) * sizeof(double2), d.stream);
  if (e == cudaSuccess && hess) e = cudaMemsetAsync(d.R + (size_t)W * r_elems, 0, r_elems * sizeof(double2), d.stream);
  if (e != cudaSuccess) {
    free_slots(d);
    ctx->err = std::string("slot initialisation failed: ") + cudaGetErrorString(e);
    return PSA_ERR_CUDA;
  }
  d.W = W;
  d.hist = hist;
  d.vec_elems = vec_elems;
  d.R_elems = r_elems;
  return PSA_OK;
}

int slot_capacity(const Context* ctx, const Device& d, long long points) {
  long long cap;
  if (is_hess_algo(ctx->algo)) {
    cap = hess_slot_capacity(ctx->n, d.num_sms);
  } else {
    // exactly ONE round of resident solve CTAs (2 per SM, NS shifts each): with more slots than that every tick whose
    // active count is not a multiple of a round pays for a partially filled extra round.  Replaying the tick loop on
    // the measured iteration counts of the bench workload (tools/sim_ticks.py) gives 0.87 strong-scaling efficiency at
    // 8 GPUs with two rounds of slots and 0.97 with one; a single GPU is unaffected (twice the ticks, half as long).
    cap = 2LL * d.num_sms * NS;
    if (solve_flavor() == 0) cap = 4LL * d.num_sms * NS;  // the round-1 kernel keeps 4 CTAs per SM resident
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
    int rc = solve_set_attrs();
    if (rc == 0) rc = cluster_set_attrs();
    if (rc == 0 && cudaFuncSetAttribute(lanczos_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        LZW_WARPS * 2 * LZW_MAX_HIST * (int)sizeof(double)) != cudaSuccess)
      rc = -1;
    if (rc == 0 && cudaFuncSetAttribute(lanczos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        2 * HIST_MAX * (int)sizeof(double)) != cudaSuccess)
      rc = -1;
    if (rc == 0) rc = hess_set_attrs();
    if (rc == 0) rc = point_set_attrs();
    if (rc == 0 && cudaFuncSetAttribute(svd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) rc = -1;
    if (rc != 0) {
      ctx->err = "cudaFuncSetAttribute failed";
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
  sp.iter = d.slot_iter;
  sp.batch = d.slot_batch;
  sp.q0 = d.q0;
  const char* ep = getenv("PSA_PAIR");  // experiments only
  sp.pair = ep ? atoi(ep) : 1;
  return sp;
}

// pack / prepare the device-resident matrix on one device (d.Tcur holds the raw ctx->m x ctx->n matrix, ld = ctx->m)
int prepare_matrix(Context* ctx, Device& d) {
  PSA_CUDA(cudaSetDevice(d.dev));
  if (ctx->algo == PSA_ALGO_SQ_LANC) {
    const long long ntiles = tiles_per_phase(ctx->nblk);
    int rc;
    if (d.pack_cap < (size_t)ntiles * TILE_ELEMS) {  // buffers are reused across set_matrix calls of the same size
      d.pack_cap = 0;
      if ((rc = dev_alloc(ctx, &d.packA, (size_t)ntiles * TILE_ELEMS))) return rc;
      if ((rc = dev_alloc(ctx, &d.packB, (size_t)ntiles * TILE_ELEMS))) return rc;
      d.pack_cap = (size_t)ntiles * TILE_ELEMS;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, d.stream);
    if ((rc = ensure_cap(ctx, &d.diagT, &d.diag_cap, (size_t)ctx->n))) return rc;
    diag_kernel<<<(ctx->n + 255) / 256, 256, 0, d.stream>>>(d.Tcur, ctx->m, ctx->n, d.diagT);
    const int grid = (int)std::min<long long>(2 * ntiles, 148LL * 32);
    pack_kernel<<<grid, 256, 0, d.stream>>>(d.Tcur, ctx->m, ctx->n, ctx->n_pad, ctx->nblk, d.packA, d.packB);
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

// Which branch of psacore the shape selects (compute.jl:478, 534, 565, 579, 588-590), with the context's thresholds.
int classify(const Context* ctx, int m, int n) {
  if (n < ctx->minlancs) return (long long)m * n <= SVD_MAX_ELEMS ? PSA_ALGO_SVD : PSA_ALGO_NONE;  // compute.jl:478
  if (m == n) return PSA_ALGO_SQ_LANC;                                                             // compute.jl:534,565,594
  if (m == n + 1 && n <= HQ_MAX_N) return m > ctx->maxstdqr ? PSA_ALGO_HESSQR : PSA_ALGO_RECT_QR;  // compute.jl:579,588
  return PSA_ALGO_NONE;
}

struct GridArgs {
  int lx, ly;
  double tol;
  int maxit;
  double bigsigma;
  volatile int* cancel;
};

void short_pause(unsigned& idle) {
  if (++idle < 20000u) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  } else {
    timespec ts{0, 20000};  // 20 us: long ticks (tens of ms) do not deserve a spinning host core
    nanosleep(&ts, nullptr);
  }
}

// Per-device progress of the tick loop.  A tick = [sched(k)] -> solve(k) -> lanczos(k).  The scheduler kernel of
// tick k publishes (k, n_active) in mapped host memory.  The host enqueues "group k" = {solve(k), lanczos(k),
// sched(k+1)} as soon as tick k - lag has been published:
//   lag = 0  exact mode: grid sizes and the panel width come from tick k's own counts (heavy ticks: the ~20 us the
//            GPU waits for the host per tick are noise against milliseconds of work);
//   lag > 0  pipelined mode for light ticks: the host runs up to `lag` ticks ahead with grid sizes taken from the
//            newest published count, which is an upper bound because the active count never grows after the first
//            tick; every kernel reads the exact count from device memory and surplus blocks exit at once.
// Either way the host never synchronises a stream inside the loop, so several devices progress independently.
struct TickState {
  unsigned pub = 0;        // newest tick seen published
  int pub_active = 0;      // its n_active
  int pub_fresh = -1;      // its n_fresh when read consistently, else -1
  unsigned groups = 0;     // groups enqueued so far (sched(groups+1) is in the stream)
  int bound = 0;           // upper bound of n_active for the next group
  bool done = false;
  size_t nev = 0;
};

// Run the tick loop on every device.  Per device: d.x, d.y, d.q0 (and d.row_batch) are resident, Zout / itout are
// the (local point order) outputs.  Rows are dealt cyclically: device i owns rows i, i+ng, ...
int run_grid(Context* ctx, const GridArgs& g, const std::vector<const int*>& row_batch, std::vector<double*>& Zout,
             std::vector<int*>& itout, int64_t* counts) {
  const int ng = (int)ctx->devs.size();
  int rc = set_solve_attr(ctx);
  if (rc) return rc;
  if (g.maxit > HIST_MAX - 2) return PSA_ERR_ARG;
  const int hist = std::max(HIST_MIN, (g.maxit + 2 + 31) / 32 * 32);
  // small square matrices: one persistent launch, a warp per grid point (psa_point.cuh); shape-only decision
  const int pt_w = ctx->algo == PSA_ALGO_SQ_LANC && !getenv("PSA_NO_POINT_KERNEL") ? pt_warps(ctx->n, hist) : 0;
  const bool point_mode = pt_w > 0;
  std::vector<int> rows(ng), P(ng);
  std::vector<const int*> order(ng, nullptr);
  std::vector<TickState> ts(ng);
  std::vector<int> dyn_stride(ng, 0);   // lattice stride of the dynamic queue order (0: static order)
  std::vector<long long> started(ng, 0);  // host view of the points started so far (exact mode: from the published counts)
  double launches = 0;
  for (int i = 0; i < ng; ++i) {
    Device& d = ctx->devs[i];
    rows[i] = (g.ly - i + ng - 1) / ng;
    P[i] = rows[i] * g.lx;
    PSA_CUDA(cudaSetDevice(d.dev));
    if (ctx->algo != PSA_ALGO_SVD && !point_mode) {  // the SVD branch has no Lanczos state, the point kernel keeps it in registers
      if ((rc = ensure_slots(ctx, d, slot_capacity(ctx, d, P[i]), hist))) return rc;
      PSA_CUDA(cudaMemsetAsync(d.slot_point, 0xff, (d.W + 1) * sizeof(int), d.stream));
    }
    PSA_CUDA(cudaMemsetAsync(d.state, 0, 8 * sizeof(int), d.stream));
    PSA_CUDA(cudaMemsetAsync(d.counts, 0, PSA_NCOUNTS * sizeof(unsigned long long), d.stream));
    d.host_state[0] = 0;
    d.host_state[1] = 0;
    d.host_state[2] = 0;
    PSA_CUDA(cudaEventRecord(d.g0, d.stream));
    if (P[i] == 0) ts[i].done = true;
    const long long in_flight = point_mode ? (long long)d.num_sms * pt_w : (long long)d.W;
    if (ctx->algo == PSA_ALGO_SQ_LANC && P[i] > in_flight && !getenv("PSA_NATURAL_ORDER")) {
      // more points than slots: queue them longest-job-first (see predict_kernel)
      if (d.ord_cap < (size_t)P[i]) {
        d.ord_cap = 0;
        if ((rc = dev_alloc(ctx, &d.key_in, (size_t)P[i]))) return rc;
        if ((rc = dev_alloc(ctx, &d.key_out, (size_t)P[i]))) return rc;
        if ((rc = dev_alloc(ctx, &d.key_static, (size_t)P[i]))) return rc;
        if ((rc = dev_alloc(ctx, &d.ord_in, (size_t)P[i]))) return rc;
        if ((rc = dev_alloc(ctx, &d.ord_out, (size_t)P[i]))) return rc;
        if ((rc = dev_alloc(ctx, &d.prog, (size_t)P[i]))) return rc;
        d.ord_cap = (size_t)P[i];
      }
      // dynamic queue order (see rekey_kernel): only worth its two small launches per tick when ticks are heavy
      const bool dynamic = !point_mode && !getenv("PSA_STATIC_ORDER") && 8.0 * ctx->n * (double)ctx->n * d.W / 2.5e10 >= 3.0;
      if (dynamic) {
        dyn_stride[i] = std::max(1, (int)std::sqrt((double)P[i] / (0.5 * d.W)));
        PSA_CUDA(cudaMemsetAsync(d.prog, 0, (size_t)P[i] * sizeof(int), d.stream));
        PSA_CUDA(cudaMemsetAsync(d.scratch64 + 3, 0, sizeof(unsigned long long), d.stream));
      }
      predict_kernel<<<(P[i] + 255) / 256, 256, 0, d.stream>>>(P[i], rows[i], i, ng, d.x, d.y, d.diagT, ctx->n,
                                                               dynamic ? d.key_static : d.key_in, d.ord_in,
                                                               dynamic ? reinterpret_cast<unsigned*>(d.scratch64 + 3) : nullptr);
      if (dynamic) {  // first order: the lattice, then the static keys
        RekeyParams rp{P[i], rows[i], g.lx, dyn_stride[i], d.prog, d.key_static, reinterpret_cast<unsigned*>(d.scratch64 + 3),
                       d.key_in, d.ord_in, d.state};
        rekey_kernel<<<(P[i] + 255) / 256, 256, 0, d.stream>>>(rp);
        launches += 1;
      }
      size_t need = 0;
      cub::DeviceRadixSort::SortPairsDescending(nullptr, need, d.key_in, d.key_out, d.ord_in, d.ord_out, P[i], 0, 32,
                                                d.stream);
      if (need > d.sort_tmp_bytes) {
        if (d.sort_tmp) cudaFree(d.sort_tmp);
        d.sort_tmp = nullptr;
        d.sort_tmp_bytes = 0;
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
      if (P[i] == 0) continue;
      SvdParams sp;
      sp.m = ctx->m;
      sp.n = ctx->n;
      sp.T = d.Tcur;
      sp.P = P[i];
      sp.rows_local = rows[i];
      sp.row0 = i;
      sp.row_step = ng;
      sp.x = d.x;
      sp.y = d.y;
      sp.Z = Zout[i];
      sp.iters = itout[i];
      sp.counts = d.counts;
      EventPair& ep = event_pair(d, ts[i].nev++);
      PSA_CUDA(cudaEventRecord(ep.a, d.stream));
      svd_kernel<<<std::min(P[i], 8 * d.num_sms), SVD_THREADS, svd_smem(ctx->m, ctx->n), d.stream>>>(sp);
      PSA_CUDA(cudaEventRecord(ep.b, d.stream));
      PSA_CUDA(cudaGetLastError());
      launches += 1;
      ts[i].done = true;
    }
  }

  bool cancelled = false;
  if (point_mode) {
    // one launch per device over all of its points; warps pull points from an atomic queue
    if (g.cancel && *g.cancel == 1) cancelled = true;
    for (int i = 0; i < ng; ++i) ts[i].done = true;
    for (int i = 0; i < ng && !cancelled; ++i) {
      Device& d = ctx->devs[i];
      PSA_CUDA(cudaSetDevice(d.dev));
      if (P[i] == 0) continue;
      const int ctas = std::min(d.num_sms, (P[i] + pt_w - 1) / pt_w);
      const size_t pt_slots = (size_t)ctas * pt_w;  // per warp: alpha and beta histories, then the two Lanczos vectors
      if ((rc = ensure_cap(ctx, &d.pt_hist, &d.pt_hist_cap, pt_slots * (2 * (size_t)hist + 4 * PT_MAX_NPL * 32)))) return rc;
      PointParams pp;
      pp.n = ctx->n;
      pp.T = d.Tcur;
      pp.ldt = ctx->m;
      pp.P = P[i];
      pp.rows_local = rows[i];
      pp.row0 = i;
      pp.row_step = ng;
      pp.x = d.x;
      pp.y = d.y;
      pp.order = order[i];
      pp.row_batch = row_batch[i];
      pp.q0 = d.q0;
      pp.n_pad = ctx->n_pad;
      pp.tol = g.tol;
      pp.maxit = g.maxit;
      pp.bigsigma = g.bigsigma;
      pp.Z = Zout[i];
      pp.iters = itout[i];
      pp.counts = d.counts;
      pp.queue = d.state + 6;
      pp.stop = g.cancel ? d.host_state_dev + 2 : nullptr;
      pp.hist_alpha = d.pt_hist;
      pp.hist_beta = d.pt_hist + pt_slots * hist;
      pp.qold = reinterpret_cast<double2*>(d.pt_hist + 2 * pt_slots * hist);
      pp.hist = hist;
      EventPair& ep = event_pair(d, ts[i].nev++);
      PSA_CUDA(cudaEventRecord(ep.a, d.stream));
      launch_point_lanczos(pp, ctas, pt_w, d.stream);
      PSA_CUDA(cudaEventRecord(ep.b, d.stream));
      PSA_CUDA(cudaGetLastError());
      launches += 1;
    }
    // the cancel flag (compute.jl:202-204) is forwarded to the running kernels through the mapped stop word
    unsigned idle_pt = 0;
    while (g.cancel && !cancelled) {
      bool busy = false;
      for (int i = 0; i < ng; ++i) {
        PSA_CUDA(cudaSetDevice(ctx->devs[i].dev));
        const cudaError_t qe = cudaStreamQuery(ctx->devs[i].stream);
        if (qe == cudaErrorNotReady) busy = true;
        else if (qe != cudaSuccess) PSA_CUDA(qe);
      }
      if (!busy) break;
      if (*g.cancel == 1) {
        for (int i = 0; i < ng; ++i) *(volatile unsigned long long*)&ctx->devs[i].host_state[2] = 1ull;
        cancelled = true;
      }
      short_pause(idle_pt);
    }
  }

  auto launch_sched = [&](int i, unsigned tick) {
    Device& d = ctx->devs[i];
    SchedParams sp;
    sp.W = d.W;
    sp.P = P[i];
    sp.rows_local = rows[i];
    sp.row0 = i;
    sp.row_step = ng;
    sp.x = d.x;
    sp.y = d.y;
    sp.order = order[i];
    sp.row_batch = row_batch[i];
    sp.slot_point = d.slot_point;
    sp.slot_z = d.slot_z;
    sp.slot_iter = d.slot_iter;
    sp.slot_beta = d.slot_beta;
    sp.slot_sigold = d.slot_sigold;
    sp.slot_sel = d.slot_sel;
    sp.slot_batch = d.slot_batch;
    sp.active = d.active;
    sp.fresh = d.fresh;
    sp.state = d.state;
    sp.prog = dyn_stride[i] ? d.prog : nullptr;
    sp.host_state = d.host_state_dev;
    sp.tick = tick;
    sched_kernel<<<1, 1024, 0, d.stream>>>(sp);
    launches += 1;
  };

  // estimated device time of a full tick decides between exact and pipelined launching (see TickState)
  std::vector<int> lag(ng, 0);
  for (int i = 0; i < ng; ++i) {
    if (ts[i].done) continue;
    const Device& d = ctx->devs[i];
    const double nn = (double)ctx->n * ctx->n, act = (double)std::min(d.W, std::max(P[i], 1));
    const double est_ms = is_hess_algo(ctx->algo) ? 16.0 * nn * act / 3.0e9 : 8.0 * nn * act / 2.5e10;
    lag[i] = est_ms < 3.0 ? 2 : 0;
    if (const char* e = getenv("PSA_TICK_LAG")) lag[i] = std::max(0, std::min(8, atoi(e)));  // experiments only
    ts[i].bound = std::min(d.W, (P[i] + APAD - 1) / APAD * APAD);
    PSA_CUDA(cudaSetDevice(d.dev));
    launch_sched(i, 1);
  }

  const int lanczos_maxit = g.maxit;
  auto launch_group = [&](int i, unsigned k) -> int {  // solve(k), lanczos(k), sched(k+1)
    Device& d = ctx->devs[i];
    TickState& t = ts[i];
    PSA_CUDA(cudaSetDevice(d.dev));
    const int bound = std::max(t.bound, 1);
    EventPair& ep = event_pair(d, t.nev++);
    if (ctx->algo == PSA_ALGO_SQ_LANC) {
      PSA_CUDA(cudaEventRecord(ep.a, d.stream));
      launch_solve_auto(solve_params(ctx, d), bound, d.num_sms, d.stream);
      PSA_CUDA(cudaEventRecord(ep.b, d.stream));
      launches += 1;
    } else {
      HessTick ht;
      ht.m = ctx->m;
      ht.n = ctx->n;
      ht.n_pad = ctx->n_pad;
      ht.H = d.Tcur;
      ht.R = d.R;
      ht.q0 = d.q0;
      ht.fresh = d.fresh;
      ht.n_fresh = (t.pub == k && t.pub_fresh >= 0) ? t.pub_fresh : bound;
      ht.n_fresh_dev = d.state + 2;
      ht.active = d.active;
      ht.n_active = bound;
      ht.n_active_dev = d.state + 1;
      ht.Q0 = d.Q0;
      ht.Q1 = d.Q1;
      ht.Wv = d.Wv;
      ht.sel = d.slot_sel;
      ht.iter = d.slot_iter;
      ht.batch = d.slot_batch;
      ht.z = d.slot_z;
      ht.full_sweep = ctx->algo == PSA_ALGO_RECT_QR;
      ht.num_sms = d.num_sms;
      PSA_CUDA(cudaEventRecord(ep.a, d.stream));
      launches += hess_launch_tick(ht, d.stream, ep.b);
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
    lp.slot_batch = d.slot_batch;
    lp.q0 = d.q0;
    lp.hist_alpha = d.hist_alpha;
    lp.hist_beta = d.hist_beta;
    lp.hist = d.hist;
    lp.tol = g.tol;
    lp.maxit = lanczos_maxit;
    lp.bigsigma = g.bigsigma;
    lp.Z = Zout[i];
    lp.iters = itout[i];
    lp.counts = d.counts;
    lp.prog = dyn_stride[i] ? d.prog : nullptr;
    launch_lanczos(lp, bound, d.stream);
    launches += 1;
    if (dyn_stride[i] && started[i] < P[i] && k % 4 == 0) {  // every 4th tick: the keys move one iteration per tick
      // points are still waiting: re-key them with the progress of the lattice points around them and re-sort
      RekeyParams rp{P[i], rows[i], g.lx, dyn_stride[i], d.prog, d.key_static, reinterpret_cast<unsigned*>(d.scratch64 + 3),
                     d.key_in, d.ord_in, d.state};
      rekey_kernel<<<(P[i] + 255) / 256, 256, 0, d.stream>>>(rp);
      size_t need = d.sort_tmp_bytes;
      PSA_CUDA(cub::DeviceRadixSort::SortPairsDescending(d.sort_tmp, need, d.key_in, d.key_out, d.ord_in, d.ord_out, P[i],
                                                         0, 32, d.stream));
      launches += 2;
    }
    launch_sched(i, k + 1);
    PSA_CUDA(cudaGetLastError());
    return PSA_OK;
  };

  int remaining = 0;
  for (int i = 0; i < ng; ++i) remaining += !ts[i].done;
  unsigned idle = 0;
  while (remaining > 0) {
    if (g.cancel && *g.cancel == 1) {
      cancelled = true;
      break;
    }
    bool progressed = false;
    for (int i = 0; i < ng; ++i) {
      TickState& t = ts[i];
      if (t.done) continue;
      Device& d = ctx->devs[i];
      const unsigned long long h0 = *(volatile unsigned long long*)&d.host_state[0];
      const unsigned long long h1 = *(volatile unsigned long long*)&d.host_state[1];
      const unsigned tick = (unsigned)(h0 >> 32);
      if (tick > t.pub) {
        t.pub = tick;
        t.pub_active = (int)(h0 & 0xffffffffu);
        t.pub_fresh = (unsigned)(h1 >> 32) == tick ? (int)(h1 & 0xffffffffu) : -1;
        // points started so far, as far as the host can tell (a stale value only costs a few extra re-sorts)
        if (t.pub_fresh >= 0) started[i] += t.pub_fresh;
        t.bound = std::min(t.bound, (t.pub_active + APAD - 1) / APAD * APAD);
        progressed = true;
        if (t.pub_active == 0) {  // nothing active and the queue is empty: this device is finished
          t.done = true;
          --remaining;
          continue;
        }
      }
      while (t.groups + 1 <= t.pub + (unsigned)lag[i]) {
        if ((rc = launch_group(i, t.groups + 1))) return rc;
        ++t.groups;
        progressed = true;
      }
    }
    if (progressed) idle = 0;
    else short_pause(idle);
  }
  // drain, collect counts + timing
  unsigned long long tot[PSA_NCOUNTS] = {0, 0, 0, 0};
  double max_ms = 0;
  for (int i = 0; i < ng; ++i) {
    Device& d = ctx->devs[i];
    PSA_CUDA(cudaSetDevice(d.dev));
    PSA_CUDA(cudaEventRecord(d.g1, d.stream));
    PSA_CUDA(cudaStreamSynchronize(d.stream));
    unsigned long long c[PSA_NCOUNTS];
    int st[4];
    PSA_CUDA(cudaMemcpy(c, d.counts, sizeof c, cudaMemcpyDeviceToHost));
    PSA_CUDA(cudaMemcpy(st, d.state, sizeof st, cudaMemcpyDeviceToHost));
    for (int k = 0; k < PSA_NCOUNTS; ++k) tot[k] += c[k];
    float gms = 0;
    cudaEventElapsedTime(&gms, d.g0, d.g1);
    max_ms = std::max(max_ms, (double)gms);
    if (i == 0) {
      const size_t live = (ctx->algo == PSA_ALGO_SVD || point_mode) ? ts[i].nev : std::min<size_t>(ts[i].nev, (size_t)st[3]);
      double solve_ms = 0;
      for (size_t e = 0; e < live; ++e) {  // groups launched past the last live tick are empty: not counted
        float ms = 0;
        cudaEventElapsedTime(&ms, d.ev[e].a, d.ev[e].b);
        solve_ms += ms;
      }
      ctx->stats[PSA_STAT_SOLVE_MS] = solve_ms;
      ctx->stats[PSA_STAT_GRID_MS] = gms;
      ctx->stats[PSA_STAT_SOLVE_LAUNCHES] = (double)live;
      ctx->stats[PSA_STAT_TICKS] = ctx->algo == PSA_ALGO_SVD ? 0.0 : (double)st[3];
    }
  }
  ctx->stats[PSA_STAT_MAX_DEVICE_MS] = max_ms;
  ctx->stats[PSA_STAT_TOTAL_LAUNCHES] = launches;
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

// Largest eigenpair of the l x l symmetric tridiagonal (diag a, off-diagonal b) on the host: implicit-shift QL with
// accumulated rotations (the textbook tql2 recurrence).  Stands in for `eigen(H)` at modes.jl:255-257; l <= num_its.
bool tridiag_top_eigvec(int l, const double* a, const double* b, double* lambda, std::vector<double>& vec) {
  std::vector<double> d(a, a + l), e(l, 0.0), z((size_t)l * l, 0.0);
  for (int i = 0; i + 1 < l; ++i) e[i] = b[i];
  for (int i = 0; i < l; ++i) z[(size_t)i * l + i] = 1.0;  // z[k*l + i]: component k of eigenvector i (row-major by k)
  for (int j = 0; j < l; ++j) {
    int iter = 0, mm;
    do {
      for (mm = j; mm + 1 < l; ++mm) {
        const double dd = std::fabs(d[mm]) + std::fabs(d[mm + 1]);
        if (std::fabs(e[mm]) <= std::numeric_limits<double>::epsilon() * dd) break;
      }
      if (mm != j) {
        if (++iter > 60) return false;
        double gg = (d[j + 1] - d[j]) / (2.0 * e[j]);
        double r = std::hypot(gg, 1.0);
        gg = d[mm] - d[j] + e[j] / (gg + (gg >= 0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = mm - 1; i >= j; --i) {
          double f = s * e[i], bb = c * e[i];
          r = std::hypot(f, gg);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[mm] = 0.0;
            break;
          }
          s = f / r;
          c = gg / r;
          gg = d[i + 1] - p;
          r = (d[i] - gg) * s + 2.0 * c * bb;
          p = s * r;
          d[i + 1] = gg + p;
          gg = c * r - bb;
          for (int k = 0; k < l; ++k) {
            f = z[(size_t)k * l + i + 1];
            z[(size_t)k * l + i + 1] = s * z[(size_t)k * l + i] + c * f;
            z[(size_t)k * l + i] = c * z[(size_t)k * l + i] - s * f;
          }
        }
        if (r == 0.0 && i >= j) continue;
        d[j] -= p;
        e[j] = gg;
        e[mm] = 0.0;
      }
    } while (mm != j);
  }
  int best = 0;
  for (int i = 1; i < l; ++i)
    if (d[i] > d[best]) best = i;
  *lambda = d[best];
  vec.resize(l);
  for (int k = 0; k < l; ++k) vec[k] = z[(size_t)k * l + best];
  return true;
}

__global__ void pseudo_init_kernel(int W, double zre, double zim, int* slot_point, double2* slot_z, int* slot_iter,
                                   double* slot_beta, double* slot_sigold, int* slot_sel, int* slot_batch, int* active,
                                   int* state) {
  const int t = threadIdx.x;
  if (t == 0) {
    slot_point[0] = 0;
    slot_z[0] = make_double2(zre, zim);
    slot_iter[0] = 0;
    slot_beta[0] = 0.0;
    slot_sigold[0] = 0.0;
    slot_sel[0] = 0;
    slot_batch[0] = 0;
    state[0] = 1;
    state[1] = 1;
    state[2] = 1;
  }
  if (t < APAD) active[t] = t == 0 ? 0 : W;
}

// upload nbatch start vectors (n complex each, host or device) into the zero-padded [nbatch][n_pad] store
int upload_q0(Context* ctx, Device& d, const double* q0s, int nbatch, bool on_device) {
  int rc;
  const size_t n = ctx->n, n_pad = ctx->n_pad;
  if ((rc = ensure_cap(ctx, &d.q0, &d.q0_cap, (size_t)nbatch * n_pad))) return rc;
  PSA_CUDA(cudaMemsetAsync(d.q0, 0, (size_t)nbatch * n_pad * 16, d.stream));
  PSA_CUDA(cudaMemcpy2DAsync(d.q0, n_pad * 16, q0s, n * 16, n * 16, nbatch,
                             on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, d.stream));
  return PSA_OK;
}

int grid_host(Context* ctx, const double* x, int lx, const double* y, int ly, const double* q0s, int nbatch,
              const int32_t* row_batch, double tol, int maxit, double bigsigma, double* Z, int32_t* iters,
              int64_t* counts, int* algo, volatile int* cancel) {
  if (ctx->algo == PSA_ALGO_NONE) return PSA_ERR_NOMATRIX;
  if (!x || !y || !q0s || lx < 1 || ly < 1 || maxit < 1 || maxit > HIST_MAX - 2 || nbatch < 1) return PSA_ERR_ARG;
  if (nbatch > 1 && !row_batch) return PSA_ERR_ARG;
  if (row_batch)
    for (int j = 0; j < ly; ++j)
      if (row_batch[j] < 0 || row_batch[j] >= nbatch) return PSA_ERR_ARG;
  if (algo) *algo = ctx->algo;
  ctx->last_Z = nullptr;
  const int ng = (int)ctx->devs.size();
  std::vector<double*> Zout(ng);
  std::vector<int*> itout(ng);
  std::vector<const int*> rb(ng, nullptr);
  int rc;
  for (int i = 0; i < ng; ++i) {
    Device& d = ctx->devs[i];
    PSA_CUDA(cudaSetDevice(d.dev));
    const int rows = (ly - i + ng - 1) / ng;
    const size_t P = (size_t)rows * lx;
    if ((rc = ensure_cap(ctx, &d.x, &d.x_cap, (size_t)lx))) return rc;
    if ((rc = ensure_cap(ctx, &d.y, &d.y_cap, (size_t)ly))) return rc;
    if ((rc = ensure_cap(ctx, &d.Z, &d.Z_cap, P))) return rc;
    if ((rc = ensure_cap(ctx, &d.iters, &d.it_cap, P))) return rc;
    PSA_CUDA(cudaMemcpyAsync(d.x, x, lx * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    PSA_CUDA(cudaMemcpyAsync(d.y, y, ly * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    if ((rc = upload_q0(ctx, d, q0s, nbatch, false))) return rc;
    if (row_batch) {
      if ((rc = ensure_cap(ctx, &d.row_batch, &d.rb_cap, (size_t)ly))) return rc;
      PSA_CUDA(cudaMemcpyAsync(d.row_batch, row_batch, ly * sizeof(int), cudaMemcpyHostToDevice, d.stream));
      rb[i] = d.row_batch;
    }
    Zout[i] = d.Z;
    itout[i] = d.iters;
  }
  GridArgs g{lx, ly, tol, maxit, bigsigma, cancel};
  rc = run_grid(ctx, g, rb, Zout, itout, counts);
  if (rc != PSA_OK) return rc;
  // gather: device i owns rows i, i+ng, ... ; local layout is rows_i x lx column-major
  Device& d0 = ctx->devs[0];
  const size_t Ptot = (size_t)ly * lx;
  PSA_CUDA(cudaSetDevice(d0.dev));
  const double* Zdev = d0.Z;
  const int* itdev = d0.iters;
  if (ng > 1) {
    // the sigma_min tiles travel to device 0 over NVLink (grouped ncclSend/ncclRecv) and are interleaved there
    if ((rc = ensure_cap(ctx, &d0.Zfull, &d0.Zfull_cap, Ptot))) return rc;
    if (iters && (rc = ensure_cap(ctx, &d0.itfull, &d0.itfull_cap, Ptot))) return rc;
    for (int i = 1; i < ng; ++i) {
      const size_t P = (size_t)((ly - i + ng - 1) / ng) * lx;
      if (ctx->gather_cap[i] < P) {
        ctx->gather_cap[i] = 0;
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
    PSA_CUDA(cudaSetDevice(d0.dev));
    for (int i = 0; i < ng; ++i) {
      const int rows = (ly - i + ng - 1) / ng;
      const size_t P = (size_t)rows * lx;
      if (P == 0) continue;
      const int grid = (int)std::min<size_t>((P + 255) / 256, 4096);
      interleave_rows_kernel<double><<<grid, 256, 0, d0.stream>>>(i == 0 ? d0.Z : ctx->gather_buf[i], rows, lx, i, ng, ly, d0.Zfull);
      if (iters)
        interleave_rows_kernel<int><<<grid, 256, 0, d0.stream>>>(i == 0 ? d0.iters : ctx->gather_ibuf[i], rows, lx, i, ng, ly, d0.itfull);
    }
    PSA_CUDA(cudaGetLastError());
    for (int i = 1; i < ng; ++i) {
      PSA_CUDA(cudaSetDevice(ctx->devs[i].dev));
      PSA_CUDA(cudaStreamSynchronize(ctx->devs[i].stream));
    }
    PSA_CUDA(cudaSetDevice(d0.dev));
    Zdev = d0.Zfull;
    itdev = d0.itfull;
  }
  if (Z) PSA_CUDA(cudaMemcpyAsync(Z, Zdev, Ptot * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
  if (iters) PSA_CUDA(cudaMemcpyAsync(iters, itdev, Ptot * sizeof(int), cudaMemcpyDeviceToHost, d0.stream));
  PSA_CUDA(cudaStreamSynchronize(d0.stream));
  ctx->last_Z = Zdev;
  ctx->last_lx = lx;
  ctx->last_ly = ly;
  return PSA_OK;
}

}  // namespace

extern "C" {

const char* psa_gpu_version(void) { return "psagpu 0.2.0 sm_100a"; }

int psa_gpu_plan(int m, int n, int maxit, int* warps_per_sm) {
  if (warps_per_sm) *warps_per_sm = 0;
  if (m < 1 || n < 1 || maxit < 1 || maxit > HIST_MAX - 2) return PSA_ALGO_NONE;
  Context defaults;  // the reference's default thresholds
  const int algo = classify(&defaults, m, n);
  if (algo == PSA_ALGO_SQ_LANC && warps_per_sm && !getenv("PSA_NO_POINT_KERNEL"))
    *warps_per_sm = pt_warps(n, std::max(HIST_MIN, (maxit + 2 + 31) / 32 * 32));
  return algo;
}

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
    ok = ok && cudaMalloc((void**)&d.state, 8 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&d.counts, PSA_NCOUNTS * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&d.scratch64, 4 * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaHostAlloc((void**)&d.host_state, 4 * sizeof(unsigned long long), cudaHostAllocMapped) == cudaSuccess;
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

int psa_gpu_set_thresholds(void* vctx, int minlancs4psa, int maxstdqr4hess) {
  if (!vctx || minlancs4psa < 1 || maxstdqr4hess < 1) return PSA_ERR_ARG;
  Context* ctx = static_cast<Context*>(vctx);
  ctx->minlancs = minlancs4psa;
  ctx->maxstdqr = maxstdqr4hess;
  return PSA_OK;
}

static int set_matrix_common(Context* ctx, const double* T, int64_t ldt, int m, int n, bool on_device, int flags) {
  if (!T || m < n || n < 1 || ldt < m) return PSA_ERR_ARG;
  const int algo = classify(ctx, m, n);
  if (algo == PSA_ALGO_NONE) return PSA_ERR_UNSUPPORTED;
  ctx->algo = ctx->algo_full = PSA_ALGO_NONE;
  ctx->last_Z = nullptr;
  // raw copy to device 0
  Device& d0 = ctx->devs[0];
  PSA_CUDA(cudaSetDevice(d0.dev));
  int rc;
  if ((rc = ensure_cap(ctx, &d0.T, &d0.T_cap, (size_t)m * n))) return rc;
  PSA_CUDA(cudaMemcpy2DAsync(d0.T, (size_t)m * 16, T, (size_t)ldt * 16, (size_t)m * 16, n,
                             on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, d0.stream));
  // structure: square input must be upper triangular (compute.jl:597-601) unless the caller passed the parent of an
  // UpperTriangular wrapper; rectangular input must be upper Hessenberg (compute.jl:579-590).  Checked on the
  // device copy: O(n^2) reads, once per matrix.
  const int band = (algo == PSA_ALGO_SQ_LANC && !(flags & PSA_MATRIX_UPPER_WRAPPER)) ? 1 : (is_hess_algo(algo) ? 2 : 0);
  if (band) {
    PSA_CUDA(cudaMemsetAsync(d0.scratch64, 0, sizeof(unsigned long long), d0.stream));
    below_band_count_kernel<<<std::min(n, 8 * d0.num_sms), 256, 0, d0.stream>>>(d0.T, m, m, n, band, d0.scratch64);
    unsigned long long nnz = 0;
    PSA_CUDA(cudaMemcpyAsync(&nnz, d0.scratch64, sizeof nnz, cudaMemcpyDeviceToHost, d0.stream));
    PSA_CUDA(cudaStreamSynchronize(d0.stream));
    if (nnz != 0) {
      ctx->err = band == 1 ? "square input is not upper triangular (the reference factorises it: host path)"
                           : "rectangular input is not upper Hessenberg (host path)";
      return PSA_ERR_UNSUPPORTED;
    }
  } else {
    PSA_CUDA(cudaStreamSynchronize(d0.stream));
  }
  ctx->m = m;
  ctx->n = n;
  ctx->n_pad = (n + MB - 1) / MB * MB;
  ctx->nblk = ctx->n_pad / MB;
  if (ctx->devs.size() > 1) {
    // replicate T with one NCCL broadcast over NVLink (root = device 0)
    if ((rc = ensure_comms(ctx))) return rc;
    for (size_t i = 1; i < ctx->devs.size(); ++i) {
      Device& d = ctx->devs[i];
      PSA_CUDA(cudaSetDevice(d.dev));
      if ((rc = ensure_cap(ctx, &d.T, &d.T_cap, (size_t)m * n))) return rc;
    }
    PSA_NCCL(ctx->nccl.GroupStart());
    for (size_t i = 0; i < ctx->devs.size(); ++i) {
      Device& d = ctx->devs[i];
      PSA_NCCL(ctx->nccl.Broadcast(d0.T, d.T, (size_t)m * n * 2, ncclDouble, 0, ctx->comms[i], d.stream));
    }
    PSA_NCCL(ctx->nccl.GroupEnd());
  }
  ctx->algo = algo;
  ctx->m_full = m;
  ctx->n_full = n;
  ctx->algo_full = algo;
  for (auto& d : ctx->devs) {
    d.Tcur = d.T;
    if ((rc = prepare_matrix(ctx, d))) {
      ctx->algo = ctx->algo_full = PSA_ALGO_NONE;
      return rc;
    }
  }
  return PSA_OK;
}

// make (m, n, raw pointer) the current matrix on every device and re-pack
static int activate_matrix(Context* ctx, int m, int n, bool projected) {
  const int algo = classify(ctx, m, n);
  if (algo == PSA_ALGO_NONE) return PSA_ERR_UNSUPPORTED;
  ctx->m = m;
  ctx->n = n;
  ctx->n_pad = (n + MB - 1) / MB * MB;
  ctx->nblk = ctx->n_pad / MB;
  ctx->algo = algo;
  ctx->last_Z = nullptr;
  for (auto& d : ctx->devs) {
    d.Tcur = projected ? d.Tp : d.T;
    int rc = prepare_matrix(ctx, d);
    if (rc) {
      ctx->algo = PSA_ALGO_NONE;
      return rc;
    }
  }
  return PSA_OK;
}

int psa_gpu_set_matrix(void* vctx, const double* T, int64_t ldt, int m, int n) {
  if (!vctx) return PSA_ERR_ARG;
  return set_matrix_common(static_cast<Context*>(vctx), T, ldt, m, n, false, PSA_MATRIX_DEFAULT);
}

int psa_gpu_set_matrix_ex(void* vctx, const double* T, int64_t ldt, int m, int n, int flags) {
  if (!vctx) return PSA_ERR_ARG;
  return set_matrix_common(static_cast<Context*>(vctx), T, ldt, m, n, false, flags);
}

int psa_gpu_set_matrix_device(void* vctx, const double* d_T, int64_t ldt, int m, int n) {
  if (!vctx) return PSA_ERR_ARG;
  return set_matrix_common(static_cast<Context*>(vctx), d_T, ldt, m, n, true, PSA_MATRIX_DEFAULT);
}

int psa_gpu_project(void* vctx, const int32_t* selection, int npick, double* Tproj, int64_t ldp) {
  if (!vctx) return PSA_ERR_ARG;
  Context* ctx = static_cast<Context*>(vctx);
  if (ctx->algo_full == PSA_ALGO_NONE) return PSA_ERR_NOMATRIX;
  if (ctx->algo_full != PSA_ALGO_SQ_LANC) return PSA_ERR_UNSUPPORTED;  // the reference projects square dense input only
  const int n = ctx->n_full;
  if (npick < 1 || npick > n || (npick < n && !selection) || (Tproj && ldp < npick)) return PSA_ERR_ARG;
  if (npick == n) {  // every eigenvalue selected: the full matrix is (again) the current one (compute.jl:322-328)
    if (ctx->n == n && ctx->devs[0].Tcur == ctx->devs[0].T) return PSA_OK;
    return activate_matrix(ctx, ctx->m_full, n, false);
  }
  for (int i = 0; i < npick; ++i)
    if (selection[i] < 0 || selection[i] >= n || (i > 0 && selection[i] <= selection[i - 1])) return PSA_ERR_ARG;
  if (classify(ctx, npick, npick) == PSA_ALGO_NONE) return PSA_ERR_UNSUPPORTED;
  Device& d0 = ctx->devs[0];
  PSA_CUDA(cudaSetDevice(d0.dev));
  int rc;
  if ((rc = ensure_cap(ctx, &d0.Twork, &d0.Twork_cap, (size_t)n * n))) return rc;
  if (d0.proj_cap < (size_t)npick) {
    d0.proj_cap = 0;
    if ((rc = dev_alloc(ctx, &d0.proj_sel, (size_t)npick))) return rc;
    if ((rc = dev_alloc(ctx, &d0.proj_start, (size_t)npick))) return rc;
    if ((rc = dev_alloc(ctx, &d0.proj_rot, (size_t)npick))) return rc;
    d0.proj_cap = (size_t)npick;
  }
  // wavefront schedule: bubble i starts its swaps at step start[i], two positions behind bubble i-1
  std::vector<int> start(npick, 0);
  int steps = 0;
  for (int i = 0; i < npick; ++i) {
    if (i > 0) start[i] = start[i - 1] + std::max(0, 2 - (selection[i] - selection[i - 1]));
    const int len = selection[i] - i;
    if (len > 0) steps = std::max(steps, start[i] + len);
  }
  PSA_CUDA(cudaMemcpyAsync(d0.proj_sel, selection, npick * sizeof(int), cudaMemcpyHostToDevice, d0.stream));
  PSA_CUDA(cudaMemcpyAsync(d0.proj_start, start.data(), npick * sizeof(int), cudaMemcpyHostToDevice, d0.stream));
  proj_copy_triu_kernel<<<n, 256, 0, d0.stream>>>(d0.T, ctx->m_full, n, d0.Twork, n);
  const int gy = std::max(1, std::min(8, (n + 1023) / 1024));
  for (int t = 0; t < steps; ++t) {
    proj_params_kernel<<<(npick + 127) / 128, 128, 0, d0.stream>>>(t, npick, d0.proj_sel, d0.proj_start, d0.Twork, n, d0.proj_rot);
    proj_cols_kernel<<<dim3(npick, gy), 256, 0, d0.stream>>>(npick, d0.proj_rot, d0.Twork, n);
    proj_rows_kernel<<<dim3(npick, gy), 256, 0, d0.stream>>>(npick, n, d0.proj_rot, d0.Twork, n);
  }
  if ((rc = ensure_cap(ctx, &d0.Tp, &d0.Tp_cap, (size_t)npick * npick))) return rc;
  proj_extract_kernel<<<npick, 256, 0, d0.stream>>>(d0.Twork, n, npick, d0.Tp);
  PSA_CUDA(cudaGetLastError());
  if (Tproj)
    PSA_CUDA(cudaMemcpy2DAsync(Tproj, (size_t)ldp * 16, d0.Tp, (size_t)npick * 16, (size_t)npick * 16, npick,
                               cudaMemcpyDeviceToHost, d0.stream));
  PSA_CUDA(cudaStreamSynchronize(d0.stream));
  if (ctx->devs.size() > 1) {
    if ((rc = ensure_comms(ctx))) return rc;
    for (size_t i = 1; i < ctx->devs.size(); ++i) {
      Device& d = ctx->devs[i];
      PSA_CUDA(cudaSetDevice(d.dev));
      if ((rc = ensure_cap(ctx, &d.Tp, &d.Tp_cap, (size_t)npick * npick))) return rc;
    }
    PSA_NCCL(ctx->nccl.GroupStart());
    for (size_t i = 0; i < ctx->devs.size(); ++i) {
      Device& d = ctx->devs[i];
      PSA_NCCL(ctx->nccl.Broadcast(d0.Tp, d.Tp, (size_t)npick * npick * 2, ncclDouble, 0, ctx->comms[i], d.stream));
    }
    PSA_NCCL(ctx->nccl.GroupEnd());
  }
  return activate_matrix(ctx, npick, npick, true);
}

int psa_gpu_grid(void* vctx, const double* x, int lx, const double* y, int ly, const double* q0, double tol, int maxit,
                 double bigsigma, double* Z, int32_t* iters, int64_t* counts, int* algo, volatile int* cancel) {
  if (!vctx) return PSA_ERR_ARG;
  return grid_host(static_cast<Context*>(vctx), x, lx, y, ly, q0, 1, nullptr, tol, maxit, bigsigma, Z, iters, counts,
                   algo, cancel);
}

int psa_gpu_grid_batched(void* vctx, const double* x, int lx, const double* y, int ly, const double* q0s, int nbatch,
                         const int32_t* row_batch, double tol, int maxit, double bigsigma, double* Z, int32_t* iters,
                         int64_t* counts, int* algo, volatile int* cancel) {
  if (!vctx) return PSA_ERR_ARG;
  return grid_host(static_cast<Context*>(vctx), x, lx, y, ly, q0s, nbatch, row_batch, tol, maxit, bigsigma, Z, iters,
                   counts, algo, cancel);
}

int psa_gpu_grid_device(void* vctx, const double* d_x, int lx, const double* d_y, int ly, const double* d_q0, double tol,
                        int maxit, double bigsigma, double* d_Z, int32_t* d_iters, int64_t* counts, int* algo,
                        volatile int* cancel) {
  if (!vctx) return PSA_ERR_ARG;
  Context* ctx = static_cast<Context*>(vctx);
  if (ctx->algo == PSA_ALGO_NONE) return PSA_ERR_NOMATRIX;
  if (ctx->devs.size() != 1) return PSA_ERR_UNSUPPORTED;
  if (!d_x || !d_y || !d_q0 || !d_Z || lx < 1 || ly < 1 || maxit < 1 || maxit > HIST_MAX - 2) return PSA_ERR_ARG;
  if (algo) *algo = ctx->algo;
  ctx->last_Z = nullptr;
  Device& d = ctx->devs[0];
  PSA_CUDA(cudaSetDevice(d.dev));
  int rc = upload_q0(ctx, d, d_q0, 1, true);  // into the zero-padded store the kernels read in place
  if (rc) return rc;
  // borrow the caller's buffers for this call
  double *sx = d.x, *sy = d.y;
  d.x = const_cast<double*>(d_x);
  d.y = const_cast<double*>(d_y);
  std::vector<double*> Zout{d_Z};
  std::vector<int*> itout{d_iters};
  std::vector<const int*> rb{nullptr};
  GridArgs g{lx, ly, tol, maxit, bigsigma, cancel};
  rc = run_grid(ctx, g, rb, Zout, itout, counts);
  d.x = sx;
  d.y = sy;
  if (rc == PSA_OK) {
    ctx->last_Z = d_Z;
    ctx->last_lx = lx;
    ctx->last_ly = ly;
  }
  return rc;
}

int psa_gpu_postprocess(void* vctx, const double* y, int ly, int lx, int n_mirror, double ps_tiny, double small_sigma,
                        double* Zfull, int64_t ldz, double* yfull, double* zstats) {
  if (!vctx) return PSA_ERR_ARG;
  Context* ctx = static_cast<Context*>(vctx);
  if (!ctx->last_Z) return PSA_ERR_NOMATRIX;
  if (!y || !Zfull || lx != ctx->last_lx || ly != ctx->last_ly) return PSA_ERR_ARG;
  const int nm = n_mirror < 0 ? -n_mirror : n_mirror;
  const int y0_is_zero = (n_mirror > 0 && y[0] == 0.0) ? 1 : 0;
  if (nm > ly || (y0_is_zero && nm + 1 > ly)) return PSA_ERR_ARG;
  const int lyf = ly + nm;
  if (ldz < lyf) return PSA_ERR_ARG;
  Device& d = ctx->devs[0];
  PSA_CUDA(cudaSetDevice(d.dev));
  int rc;
  if ((rc = ensure_cap(ctx, &d.post, &d.post_cap, (size_t)lyf * lx))) return rc;
  const unsigned long long init[3] = {~0ull, 0ull, ~0ull};
  PSA_CUDA(cudaMemcpyAsync(d.scratch64, init, sizeof init, cudaMemcpyHostToDevice, d.stream));
  PostParams pp;
  pp.Z = ctx->last_Z;
  pp.ly = ly;
  pp.lx = lx;
  pp.n_mirror = n_mirror;
  pp.y0_is_zero = y0_is_zero;
  pp.ps_tiny = ps_tiny;
  pp.small_sigma = small_sigma;
  pp.out = d.post;
  pp.lyf = lyf;
  pp.stats = d.scratch64;
  const int grid = (int)std::min<size_t>(((size_t)lyf * lx + 255) / 256, (size_t)8 * d.num_sms);
  postprocess_kernel<<<grid, 256, 0, d.stream>>>(pp);
  PSA_CUDA(cudaGetLastError());
  unsigned long long bits[3];
  PSA_CUDA(cudaMemcpyAsync(bits, d.scratch64, sizeof bits, cudaMemcpyDeviceToHost, d.stream));
  PSA_CUDA(cudaMemcpy2DAsync(Zfull, (size_t)ldz * 8, d.post, (size_t)lyf * 8, (size_t)lyf * 8, lx, cudaMemcpyDeviceToHost,
                             d.stream));
  PSA_CUDA(cudaStreamSynchronize(d.stream));
  if (zstats) {
    const double inf = std::numeric_limits<double>::infinity();
    double v[3];
    for (int k = 0; k < 3; ++k) memcpy(&v[k], &bits[k], 8);
    zstats[0] = bits[0] == ~0ull ? inf : v[0];
    zstats[1] = v[1];
    zstats[2] = bits[2] == ~0ull ? inf : v[2];
  }
  if (yfull) {  // compute.jl:222-235
    if (n_mirror < 0) {
      for (int r = 0; r < ly; ++r) yfull[r] = y[r];
      for (int t = 0; t < nm; ++t) yfull[ly + t] = -y[ly - 1 - t];
    } else {
      for (int t = 0; t < nm; ++t) yfull[t] = -y[y0_is_zero ? nm - t : nm - 1 - t];
      for (int r = 0; r < ly; ++r) yfull[nm + r] = y[r];
    }
  }
  return PSA_OK;
}

int psa_gpu_pseudomode(void* vctx, double zre, double zim, const double* q0, double tol, int num_its, double* sigmin,
                       double* mode, int* iters_out, int* info) {
  if (!vctx) return PSA_ERR_ARG;
  Context* ctx = static_cast<Context*>(vctx);
  if (ctx->algo == PSA_ALGO_NONE) return PSA_ERR_NOMATRIX;
  if (ctx->algo != PSA_ALGO_SQ_LANC || ctx->n < 200) return PSA_ERR_UNSUPPORTED;  // modes.jl:193: plain SVD below 200
  if (!q0 || !sigmin || !mode || num_its < 1 || num_its > HIST_MAX - 2) return PSA_ERR_ARG;
  Device& d = ctx->devs[0];
  PSA_CUDA(cudaSetDevice(d.dev));
  int rc = set_solve_attr(ctx);
  if (rc) return rc;
  const int hist = std::max(HIST_MIN, (num_its + 2 + 31) / 32 * 32);
  if ((rc = ensure_slots(ctx, d, std::max(d.W, APAD), hist))) return rc;
  const size_t n = ctx->n;
  if ((rc = ensure_cap(ctx, &d.basis, &d.basis_cap, (size_t)num_its * n))) return rc;
  if ((rc = ensure_cap(ctx, &d.Z, &d.Z_cap, (size_t)1))) return rc;
  if ((rc = ensure_cap(ctx, &d.iters, &d.it_cap, (size_t)1))) return rc;
  if ((rc = upload_q0(ctx, d, q0, 1, false))) return rc;
  ctx->last_Z = nullptr;
  PSA_CUDA(cudaMemsetAsync(d.counts, 0, PSA_NCOUNTS * sizeof(unsigned long long), d.stream));
  pseudo_init_kernel<<<1, APAD, 0, d.stream>>>(d.W, zre, zim, d.slot_point, d.slot_z, d.slot_iter, d.slot_beta,
                                               d.slot_sigold, d.slot_sel, d.slot_batch, d.active, d.state);
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
  lp.slot_batch = d.slot_batch;
  lp.q0 = d.q0;
  lp.hist_alpha = d.hist_alpha;
  lp.hist_beta = d.hist_beta;
  lp.hist = d.hist;
  lp.tol = tol;
  lp.maxit = num_its;
  lp.bigsigma = 1e308;  // modes.jl:243
  lp.Z = d.Z;
  lp.iters = d.iters;
  lp.counts = d.counts;
  lp.prog = nullptr;
  const SolveParams sp = solve_params(ctx, d);
  int l = 0, point = 0;
  const int cgrid = (int)std::min<size_t>((n + 255) / 256, 4 * (size_t)d.num_sms);
  for (int it = 1; it <= num_its && point >= 0; ++it) {
    // Q[:, it] = q  (modes.jl:236; the start vector while the slot is fresh, then the buffer selected by the parity)
    const double2* qcur = it == 1 ? d.q0 : ((it - 1) & 1 ? d.Q1 : d.Q0);
    basis_store_kernel<<<cgrid, 256, 0, d.stream>>>((int)n, qcur, d.basis + (size_t)(it - 1) * n);
    launch_solve_auto(sp, 1, d.num_sms, d.stream);
    launch_lanczos(lp, 1, d.stream);
    PSA_CUDA(cudaGetLastError());
    PSA_CUDA(cudaMemcpyAsync(&point, d.slot_point, sizeof(int), cudaMemcpyDeviceToHost, d.stream));
    PSA_CUDA(cudaStreamSynchronize(d.stream));
    l = it;
  }
  std::vector<double> ha(l), hb(l);
  double Zv = 0;
  unsigned long long c[PSA_NCOUNTS];
  PSA_CUDA(cudaMemcpy(ha.data(), d.hist_alpha, l * sizeof(double), cudaMemcpyDeviceToHost));
  PSA_CUDA(cudaMemcpy(hb.data(), d.hist_beta, l * sizeof(double), cudaMemcpyDeviceToHost));
  PSA_CUDA(cudaMemcpy(&Zv, d.Z, sizeof(double), cudaMemcpyDeviceToHost));
  PSA_CUDA(cudaMemcpy(c, d.counts, sizeof c, cudaMemcpyDeviceToHost));
  double lam = 0;
  std::vector<double> yv;
  const bool eig_ok = c[PSA_COUNT_EIGFAIL] == 0 && tridiag_top_eigvec(l, ha.data(), hb.data(), &lam, yv);
  if (!eig_ok) {  // `catch; l = num_its` (modes.jl:258-260)
    l = num_its;
    yv.assign(std::max(l, 1), 0.0);
  }
  if ((rc = ensure_cap(ctx, &d.yvec, &d.yvec_cap, yv.size()))) return rc;
  PSA_CUDA(cudaMemcpyAsync(d.yvec, yv.data(), yv.size() * sizeof(double), cudaMemcpyHostToDevice, d.stream));
  basis_combine_kernel<<<(int)((n + 255) / 256), 256, 0, d.stream>>>((int)n, eig_ok ? l : 0, d.basis, d.yvec, d.Wv);
  PSA_CUDA(cudaGetLastError());
  PSA_CUDA(cudaMemcpyAsync(mode, d.Wv, n * 16, cudaMemcpyDeviceToHost, d.stream));
  PSA_CUDA(cudaStreamSynchronize(d.stream));
  *sigmin = Zv;
  if (iters_out) *iters_out = l;
  if (info) *info = (l >= num_its || !eig_ok) ? 1 : 0;  // modes.jl:262: `if l >= num_its` -> NaN / dense-SVD fallback
  return PSA_OK;
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
  if ((rc = ensure_slots(ctx, d, std::max(W, d.W), HIST_MIN))) return rc;
  if ((rc = ensure_cap(ctx, &d.q0, &d.q0_cap, (size_t)ctx->n_pad))) return rc;  // never read here, but must exist
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
  load_slots_kernel<<<W, 256, 0, d.stream>>>(ns, ctx->n, ctx->n_pad, dQ, dz, d.Q0, d.slot_z, d.slot_sel, d.slot_iter,
                                             d.active, d.W, d.state);
  EventPair& ep = event_pair(d, 0);
  PSA_CUDA(cudaEventRecord(ep.a, d.stream));
  launch_solve_auto(solve_params(ctx, d), ns, d.num_sms, d.stream);
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
