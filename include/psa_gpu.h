/* psa_gpu.h -- C ABI of libpsagpu.so: B200-native resolvent-norm grid kernel for Pseudospectra.jl.
 *
 * This is the drop-in boundary for ONE call of the reference: the per-batch core
 *     psacore(T, S, q0, x, y, bw; tol, ...) -> (Z, algo, warned)        /root/reference/src/compute.jl:448-617
 * as invoked from the dense row loop of psa_compute (src/compute.jl:199-219).  A Julia `ccall` shim
 * (julia/PseudospectraGPU.jl, shown in INTEGRATION.md) replaces that loop when opts[:gpu] is set and keeps
 * everything before (options, grid, projection: compute.jl:94-153) and after (mirror, clamp, levels:
 * compute.jl:222-267) unchanged.
 *
 * Conventions: plain pointers and sizes only; all matrices column-major like Julia's; complex numbers are
 * interleaved (re, im) doubles (`ComplexF64` / `double _Complex`).  Every function returns an int status
 * (0 = ok, < 0 = error, see psa_gpu_strerror); no C++ exception crosses this boundary.  Host pointers are
 * never retained after a call returns.  A context is not re-entrant.  There is NO CPU fallback: every entry
 * point fails with PSA_ERR_CUDA / PSA_ERR_NODEVICE when no sm_100 device is usable.
 */
#ifndef PSA_GPU_H
#define PSA_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define PSA_OK 0
#define PSA_ERR_ARG (-1)         /* bad argument (null pointer, m < n, sizes <= 0, ...) */
#define PSA_ERR_CUDA (-2)        /* CUDA runtime / launch failure; text via psa_gpu_last_error */
#define PSA_ERR_CANCELLED (-3)   /* *cancel became 1: mirrors `info = -3` / `return nothing` (compute.jl:83,202-204) */
#define PSA_ERR_UNSUPPORTED (-4) /* shape the GPU path does not cover: the shim falls through to the Julia code */
#define PSA_ERR_NODEVICE (-5)    /* no CUDA device / not sm_100 */
#define PSA_ERR_NOMATRIX (-6)    /* psa_gpu_grid before psa_gpu_set_matrix */
#define PSA_ERR_NCCL (-7)        /* multi-GPU transport failure */
#define PSA_ERR_NOMEM (-8)       /* device allocation failed */

/* algorithm actually used, mirrors the `algo` Symbol psacore returns (compute.jl:535,580,594) */
#define PSA_ALGO_NONE 0
#define PSA_ALGO_SQ_LANC 1 /* :sq_lanc -- square upper-triangular T, m == n >= minlancs4psa (55)     */
#define PSA_ALGO_HESSQR 2  /* :HessQR  -- rectangular Hessenberg, bw == 2 and m > maxstdqr4hess (200) */
#define PSA_ALGO_RECT_QR 3 /* :rect_qr -- rectangular (n+1) x n upper Hessenberg with m <= 200 (compute.jl:588-590)   */
#define PSA_ALGO_SVD 4     /* :SVD     -- n < minlancs4psa (55): smallest singular value directly (compute.jl:478-519) */

/* indices into the counts[] array returned by psa_gpu_grid */
#define PSA_COUNT_NONCONVERGED 0 /* points where `lancz_converged` stayed false (maxit reached, or the bigsigma fallback,
                                    which `break`s before the convergence test, compute.jl:654-656): warning of 603-606 */
#define PSA_COUNT_EIGFAIL 1      /* points with the bigsigma fallback (non-finite recurrence): warning of compute.jl:607-610 */
#define PSA_COUNT_TOTAL_ITERS 2  /* sum over points of Lanczos iterations (flop accounting: 8 n^2 per iteration) */
#define PSA_COUNT_POINTS 3       /* points computed */
#define PSA_NCOUNTS 4

/* indices into the stats[] array of psa_gpu_last_stats */
#define PSA_STAT_SOLVE_MS 0      /* device time spent in the dominant kernel (CUDA events on the launching stream) */
#define PSA_STAT_SOLVE_LAUNCHES 1
#define PSA_STAT_TOTAL_LAUNCHES 2 /* all kernels launched by the last grid call */
#define PSA_STAT_GRID_MS 3       /* device time of the whole grid call, first launch to last */
#define PSA_STAT_TICKS 4         /* scheduler ticks (one Lanczos iteration of every active shift per tick); 0 when the grid
                                    ran as ONE launch: the :SVD branch, and sq_lanc with n <= ~163 (point-resident kernel) */
#define PSA_STAT_FLOPS 5         /* algorithmic real FP64 flops of the last call: 8 n^2 sum(iters) (+ QR for HessQR) */
#define PSA_STAT_BYTES 6         /* algorithmic bytes for the HBM-bound HessQR path, 0 for sq_lanc */
#define PSA_STAT_PACK_MS 7       /* device time of the one-time T packing in the last set_matrix */
#define PSA_STAT_MAX_DEVICE_MS 8 /* multi-device contexts: the slowest device's grid time (PSA_STAT_GRID_MS is device 0's) */
#define PSA_NSTATS 9

/* flags of psa_gpu_set_matrix_ex */
#define PSA_MATRIX_DEFAULT 0
#define PSA_MATRIX_UPPER_WRAPPER 1 /* T is the parent array of an `UpperTriangular` wrapper: ignore i > j without looking */

/* Create a context driving `ngpus` devices from this process (devices 0..ngpus-1, or device_ids[0..ngpus-1]
 * when device_ids is non-null).  With ngpus > 1 the grid rows are sharded cyclically over the devices
 * (SURVEY §8e); T is replicated by NCCL broadcast.  A one-process-per-GPU host (bench.py under torchrun)
 * uses ngpus = 1 per rank and shards rows itself. */
int psa_gpu_init(int ngpus, const int* device_ids, void** ctx);

/* Upload the (decomposed) matrix and pre-pack it for the kernels; it stays device-resident until the next
 * set_matrix / free so that zoom re-computations (src/zooming.jl) reuse it.
 *   m == n : T is upper triangular (`UpperTriangular{ComplexF64}` parent array); only i <= j is read, exactly
 *            as `UpperTriangular` ignores the strict lower part (compute.jl:600).                 -> :sq_lanc
 *   m == n+1, m > 200 : rectangular Hessenberg (xeigs.jl:226), lower bandwidth 2.                  -> :HessQR
 *   m == n+1, m <= 200, upper Hessenberg : the reference's `qr(T1)` branch (compute.jl:588-590).    -> :rect_qr
 *       (new_matrix only routes Hessenberg input here; dense non-Hessenberg rectangular input is turned into a
 *        generalised problem by rect_fact, src/Pseudospectra.jl:288-296, and stays on the host.)
 *   n < 55, m*n <= 12288 : per-point SVD, Z = sigma_min directly (compute.jl:510-518).              -> :SVD
 *   anything else (generalised S != I, non-Hessenberg rectangular, ...) -> PSA_ERR_UNSUPPORTED.
 * T: column-major, leading dimension ldt (>= m), interleaved complex128. */
int psa_gpu_set_matrix(void* ctx, const double* T, int64_t ldt, int m, int n);

/* Same with flags.  Without PSA_MATRIX_UPPER_WRAPPER a square T must satisfy `istriu(T)` -- the reference factorises
 * anything else (`if !istriu(T1) F1 = factorize(T1)`, compute.jl:597-601), which is outside the GPU path: such input
 * returns PSA_ERR_UNSUPPORTED so that the caller falls through to the host code.  Rectangular (n+1) x n input is
 * always checked for upper-Hessenberg structure (lower bandwidth 2 is what `bw == 2` promises, compute.jl:579). */
int psa_gpu_set_matrix_ex(void* ctx, const double* T, int64_t ldt, int m, int n, int flags);

/* The reference's mutable `ComputeThresholds` (compute.jl:18-30): minlancs4psa (55; the SVD branch is used for
 * n < minlancs4psa) and maxstdqr4hess (200; the Givens sweep is used for m > maxstdqr4hess, Householder-equivalent
 * QR below).  Raising minlancs4psa is the reference's own way to get SVD "reference solutions" (compute.jl:18-21);
 * the SVD kernel covers m*n <= 12288, larger shapes then return PSA_ERR_UNSUPPORTED at set_matrix.
 * Takes effect at the next psa_gpu_set_matrix*.  maxit_lancs is the `maxit` argument of psa_gpu_grid. */
int psa_gpu_set_thresholds(void* ctx, int minlancs4psa, int maxstdqr4hess);

/* Device-resident variant: d_T is a device pointer on the context's first device (same layout). */
int psa_gpu_set_matrix_device(void* ctx, const double* d_T, int64_t ldt, int m, int n);

/* The psacore replacement.  For every grid point z = x[k] + i*y[j] computes sigma_max of
 * (T-z)^{-1}(T-z)^{-H} by the inverse-Lanczos recurrence of `_psa_lanczos!` (compute.jl:619-665) started from
 * q0 (unit norm, length >= n; the same vector for every point, compute.jl:207-210,622) and returns
 * Z[j + k*ly] = 1/sqrt(sigma) (compute.jl:611), column-major ly x lx.
 *   tol, maxit, bigsigma : psatol (1e-5), thresholds.maxit_lancs (99), 0.1*floatmax (compute.jl:88,27,475);
 *                          1 <= maxit <= 4094 (PSA_ERR_ARG above: the per-shift alpha/beta history is sized from it)
 *   iters  : nullable, ly x lx int32, Lanczos iterations per point
 *   counts : nullable, PSA_NCOUNTS int64 (see PSA_COUNT_*)
 *   cancel : nullable; polled between scheduler ticks (single-launch paths: forwarded to the running kernel, which
 *            stops fetching grid points); *cancel == 1 -> PSA_ERR_CANCELLED (compute.jl:202-204)
 * Returns the algorithm used (PSA_ALGO_*) through *algo when non-null. */
int psa_gpu_grid(void* ctx, const double* x, int lx, const double* y, int ly, const double* q0, double tol,
                 int maxit, double bigsigma, double* Z, int32_t* iters, int64_t* counts, int* algo,
                 volatile int* cancel);

/* Device-side Schur reordering + truncation: the heavy part of the Trefethen/Wright projection `_maybe_project`
 * (compute.jl:339-357).  The caller selects the eigenvalues (compute.jl:295-316: cheap host logic on eigA) and passes
 * their 0-based, ascending positions on the diagonal of the resident square T; the library moves them to the top by
 * the reference's sequence of Givens swaps -- run as a systolic wavefront on the device instead of one after the other --
 * truncates to npick x npick and makes that the current matrix of the context (later grid / pseudomode calls use it).
 * The full matrix stays resident: another psa_gpu_project (a zoom with a different window) starts from it again, and
 * npick == n restores it.  Tproj: nullable, out, npick x npick column-major interleaved complex (ldp >= npick) -- the
 * `Tproj` of the reference's return tuple.  Square (:sq_lanc-class) matrices only; multi-device contexts broadcast the
 * projected matrix over NCCL. */
int psa_gpu_project(void* ctx, const int32_t* selection, int npick, double* Tproj, int64_t ldp);

/* Row-batched start vectors: the reference draws a NEW random unit q for every batch of rows
 * (`for j=ly:-step:1 ... q = randn(n) + randn(n)*im`, compute.jl:199-208).  q0s holds nbatch vectors of n
 * interleaved complex (vector b at offset b*n); row_batch[j] in [0, nbatch) names the vector used by grid row j
 * (for the reference's loop: row_batch[j] = (ly-1-j) / step with 0-based j).  With the shim drawing the vectors in
 * the reference's order the whole psa_compute is reproducible draw for draw.  psa_gpu_grid == nbatch 1.
 * Z may be NULL: the result then stays on the device for psa_gpu_postprocess. */
int psa_gpu_grid_batched(void* ctx, const double* x, int lx, const double* y, int ly, const double* q0s, int nbatch,
                         const int32_t* row_batch, double tol, int maxit, double bigsigma, double* Z, int32_t* iters,
                         int64_t* counts, int* algo, volatile int* cancel);

/* Device-side post-processing of the result of the last psa_gpu_grid* call, which is still resident on the
 * context's first device: un-mirror for real matrices (compute.jl:222-235), `clamp!(Z, ps_tiny, Inf)` (236-238) and
 * the reductions `recalc_levels` starts from (utils.jl:24-31).
 *   y        : the ly computed grid rows (as passed to psa_gpu_grid); n_mirror as returned by shift_axes (0: none)
 *   Zfull    : out, lyf x lx column-major (ldz >= lyf), lyf = ly + |n_mirror|;  yfull: out, lyf values
 *   zstats   : out, [0] minimum(Z), [1] maximum(Z), [2] minimum over Z >= small_sigma (utils.jl:30; +Inf if none)
 * Only one D2H copy of the final array is made. */
int psa_gpu_postprocess(void* ctx, const double* y, int ly, int lx, int n_mirror, double ps_tiny, double small_sigma,
                        double* Zfull, int64_t ldz, double* yfull, double* zstats);

/* Pseudomode at one point (psmode_inv_lanczos, src/modes.jl:185-274, dense branch n >= 200): the same
 * inverse-Lanczos recurrence on the resident square T, but the Lanczos basis is kept -- on the device -- and the
 * right singular vector for sigma_min(zI - A) is assembled from the Ritz vector of the largest eigenvalue
 * (`Q[:,1:end-1] * E.vectors[:,end]`, modes.jl:255-258).
 *   q0 : start vector, n interleaved complex;  tol, num_its as in the reference
 *   sigmin : out, Z = 1/sqrt(sigma);  mode : out, n interleaved complex;  iters : out (nullable)
 *   info : out, 0 = converged; 1 = num_its reached -- the reference then returns NaN or falls back to a dense SVD
 *          (modes.jl:262-272); sigmin/mode still hold the last iterate so the caller can choose.
 * Needs the :sq_lanc path (m == n >= minlancs4psa); n < 200 is the reference's plain-SVD branch and stays on the
 * host (PSA_ERR_UNSUPPORTED). */
int psa_gpu_pseudomode(void* ctx, double zre, double zim, const double* q0, double tol, int num_its, double* sigmin,
                       double* mode, int* iters, int* info);

/* Same computation with every array already resident in device memory of the context's first device
 * (single-device contexts only): d_x[lx], d_y[ly], d_q0[2n], d_Z[ly*lx], d_iters[ly*lx] (nullable).
 * counts stays a host pointer.  Used to time the path without host<->device traffic. */
int psa_gpu_grid_device(void* ctx, const double* d_x, int lx, const double* d_y, int ly, const double* d_q0,
                        double tol, int maxit, double bigsigma, double* d_Z, int32_t* d_iters, int64_t* counts,
                        int* algo, volatile int* cancel);

/* Diagnostic hook used by the parity tests: v_s = (T - z_s I)^{-1} (T - z_s I)^{-H} q_s for ns shifts
 * (one application of the operator at compute.jl:630, through the same solve kernel the grid uses).
 * z: ns interleaved complex; Q, V: ns vectors of n interleaved complex each (vector s at offset s*n). */
int psa_gpu_apply_resolvent(void* ctx, const double* z, int ns, const double* Q, double* V);

/* Timing / accounting of the last psa_gpu_grid* call (PSA_STAT_*), device 0 of the context. */
int psa_gpu_last_stats(void* ctx, double* stats);

int psa_gpu_free(void* ctx);

const char* psa_gpu_strerror(int status);
/* Text of the last CUDA / NCCL error seen by this context (empty string if none). */
const char* psa_gpu_last_error(void* ctx);
/* "psagpu <version> sm_100a" */
const char* psa_gpu_version(void);

/* Which machine a grid call would use for an m x n matrix with the default thresholds of `ComputeThresholds`
 * (compute.jl:20-27) and the given maxit -- pure host logic, no device needed, nothing is allocated.  Returns the
 * PSA_ALGO_* branch (PSA_ALGO_NONE: shape not covered) and, through *warps_per_sm when non-null, the number of grid
 * points one SM keeps in flight in the single-launch point-resident kernel (csrc/psa_point.cuh), or 0 when the shape
 * runs on the tick engine / the HessQR / SVD kernels.  The choice depends on (m, n, maxit) only, never on scheduling. */
int psa_gpu_plan(int m, int n, int maxit, int* warps_per_sm);

#ifdef __cplusplus
}
#endif
#endif /* PSA_GPU_H */
