/* CPU oracle (C restatement) for Pseudospectra.jl's resolvent-norm grid path.
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs -- never by the product path.
 *
 * PARITY with the Julia implementation itself is UNPINNED (no Julia here, no golden Z vectors in the
 * reference's tests); see oracle/psa_oracle.py for what pins the oracle instead: the Guglielmi-Overton
 * Grcar(100) values the reference's test-suite holds (test/radius_abscissa.jl:20-28), dense svdvals,
 * and three independent restatements agreeing to 1e-14.
 *
 * Restates, with plain loops instead of BLAS/LAPACK calls:
 *   _psa_lanczos!            /root/reference/src/compute.jl:619-665
 *   _Sq_Lancs_Server + the threaded branch of psacore (work queue of (j,k) points, one worker per
 *   thread, diagonal shifted per point)        src/compute.jl:388-417, 534-562
 *   psacore HessQR branch (Givens sweep jj=1..n-1, triu(T1[1:n,1:n]))   src/compute.jl:571-592
 * The triangular solves stand in for OpenBLAS ztrsv('U','C') / ztrsv('U','N') (compute.jl:630);
 * the tridiagonal eigenvalue is a Sturm-count bisection standing in for eigvals -> dsyevr
 * (compute.jl:640); the numpy oracle calls dsyevr itself, so the two oracles cross-check it.
 *
 * Build: gcc -O3 -fPIC -shared -pthread -o oracle/libpsa_oracle.so oracle/psa_oracle.c -lm
 */
#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef double _Complex zc;

/* largest eigenvalue of the l x l symmetric tridiagonal (d, e): Sturm-count bisection on the
 * matrix scaled to unit max-norm (so e^2 cannot overflow), to the last representable midpoint. */
static int sturm_count(int l, const double* d, const double* e2, double x) {
  int cnt = 0;
  double q = d[0] - x;
  if (q == 0.0) q = -1e-300;
  if (q < 0) ++cnt;
  for (int i = 1; i < l; ++i) {
    q = d[i] - x - e2[i - 1] / q;
    if (q == 0.0) q = -1e-300;
    if (q < 0) ++cnt;
  }
  return cnt; /* number of eigenvalues < x */
}

static int tridiag_max_eig(int l, const double* d_in, const double* e_in, double* out) {
  double dbuf[128], ebuf[128];
  double *d = dbuf, *e2 = ebuf;
  if (l > 127) { /* maxit_lancs is mutable in the reference (compute.jl:27): no fixed limit here either */
    d = (double*)malloc(sizeof(double) * 2 * (size_t)l);
    if (!d) return -1;
    e2 = d + l;
  }
  double sc = 0;
  for (int i = 0; i < l; ++i) sc = fmax(sc, fabs(d_in[i]));
  for (int i = 0; i + 1 < l; ++i) sc = fmax(sc, fabs(e_in[i]));
  if (sc == 0.0) { *out = 0.0; if (d != dbuf) free(d); return 0; }
  double lo = 1e300, hi = -1e300;
  for (int i = 0; i < l; ++i) {
    d[i] = d_in[i] / sc;
    double el = (i > 0) ? fabs(e_in[i - 1] / sc) : 0.0, er = (i + 1 < l) ? fabs(e_in[i] / sc) : 0.0;
    if (i + 1 < l) e2[i] = er * er;
    lo = fmin(lo, d[i] - el - er);
    hi = fmax(hi, d[i] + el + er);
  }
  hi += 1e-15 * fmax(fabs(hi), 1.0) + 1e-300; /* count(hi) == l */
  for (int it = 0; it < 200; ++it) {
    double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (sturm_count(l, d, e2, mid) >= l) hi = mid; else lo = mid;
  }
  *out = 0.5 * (lo + hi) * sc;
  if (d != dbuf) free(d);
  return 0;
}

/* w = (U - z I)^{-H} q  : forward substitution, column (dot) form on col-major U, ld = ldt */
static void solve_UH(int n, const zc* U, int64_t ldt, zc z, const zc* q, zc* w) {
  for (int j = 0; j < n; ++j) {
    const zc* col = U + (int64_t)j * ldt;
    double sr = 0, si = 0;
    for (int i = 0; i < j; ++i) { /* sum conj(U[i,j]) * w[i] */
      double ar = creal(col[i]), ai = -cimag(col[i]);
      double br = creal(w[i]), bi = cimag(w[i]);
      sr += ar * br - ai * bi;
      si += ar * bi + ai * br;
    }
    zc num = q[j] - (sr + si * I);
    w[j] = num / conj(col[j] - z);
  }
}

/* v = (U - z I)^{-1} w : back substitution, column (axpy) form */
static void solve_U(int n, const zc* U, int64_t ldt, zc z, zc* v /* in: w, out: v */) {
  for (int j = n - 1; j >= 0; --j) {
    const zc* col = U + (int64_t)j * ldt;
    zc xj = v[j] / (col[j] - z);
    v[j] = xj;
    double xr = creal(xj), xi = cimag(xj);
    for (int i = 0; i < j; ++i) {
      double ar = creal(col[i]), ai = cimag(col[i]);
      v[i] -= (ar * xr - ai * xi) + (ar * xi + ai * xr) * I;
    }
  }
}

/* src/compute.jl:619-665.  U is upper triangular (strict lower part ignored), its diagonal is
 * taken as (U[i,i] - z) so no per-point copy is needed.  work: 3n complex.  Returns sigma. */
double psa_oracle_lanczos(int n, const zc* U, int64_t ldt, double zre, double zim, const zc* q0, double tol,
                          int maxit, double bigsigma, zc* work, int32_t* iters_out, int32_t* flags_out) {
  zc z = zre + zim * I;
  zc *q = work, *qold = work + n, *v = work + 2 * (int64_t)n;
  double abuf[128], bbuf[128];
  double *alpha = abuf, *beta_ = bbuf;
  if (maxit > 126) {
    alpha = (double*)malloc(sizeof(double) * 2 * ((size_t)maxit + 2));
    beta_ = alpha + maxit + 2;
  }
  for (int i = 0; i < n; ++i) { q[i] = q0[i]; qold[i] = 0; }
  double beta = 0.0, sigold = 0.0, sigma = bigsigma;
  int converged = 0, failed = 0, l;
  for (l = 1; l <= maxit; ++l) {
    solve_UH(n, U, ldt, z, q, v);
    solve_U(n, U, ldt, z, v);
    double a = 0;
    for (int i = 0; i < n; ++i) { v[i] -= beta * qold[i]; a += creal(q[i]) * creal(v[i]) + cimag(q[i]) * cimag(v[i]); }
    double nrm = 0;
    for (int i = 0; i < n; ++i) { v[i] -= a * q[i]; nrm += creal(v[i]) * creal(v[i]) + cimag(v[i]) * cimag(v[i]); }
    beta = sqrt(nrm);
    for (int i = 0; i < n; ++i) { qold[i] = q[i]; q[i] = v[i] / beta; }
    alpha[l - 1] = a; beta_[l - 1] = beta;
    if (!isfinite(a) || !isfinite(beta)) { failed = 1; sigma = bigsigma; break; }
    if (tridiag_max_eig(l, alpha, beta_, &sigma) != 0) { failed = 1; sigma = bigsigma; break; }
    if (fabs(sigold / sigma - 1.0) < tol || beta == 0.0) { converged = 1; break; }
    sigold = sigma;
  }
  if (l > maxit) l = maxit;
  if (alpha != abuf) free(alpha);
  if (iters_out) *iters_out = l;
  if (flags_out) *flags_out = (converged ? 0 : 1) | (failed ? 2 : 0);
  return sigma;
}

/* zlartg-convention Givens: [c s; -conj(s) c] [f; g] = [r; 0]   (Julia givens(f,g,1,2), compute.jl:585) */
static void givens(zc f, zc g, double* c, zc* s, zc* r) {
  if (g == 0) { *c = 1; *s = 0; *r = f; return; }
  if (f == 0) { double ag = cabs(g); *c = 0; *s = conj(g) / ag; *r = ag; return; }
  double af = cabs(f), d = hypot(af, cabs(g));
  zc ph = f / af;
  *c = af / d; *s = ph * conj(g) / d; *r = ph * d;
}

/* HessQR point: R = triu((G_{n-1}...G_1 (H - z I_{m,n}))[1:n,1:n]) then Lanczos on R.
 * H is m x n col-major (ldh), Rwork is n*n complex, work 3n complex. */
double psa_oracle_hessqr_point(int m, int n, const zc* H, int64_t ldh, double zre, double zim, const zc* q0,
                               double tol, int maxit, double bigsigma, zc* Rwork, zc* work, int32_t* iters_out,
                               int32_t* flags_out) {
  zc z = zre + zim * I;
  zc* T1 = (zc*)malloc(sizeof(zc) * (size_t)m * n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) T1[(size_t)j * m + i] = H[(int64_t)j * ldh + i] - (i == j ? z : 0);
  for (int jj = 0; jj < n - 1; ++jj) {
    double c; zc s, r;
    givens(T1[(size_t)jj * m + jj], T1[(size_t)jj * m + jj + 1], &c, &s, &r);
    for (int k = jj; k < n; ++k) {
      zc a = T1[(size_t)k * m + jj], b = T1[(size_t)k * m + jj + 1];
      T1[(size_t)k * m + jj] = c * a + s * b;
      T1[(size_t)k * m + jj + 1] = -conj(s) * a + c * b;
    }
  }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) Rwork[(size_t)j * n + i] = (i <= j) ? T1[(size_t)j * m + i] : 0;
  free(T1);
  return psa_oracle_lanczos(n, Rwork, n, 0.0, 0.0, q0, tol, maxit, bigsigma, work, iters_out, flags_out);
}

typedef struct {
  int m, n; const zc* T; int64_t ldt; const double *x, *y; int lx, ly; const zc* q0;
  double tol; int maxit; double bigsigma; double* Z; int32_t* iters; int32_t* flags;
  atomic_long* next; long total;
} job_t;

static void* worker(void* arg) {
  job_t* jb = (job_t*)arg;
  int n = jb->n;
  zc* work = (zc*)malloc(sizeof(zc) * 3 * (size_t)n);
  zc* Rwork = (jb->m != n) ? (zc*)malloc(sizeof(zc) * (size_t)n * n) : NULL;
  for (;;) {
    long p = atomic_fetch_add(jb->next, 1); /* the Channel{Tuple{Int,Int}} of compute.jl:540-555 */
    if (p >= jb->total) break;
    int j = (int)(p / jb->lx), k = (int)(p % jb->lx);
    int32_t it, fl;
    double sig;
    if (jb->m == n)
      sig = psa_oracle_lanczos(n, jb->T, jb->ldt, jb->x[k], jb->y[j], jb->q0, jb->tol, jb->maxit, jb->bigsigma, work, &it, &fl);
    else
      sig = psa_oracle_hessqr_point(jb->m, n, jb->T, jb->ldt, jb->x[k], jb->y[j], jb->q0, jb->tol, jb->maxit,
                                    jb->bigsigma, Rwork, work, &it, &fl);
    jb->Z[(size_t)k * jb->ly + j] = 1.0 / sqrt(sig); /* col-major ly x lx, compute.jl:410/611 */
    if (jb->iters) jb->iters[(size_t)k * jb->ly + j] = it;
    if (jb->flags) jb->flags[(size_t)k * jb->ly + j] = fl;
  }
  free(work); free(Rwork);
  return NULL;
}

/* Threaded grid: Z is ly x lx column-major (rows = y).  m == n -> sq_lanc, m == n+1 -> HessQR. */
int psa_oracle_grid(int m, int n, const zc* T, int64_t ldt, const double* x, int lx, const double* y, int ly,
                    const zc* q0, double tol, int maxit, double bigsigma, int nthreads, double* Z, int32_t* iters,
                    int32_t* flags) {
  atomic_long next = 0;
  job_t jb = {m, n, T, ldt, x, y, lx, ly, q0, tol, maxit, bigsigma, Z, iters, flags, &next, (long)lx * ly};
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, &jb);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th);
  return 0;
}
