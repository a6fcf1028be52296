"""CPU oracle for Pseudospectra.jl's resolvent-norm grid path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy *restatement* of the reference algorithm (RalphAS/Pseudospectra.jl,
`src/compute.jl`, `src/utils.jl`); every function cites the reference file:line it follows.  It is
the checker for the CUDA path, never the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.

PARITY PIN STATUS.  The reference is 100 % Julia, Julia is not installed in the build container, and the reference's
test-suite holds no golden Z arrays for this path (only `test/threads.jl:18`, a 1e-4 self-consistency check, and
`algo`-symbol checks), so bit-level parity with the Julia implementation itself remains UNPINNED.  The arithmetic lives
in un-vendored libraries: Julia's LinearAlgebra stdlib (compat "1", Project.toml:27) -> OpenBLAS `ztrsv` / LAPACK
`dsyevr`, `zgesdd`, `zgeqrf`.  This oracle calls the *same* library routines through scipy (OpenBLAS 0.3.x): `ztrsv` for
`F1\\(F1'\\q)` (compute.jl:630) and `dsyevr` (scipy.linalg.eigvalsh) for `eigvals(H[1:l,1:l])` (compute.jl:640).
What pins it:
 (i)   the ONE published known answer about sigma_min that the reference's test-suite holds: Grcar(100), eps = 1e-4,
       pseudospectral abscissa 2.41276 and radius 2.85216 to 1e-5 (Guglielmi & Overton 2012; test/radius_abscissa.jl:
       25-28) and the top eigenvalue 2.2617668 (:20) -- checked through the definition of the level set
       (tests/test_oracle.py: check_go) for the SVD checker, both Lanczos restatements and the GPU path;
 (ii)  dense `svdvals(T - zI)[-1]`, which is exactly the quantity the reference's own SVD branch returns
       (compute.jl:510-518), on every BASELINE matrix family and, every bench run, on the n = 4000 workload;
 (iii) the C restatement in oracle/psa_oracle.c (independent triangular solves + Sturm-bisection eigenvalues) and the
       OpenBLAS-worker arm (oracle/blas_arm.py) agreeing with this file to 1e-14;
 (iv)  the reference's own tests at this boundary restated with their tolerance (threads.jl 1e-4; medcplx / medrect /
       smrect `algo` symbols), and hand-derived known answers for the grid rules (tests/test_host_kat.py);
 (v)   closed-form sigma_min (a diagonal matrix: distance to the spectrum; lambda*I plus one corner entry: 2x2 closed
       form), tests/test_oracle.py: closed_form_cases -- implementation-independent known answers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import scipy.linalg as sla
from scipy.linalg import blas as _blas

# --------------------------------------------------------------------------------------
# thresholds                                                       src/compute.jl:20-30
# --------------------------------------------------------------------------------------


@dataclass
class ComputeThresholds:
    minlancs4psa: int = 55  # use SVD for n < this
    maxstdqr4hess: int = 200  # use HessQR for m > this (rectangular case)
    minnev: int = 20
    maxit_lancs: int = 99


DEFAULT_THRESHOLDS = ComputeThresholds()
FLOATMAX = np.finfo(np.float64).max
FLOATMIN = np.finfo(np.float64).tiny
BIGSIGMA = 0.1 * FLOATMAX  # src/compute.jl:475
PS_TINY = 10.0 * math.sqrt(FLOATMIN)  # src/compute.jl:236
SMALL_SIGMA = 1e-150  # `smallσ`, src/Pseudospectra.jl:46


def get_step_size(n: int, ly: int) -> int:
    """Rows per psacore batch for Float64.  src/compute.jl:270-279."""
    nsmall = 8
    if n < 100:
        return max(1, ly // nsmall)
    return min(ly, max(1, (4 * ly) // n))


# --------------------------------------------------------------------------------------
# grid construction                       src/compute.jl:122-145, src/utils.jl:112-141
# --------------------------------------------------------------------------------------


def _jl_range(a: float, b: float, n: int) -> np.ndarray:
    """`collect(range(a, stop=b, length=n))`.  Julia uses twice-precision stepping; numpy.linspace
    can differ in the last ulp -- the grid is an *input* at the GPU boundary, so this
    only matters for documentation-grade reproduction of the reference's axes."""
    return np.linspace(a, b, int(n))


def shift_axes(ax, npts: int):
    """Mesh for the imaginary axis exploiting conjugate symmetry.  src/utils.jl:112-141."""
    ax = [float(v) for v in ax]
    if ax[3] > -ax[2]:
        tpts = int(math.floor((ax[3] / (ax[3] - ax[2])) * npts))
        y = _jl_range(ax[3] / (2 * (tpts - 1) + 1), ax[3], tpts)
        n = int(np.count_nonzero(y < -ax[2]))
        if abs(y[min(len(y), n + 1) - 1] + ax[2]) > 1e-15:
            y = np.concatenate([y[:n], [-ax[2]], y[n:]])
        num_mirror = n + 1
    elif ax[3] < -ax[2]:
        bpts = int(math.floor((ax[2] / (ax[2] - ax[3])) * npts))
        y = _jl_range(ax[2], ax[2] / (2 * (bpts - 1) + 1), bpts)
        n = int(np.count_nonzero(y < -ax[3]))
        if abs(y[max(1, n) - 1] + ax[3]) > 1e-15:
            y = np.concatenate([y[:n], [-ax[3]], y[n:]])
        num_mirror = -(len(y) - (n + 1) + 1)
    else:
        if npts % 2 == 1:
            y = _jl_range(0.0, ax[3], int(round((npts + 1) / 2)))
            num_mirror = ((npts + 1) >> 1) - 1
        else:
            step = (ax[3] - ax[2]) / npts
            y = _jl_range(step / 2, ax[3], int(round(npts / 2)))
            num_mirror = npts >> 1
    return np.asarray(y, dtype=np.float64), int(num_mirror)


def make_grid(ax, npts: int, real_matrix: bool, scale_equal: bool = False):
    """x, y, n_mirror_pts exactly as psa_compute builds them.  src/compute.jl:122-145."""
    ax = [float(v) for v in ax]
    if scale_equal:
        y_dist = ax[3] - ax[2]
        x_dist = ax[1] - ax[0]
        if x_dist > y_dist:
            x_npts = npts
            y_npts = max(5, int(math.ceil(y_dist / x_dist * npts)))
        else:
            y_npts = npts
            x_npts = max(5, int(math.ceil(x_dist / y_dist * npts)))
    else:
        x_npts = y_npts = npts
    if real_matrix and ax[3] > 0 and ax[2] < 0:
        y, n_mirror = shift_axes(ax, y_npts)
    else:
        n_mirror = 0
        y = _jl_range(ax[2], ax[3], y_npts)
    x = _jl_range(ax[0], ax[1], x_npts)
    return x, y, n_mirror


def unmirror(Z: np.ndarray, y: np.ndarray, n_mirror: int):
    """Map data (and y) back when symmetry was used.  src/compute.jl:222-235."""
    if n_mirror < 0:
        k = -n_mirror
        Z = np.vstack([Z, Z[Z.shape[0] - k:, :][::-1, :]])
        y = np.concatenate([y, -y[len(y) - k:][::-1]])
    else:
        if y[0] != 0:
            Z = np.vstack([Z[:n_mirror, :][::-1, :], Z])
            y = np.concatenate([-y[:n_mirror][::-1], y])
        else:
            Z = np.vstack([Z[1:n_mirror + 1, :][::-1, :], Z])
            y = np.concatenate([-y[1:n_mirror + 1][::-1], y])
    return Z, y


def tidyaxes(vmin: float, vmax: float, tol: float):
    """Round [vmin, vmax] outward to cleaner decimals.  src/utils.jl:237-248."""
    span = vmax - vmin
    round_to = span * tol
    lround_to = math.log10(round_to)
    expon = math.floor(lround_to)
    mant = round(10 ** (lround_to - expon))
    round_to = mant * 10.0 ** expon
    return round_to * math.floor(vmin / round_to), round_to * math.ceil(vmax / round_to)


def vec2ax(dispvec):
    """Axis limits around a vector of points.  src/utils.jl:175-191."""
    v = np.asarray(dispvec)
    ca = [v.real.min(), v.real.max(), v.imag.min(), v.imag.max()]
    if ca[0] == ca[1]:
        ca[0] -= 0.5
        ca[1] += 0.5
    if ca[2] == ca[3]:
        ca[2] -= 0.5
        ca[3] += 0.5
    ext = max(ca[1] - ca[0], ca[3] - ca[2]) / 3
    ax = [ca[0] - ext, ca[1] + ext, ca[2] - ext, ca[3] + ext]
    ax[0], ax[1] = tidyaxes(ax[0], ax[1], 0.02)
    ax[2], ax[3] = tidyaxes(ax[2], ax[3], 0.02)
    return [float(a) for a in ax]


def recalc_levels(sigmin: np.ndarray, ax):
    """Contour levels for log10(sigma).  src/utils.jl:21-73."""
    err = 0
    smin, smax = float(np.min(sigmin)), float(np.max(sigmin))
    if smax <= SMALL_SIGMA:
        return np.zeros(0), -2
    if smin <= SMALL_SIGMA:
        smin = float(np.min(sigmin[sigmin >= SMALL_SIGMA]))
    max_val = math.log10(smax)
    min_val = math.log10(smin)
    num_lines = 8
    scale = min(ax[3] - ax[2], ax[1] - ax[0])
    last_lev = math.log10(0.03 * scale)
    if max_val >= last_lev:
        num_lines = int(math.ceil(num_lines * (last_lev - min_val) / (max_val - min_val)))
        max_val = last_lev
    if max_val < min_val:
        max_val = math.log10(smin) + 0.1 * math.log10(smax / smin)
        num_lines = 3
    max_lines = num_lines
    levels = np.zeros(0)
    if num_lines > 0:
        jround = lambda v: float(np.round(v))  # Julia round(): ties-to-even, as numpy
        stepsize = max(jround((max_val - min_val) / num_lines * 4) / 4, 0.25)
        if stepsize >= 0.75:
            stepsize = 1.0
        if (max_val - min_val) / stepsize < max_lines:
            stepsize = max(jround((max_val - min_val) / num_lines * 10) / 10, 0.1)
            if stepsize == 0.3:
                stepsize = 0.2
            if stepsize == 0.4:
                stepsize = 0.5
            if 0.6 <= stepsize <= 0.8:
                stepsize = 0.5
            if stepsize >= 0.9:
                stepsize = 1.0
        if (max_val - min_val) / stepsize < math.ceil(max_lines / 2):
            stepsize = (max_val - min_val) / max_lines
        else:
            min_val = math.ceil(min_val / stepsize) * stepsize
            max_val = math.floor(max_val / stepsize) * stepsize
        if stepsize > 0:
            nlev = int(math.floor((max_val - min_val) / stepsize + 1e-12)) + 1
            levels = min_val + stepsize * np.arange(max(nlev, 0))
        else:
            levels = np.array([min_val, max_val])
    ll = len(levels)
    stride = max(ll // 9, 1)
    levels = levels[::-1][::stride][::-1]
    if len(levels) == 1:
        levels = levels * np.ones(2)
    return levels, err


# --------------------------------------------------------------------------------------
# the inverse-Lanczos kernel                                      src/compute.jl:619-665
# --------------------------------------------------------------------------------------


def psa_lanczos(F1: np.ndarray, q0: np.ndarray, tol: float = 1e-5, maxit: int = 99,
                bigsigma: float = BIGSIGMA):
    """sigma_max of (F1^-1 F1^-H) for upper-triangular F1 (Fortran-ordered complex128).

    Returns (sigma, converged, eig_failed, iters).  Follows src/compute.jl:619-665 line by line:
    v = F1\\(F1'\\q) - beta*qold (630), alpha = real(q.v) (631), v -= alpha q (632), beta = |v| (633),
    H tridiagonal fill (636-638), sigma = max eigvals(H[1:l,1:l]) (640-642), stop on
    |sigma_old/sigma - 1| < tol or beta == 0 (658-661).

    Defined behaviour where the reference has none (SURVEY §7 'Overflow'): a non-finite alpha/beta makes
    Julia's `eigvals` throw ArgumentError (re-thrown at compute.jl:651-653); this oracle, the C
    restatement and the CUDA path all map that case to sigma = bigsigma, eig_failed = True.
    """
    n = F1.shape[1]
    q = np.array(q0[:n], dtype=np.complex128)
    qold = np.zeros(n, dtype=np.complex128)
    beta = 0.0
    sigold = 0.0
    sigma = bigsigma
    H = np.zeros((maxit + 1, maxit + 1))
    converged = False
    failed = False
    iters = 0
    with np.errstate(all="ignore"):
        for l in range(1, maxit + 1):
            iters = l
            w = _blas.ztrsv(F1, q, lower=0, trans=2, diag=0)  # F1' \ q
            v = _blas.ztrsv(F1, w, lower=0, trans=0, diag=0)  # F1 \ w
            v = v - beta * qold
            alpha = float(np.real(np.vdot(q, v)))
            v = v - alpha * q
            beta = float(np.linalg.norm(v))
            qold = q
            q = v / beta
            H[l, l - 1] = beta
            H[l - 1, l] = beta
            H[l - 1, l - 1] = alpha
            if not (math.isfinite(alpha) and math.isfinite(beta)):
                failed = True
                sigma = bigsigma
                break
            try:
                ew = sla.eigvalsh(H[:l, :l], check_finite=False, driver="evr")
                sigma = float(np.max(ew))
            except sla.LinAlgError:  # LAPACKException branch, compute.jl:651-656
                failed = True
                sigma = bigsigma
                break
            if abs(sigold / sigma - 1) < tol or beta == 0:
                converged = True
                break
            sigold = sigma
    return sigma, converged, failed, iters


def givens(f: complex, g: complex):
    """(c, s, r) with [c s; -conj(s) c] [f; g] = [r; 0]; c real.  Mirrors Julia's
    `LinearAlgebra.givens(f,g,1,2)` (a zlartg port) as used at src/compute.jl:585."""
    if g == 0:
        return 1.0, 0.0 + 0.0j, f
    if f == 0:
        ag = abs(g)
        return 0.0, np.conj(g) / ag, ag + 0.0j
    af = abs(f)
    d = math.hypot(af, abs(g))
    c = af / d
    ph = f / af
    s = ph * np.conj(g) / d
    r = ph * d
    return c, s, r


def hessqr_triu(Tshift: np.ndarray) -> np.ndarray:
    """Givens sweep of src/compute.jl:581-587 followed by `triu(T1[1:n,1:n])` (592).

    NOTE the reference's quirk: the sweep runs jj = 1..n-1, so for m = n+1 the last
    sub-diagonal entry H[n+1,n] is never rotated in and is dropped by the truncation."""
    T1 = np.array(Tshift, dtype=np.complex128, order="F")
    m, n = T1.shape
    for jj in range(n - 1):
        c, s, _ = givens(T1[jj, jj], T1[jj + 1, jj])
        a = T1[jj, jj:].copy()
        b = T1[jj + 1, jj:].copy()
        T1[jj, jj:] = c * a + s * b
        T1[jj + 1, jj:] = -np.conj(s) * a + c * b
    return np.asfortranarray(np.triu(T1[:n, :n]))


def psacore(T: np.ndarray, q0: np.ndarray, x: np.ndarray, y: np.ndarray, bw: int, tol: float = 1e-5,
            thresholds: ComputeThresholds = DEFAULT_THRESHOLDS, want_iters: bool = False):
    """Serial `psacore` for S = I.  src/compute.jl:448-617.

    Returns (Z, algo, warned[, iters]); Z is (len(y), len(x)), Z[j,k] = 1/sqrt(sigma) (611) or the
    smallest singular value on the SVD branch (516)."""
    Twork = np.array(T, dtype=np.complex128, order="F")  # 450-454
    lx, ly = len(x), len(y)
    m, n = Twork.shape
    if m < n:
        raise ValueError("Matrix size must be m x n with m >= n")  # 459-461
    Z = np.zeros((ly, lx))
    iters = np.zeros((ly, lx), dtype=np.int32)
    diaga = np.diagonal(Twork).copy()
    warned = [False, False]
    idx = np.arange(n)
    algo = None
    if n < thresholds.minlancs4psa:  # 478-519
        algo = "SVD"
        for j in range(ly):
            for k in range(lx):
                zpt = x[k] + 1j * y[j]
                Twork[idx, idx] = diaga - zpt
                Z[j, k] = sla.svdvals(Twork)[-1]
    else:
        maxit = thresholds.maxit_lancs
        T1 = Twork.copy(order="F") if m == n else None
        for j in range(ly):
            for k in range(lx):
                zpt = x[k] + 1j * y[j]
                if m != n:
                    Twork[idx, idx] = diaga - zpt  # 573
                    if bw == 2 and m > thresholds.maxstdqr4hess:  # 579-587
                        algo = "HessQR"
                        F1 = hessqr_triu(Twork)
                    else:  # 589-590
                        algo = "rect_qr"
                        R = np.linalg.qr(Twork, mode="r")
                        F1 = np.asfortranarray(np.triu(R[:n, :n]))
                else:
                    algo = "sq_lanc"
                    T1[idx, idx] = diaga - zpt  # 595
                    F1 = T1
                sig, conv, fail, it = psa_lanczos(F1, q0, tol, maxit, BIGSIGMA)
                warned[0] |= not conv
                warned[1] |= fail
                with np.errstate(all="ignore"):
                    Z[j, k] = 1.0 / math.sqrt(sig) if sig > 0 else np.inf
                iters[j, k] = it
    if want_iters:
        return Z, algo, warned, iters
    return Z, algo, warned


def psa_compute(T: np.ndarray, npts: int, ax, eigA, opts: dict | None = None, *, psatol: float = 1e-5,
                q_for_batch=None, thresholds: ComputeThresholds = DEFAULT_THRESHOLDS):
    """Dense branch of `psa_compute` (S = I, proj_lev = Inf).  src/compute.jl:87-268.

    `q_for_batch(batch_index, n)` supplies the unit start vector the reference draws with
    `randn` at compute.jl:207-208 (Julia's RNG stream cannot be reproduced, so q is an input).
    Returns the reference tuple (Z, x, y, levels, err, Tproj, eigAproj, algo)."""
    opts = dict(opts or {})
    m, n = T.shape
    real_matrix = opts.get("real_matrix", not np.iscomplexobj(T))
    x, y, n_mirror = make_grid(ax, npts, real_matrix, opts.get("scale_equal", False))
    lx, ly = len(x), len(y)
    Z = np.full((ly, lx), np.inf)
    if q_for_batch is None:
        rng = np.random.default_rng(0)

        def q_for_batch(_b, nn):
            q = rng.standard_normal(nn) + 1j * rng.standard_normal(nn)
            return q / np.linalg.norm(q)
    step = get_step_size(m, ly)
    algo = None
    b = 0
    j = ly
    while j >= 1:  # for j = ly:-step:1, compute.jl:200
        last_y = max(j - step + 1, 1)
        q = q_for_batch(b, n)
        rows = np.arange(j, last_y - 1, -1) - 1  # j:-1:last_y (0-based)
        Zb, algo, _ = psacore(T, q, x, y[rows], m - n + 1, tol=psatol, thresholds=thresholds)
        Z[rows, :] = Zb
        j -= step
        b += 1
    Z, y = unmirror(Z, y, n_mirror)  # no-op when n_mirror == 0, as in the reference
    Z = np.clip(Z, PS_TINY, np.inf)  # 236-238
    err = 0
    levels = opts.get("levels")
    if levels is None or opts.get("recompute_levels", levels is None):
        levels, err = recalc_levels(Z, ax)
    return Z, x, y, levels, err, T, np.array(eigA), algo


def psmode_inv_lanczos(A: np.ndarray, q0: np.ndarray, z: complex, tol: float, num_its: int):
    """Pseudomode by inverse Lanczos for upper-triangular A, n >= 200 branch.  src/modes.jl:201-261
    (`v = T1 \\ (T2 \\ q) - beta*qold` at 229, Ritz vector `Q[:,1:end-1] * E.vectors[:,end]` at 257)."""
    n = A.shape[0]
    T1 = np.asfortranarray(np.triu(np.asarray(A, dtype=np.complex128)))
    T1[np.arange(n), np.arange(n)] -= z
    q = np.asarray(q0, dtype=np.complex128).copy()
    qold = np.zeros(n, dtype=np.complex128)
    Q = [q.copy()]
    H = np.zeros((num_its + 1, num_its + 1))
    beta = sigold = 0.0
    sigma, l = None, 0
    for _ in range(num_its):
        l += 1
        w = _blas.ztrsv(T1, q, lower=0, trans=2, diag=0)
        v = _blas.ztrsv(T1, w, lower=0, trans=0, diag=0) - beta * qold
        alpha = float(np.real(np.vdot(q, v)))
        v = v - alpha * q
        beta = float(np.linalg.norm(v))
        qold = q
        q = v / beta
        Q.append(q.copy())
        H[l, l - 1] = H[l - 1, l] = beta
        H[l - 1, l - 1] = alpha
        sigma = float(np.max(sla.eigvalsh(H[:l, :l])))
        if abs(sigold / sigma - 1) < tol or beta == 0:
            break
        sigold = sigma
    _, vecs = sla.eigh(H[:l, :l])
    return 1.0 / math.sqrt(sigma), np.column_stack(Q[:l]) @ vecs[:, -1], l


def sigmin_svd(T: np.ndarray, z: complex) -> float:
    """Independent checker: smallest singular value of (T - zI) (same quantity as compute.jl:510-518)."""
    m, n = T.shape
    A = np.array(T, dtype=np.complex128)
    A[np.arange(n), np.arange(n)] -= z
    return float(sla.svdvals(A)[-1])


# --------------------------------------------------------------------------------------
# input generators for the BASELINE configs
# --------------------------------------------------------------------------------------


def readme_matrix(n: int = 150) -> np.ndarray:
    """README.md:40-42:  B = diagm(1=>2im, 2=>-1, 3=>2, -2=>-4, -3=>-2im)."""
    B = np.zeros((n, n), dtype=np.complex128)
    B += np.diag(np.full(n - 1, 2j), 1)
    B += np.diag(np.full(n - 2, -1.0), 2)
    B += np.diag(np.full(n - 3, 2.0), 3)
    B += np.diag(np.full(n - 2, -4.0), -2)
    B += np.diag(np.full(n - 3, -2j), -3)
    return B


def grcar(N: int) -> np.ndarray:
    """examples/demo_mtx.jl:30-34."""
    return (np.eye(N) - np.diag(np.ones(N - 1), -1) + np.diag(np.ones(N - 1), 1)
            + np.diag(np.ones(N - 2), 2) + np.diag(np.ones(N - 3), 3))


def convdiff_fd(N: int):
    """examples/demo_mtx.jl:161-191 (scipy.sparse CSR, dimension 4 N^2)."""
    import scipy.sparse as sp
    e = 1e-2
    w1, w2 = 0.0, 1.0
    n1, n2 = N, 4 * N
    h1, h2 = 1 / (n1 + 1), 1 / (n2 + 1)
    alpha1, beta1 = w1 / h1, 2 * e / h1 ** 2
    alpha2, beta2 = w2 / h2, 2 * e / h2 ** 2
    A1 = sp.diags([(-alpha1 - beta1) * np.ones(n1 - 1), 2 * beta1 * np.ones(n1), (alpha1 - beta1) * np.ones(n1 - 1)],
                  [-1, 0, 1])
    A2 = sp.diags([(-alpha2 - beta2) * np.ones(n2 - 1), 2 * beta2 * np.ones(n2), (alpha2 - beta2) * np.ones(n2 - 1)],
                  [-1, 0, 1])
    return (sp.kron(sp.identity(n2), A1) + sp.kron(A2, sp.identity(n1))).tocsr()


def landau_fox_li(N: int, F: float = 0) -> np.ndarray:
    """examples/demo_mtx.jl:260-280: Landau's discretisation of the Fox-Li laser-cavity operator (Gauss-Legendre nodes
    and weights from the Jacobi matrix); the input of the reference's test/medcplx.jl (N = 225, F = 16)."""
    if F == 0:
        F = 32 if N > 200 else 12
    k = np.arange(1, N)
    beta = 0.5 * (1 - (2.0 * k) ** (-2.0)) ** (-0.5)
    d, V = np.linalg.eigh(np.diag(beta, 1) + np.diag(beta, -1))
    idx = np.argsort(d)
    nodes = d[idx]
    w = 2 * V[0, idx] ** 2
    B = w[None, :] * np.sqrt(F * 1j) * np.exp(-1j * np.pi * F * (nodes[:, None] - nodes[None, :]) ** 2)
    sw = np.sqrt(w)
    return (sw[:, None] * B) / sw[None, :]


def toeplitz_from_symbol(vc, vr, n: int) -> np.ndarray:
    """Banded Toeplitz matrix with first column `vc` and first row `vr` (examples/toeplitz.jl:45-67:
    vc/vr hold the coefficients, x^0 terms equal)."""
    c = np.zeros(n, dtype=np.result_type(np.asarray(vc).dtype, np.asarray(vr).dtype, np.float64))
    r = c.copy()
    c[:len(vc)] = vc
    r[:len(vr)] = vr
    return sla.toeplitz(c, r)


def random_complex(n: int, seed: int = 1) -> np.ndarray:
    """Config 3: A = N(0,1) + i N(0,1), numpy default_rng(seed)  (SURVEY §8d)."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))


def complex_schur(A: np.ndarray):
    """`schur(A)` complexified, as new_matrix does (src/Pseudospectra.jl:206-213): returns (T, eigA)."""
    T, _ = sla.schur(np.asarray(A, dtype=np.complex128), output="complex")
    T = np.asfortranarray(np.triu(T))
    return T, np.diagonal(T).copy()


def start_vector(n: int, seed: int = 0) -> np.ndarray:
    """Unit start vector q0 = normalize(randn + i randn) (src/compute.jl:207-208), seeded."""
    rng = np.random.default_rng(seed)
    q = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    return q / np.linalg.norm(q)


def arnoldi_hessenberg(A, m: int, seed: int = 0) -> np.ndarray:
    """(m+1) x m rectangular Hessenberg from an m-step Arnoldi process with a seeded start vector.
    Stand-in for the matrix `extract_mtx` builds from ARPACK's workspace (src/xeigs.jl:196-227):
    same shape; the reference's last row is [0 ... 0 beta] (xeigs.jl:226), which Arnoldi's
    H[m+1, :] reproduces."""
    n = A.shape[0]
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(n)
    v /= np.linalg.norm(v)
    V = np.zeros((n, m + 1))
    H = np.zeros((m + 1, m))
    V[:, 0] = v
    for j in range(m):
        w = A @ V[:, j]
        for _ in range(2):  # classical Gram-Schmidt with one re-orthogonalisation
            h = V[:, :j + 1].T @ w
            w = w - V[:, :j + 1] @ h
            H[:j + 1, j] += h
        H[j + 1, j] = np.linalg.norm(w)
        V[:, j + 1] = w / H[j + 1, j]
    return H
