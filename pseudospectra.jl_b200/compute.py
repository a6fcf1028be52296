"""Host-side mirror of the reference interface for the hot path: `psacore` and the dense branch of
`psa_compute` (/root/reference/src/compute.jl:87-268, 448-617), with the per-point work done by libpsagpu.

Same names, argument meaning and error behaviour as the reference; the differences the GPU option implies
 are: one start vector q for the whole grid instead of one per row batch (compute.jl:206-208), and
all rows handed to the library in a single call.  No CPU fallback: shapes the library does not cover raise
PsaGpuUnsupported (the Julia shim falls through to the stock code there)."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from ._lib import PSA_ERR_UNSUPPORTED, PsaGpu, PsaGpuCancelled, PsaGpuUnsupported

_FLOATMAX = float(np.finfo(np.float64).max)
_FLOATMIN = float(np.finfo(np.float64).tiny)
_SMALL_SIGMA = 1e-150  # src/Pseudospectra.jl:46


@dataclass
class ComputeThresholds:  # src/compute.jl:20-27
    minlancs4psa: int = 55
    maxstdqr4hess: int = 200
    minnev: int = 20
    maxit_lancs: int = 99


_default_thresholds = ComputeThresholds()
_ctx_cache: dict = {}


def _context(ngpus: int) -> PsaGpu:
    ctx = _ctx_cache.get(ngpus)
    if ctx is None:
        ctx = _ctx_cache[ngpus] = PsaGpu(ngpus)
    return ctx


def _matrix_key(T: np.ndarray, thresholds, upper_wrapper):
    """Cheap identity of a matrix for the zoom cache: same buffer, shape and a sampled checksum (diagonal, first row,
    last column).  A zoom re-enters psa_compute with the *same* Schur factor (src/zooming.jl:24-99 -> driver!), so
    the upload and the packing are skipped when this key is unchanged."""
    m, n = T.shape
    k = min(m, n)
    probe = complex(np.sum(T[np.arange(k), np.arange(k)])) + complex(np.sum(T[0, :])) + complex(np.sum(T[:, n - 1]))
    return (T.__array_interface__["data"][0], T.shape, T.strides, str(T.dtype), probe,
            thresholds.minlancs4psa, thresholds.maxstdqr4hess, bool(upper_wrapper))


def psacore(T, S, q0, x, y, bw, *, tol=1e-5, threaded=False, nt=None, warned=None, thresholds=_default_thresholds,
            ngpus=1, ctx: PsaGpu | None = None, want_iters=False, cancel=None, resident=False, row_batch=None,
            upper_wrapper=False, keep_on_device=False, hess_via_schur=False):
    """psacore(T,S,q,x,y,bw;tol) -> (Z, algo, warned)      [src/compute.jl:448-617]

    T: m x n (upper-triangular square, or (n+1) x n Hessenberg with bw == 2); S must be None / identity
    (generalised problems are not on the GPU path); q0: unit start vector (length >= n) -- or, with `row_batch`,
    one vector per row batch; x, y: grid lines.  `threaded`/`nt` are accepted for signature parity and ignored.
    `warned` carries the two once-only warning flags of compute.jl:603-610.  The matrix stays resident in `ctx`:
    a second call with the same array (a zoom) skips the upload; resident=True forces the reuse.

    hess_via_schur=True is the fast path for the :HessQR branch.  The reference's Givens sweep runs jj = 1..n-1 and
    then keeps `triu(T1[1:n,1:n])` (compute.jl:581-592), so R = Q'(Hs - zI) with Hs = H[1:n,:] the SQUARE part and Q
    unitary: R^-1 R^-H == (Hs - zI)^-1 (Hs - zI)^-H, the same operator for a matrix shared by all shifts.  One host
    `schur(Hs)` (n = 200: milliseconds) and q0' = U'q0 move the whole grid onto the tensor-core square path; in
    exact arithmetic the Lanczos coefficients are identical.  The per-shift Givens kernel stays the default (it is the
    literal restatement); results agree to rounding (tests, profiles/)."""
    if S is not None and not (np.isscalar(S) and S == 1):
        raise NotImplementedError("generalised problems (S != I) stay on the host path")
    warned = list(warned) if warned is not None else [False, False]
    T = np.asarray(T)
    m, n = T.shape
    if m < n:
        raise ValueError("Matrix size must be m x n with m >= n")  # compute.jl:459-461
    if m != n and n >= thresholds.minlancs4psa and bw != 2:
        # bw == 2 is the caller's promise of Hessenberg structure (compute.jl:579); anything wider is `qr(T1)` on a
        # dense rectangular matrix, which stays on the host
        raise PsaGpuUnsupported(PSA_ERR_UNSUPPORTED, "psacore", "rectangular input with bw != 2")
    if ctx is None:
        ctx = _context(ngpus)
    hess_fast = hess_via_schur and m == n + 1 and bw == 2 and m > thresholds.maxstdqr4hess and n >= thresholds.minlancs4psa
    if hess_fast:
        cache = getattr(ctx, "_hess_schur", None)
        hkey = _matrix_key(T, thresholds, False)
        if cache is None or cache[0] != hkey:
            import scipy.linalg as sla
            Ts, U = sla.schur(np.asarray(T[:n, :], dtype=np.complex128), output="complex")
            cache = ctx._hess_schur = (hkey, np.asfortranarray(np.triu(Ts)), U)
        _, T, U = cache
        q0 = np.asarray(q0)[..., :n] @ U.conj()  # rows are start vectors: (U' q)^T = q^T conj(U)
        m = n
    key = _matrix_key(T, thresholds, upper_wrapper)
    if getattr(ctx, "_matrix_key", None) == key and ctx.shape != (m, n) and m == n and not resident:
        ctx.project(np.arange(n), want_T=False)  # same matrix, but a projection of it is current: restore the full one
    if not ((resident and ctx.shape == (m, n)) or getattr(ctx, "_matrix_key", None) == key):
        ctx._matrix_key = None
        ctx.set_thresholds(thresholds.minlancs4psa, thresholds.maxstdqr4hess)
        ctx.set_matrix(T, upper_wrapper=upper_wrapper)
        ctx._matrix_key = key
    Z, iters, counts, algo = ctx.grid(x, y, q0, tol=tol, maxit=thresholds.maxit_lancs, bigsigma=0.1 * _FLOATMAX,
                                      want_iters=want_iters, cancel=cancel, row_batch=row_batch,
                                      want_Z=not keep_on_device)
    if counts[0] > 0 and not warned[0]:
        import warnings
        warnings.warn("Lanczos convergence failure(s) while computing resolvent norms")  # compute.jl:604
        warned[0] = True
    if counts[1] > 0 and not warned[1]:
        import warnings
        warnings.warn("Eigenvalue convergence failure(s) while computing resolvent norms")  # compute.jl:608
        warned[1] = True
    if hess_fast:
        algo = "HessQR"  # the branch the reference would have taken (test/medrect.jl:16 checks the symbol)
    if want_iters:
        return Z, algo, warned, iters
    return Z, algo, warned


def _get_step_size(n, ly):
    """Rows per psacore batch for Float64 (src/compute.jl:270-279)."""
    if n < 100:
        return max(1, ly // 8)
    return min(ly, max(1, (4 * ly) // n))


def row_batches(m, ly):
    """Batch index of every grid row in the reference's loop `for j=ly:-step:1` (compute.jl:199-206): rows
    j, j-1, ..., max(j-step+1, 1) share one start vector.  Returns (row_batch[ly] (0-based rows), nbatch)."""
    step = _get_step_size(m, ly)
    rb = (ly - 1 - np.arange(ly)) // step
    return rb.astype(np.int32), int(rb.max()) + 1 if ly else 0


# ---- grid / post-processing host logic (cheap; stays on the host exactly as in the reference) ----------

def _range(a, b, n):
    return np.linspace(float(a), float(b), int(n))


def shift_axes(ax, npts):
    """src/utils.jl:112-141"""
    a3, a4 = float(ax[2]), float(ax[3])
    if a4 > -a3:
        tpts = int(math.floor((a4 / (a4 - a3)) * npts))
        y = _range(a4 / (2 * (tpts - 1) + 1), a4, tpts)
        cnt = int(np.count_nonzero(y < -a3))
        if abs(y[min(len(y), cnt + 1) - 1] + a3) > 1e-15:
            y = np.concatenate([y[:cnt], [-a3], y[cnt:]])
        return y, cnt + 1
    if a4 < -a3:
        bpts = int(math.floor((a3 / (a3 - a4)) * npts))
        y = _range(a3, a3 / (2 * (bpts - 1) + 1), bpts)
        cnt = int(np.count_nonzero(y < -a4))
        if abs(y[max(1, cnt) - 1] + a4) > 1e-15:
            y = np.concatenate([y[:cnt], [-a4], y[cnt:]])
        return y, -(len(y) - cnt)
    if npts % 2 == 1:
        return _range(0.0, a4, int(round((npts + 1) / 2))), ((npts + 1) >> 1) - 1
    step = (a4 - a3) / npts
    return _range(step / 2, a4, int(round(npts / 2))), npts >> 1


def _grid(ax, npts, real_matrix, scale_equal):
    """src/compute.jl:122-145"""
    x_npts = y_npts = npts
    if scale_equal:
        y_dist, x_dist = ax[3] - ax[2], ax[1] - ax[0]
        if x_dist > y_dist:
            y_npts = max(5, int(math.ceil(y_dist / x_dist * npts)))
        else:
            x_npts = max(5, int(math.ceil(x_dist / y_dist * npts)))
    if real_matrix and ax[3] > 0 and ax[2] < 0:
        y, n_mirror = shift_axes(ax, y_npts)
    else:
        y, n_mirror = _range(ax[2], ax[3], y_npts), 0
    return _range(ax[0], ax[1], x_npts), np.asarray(y, dtype=np.float64), n_mirror


def _unmirror(Z, y, n_mirror):
    """src/compute.jl:222-235"""
    if n_mirror < 0:
        k = -n_mirror
        return np.vstack([Z, Z[Z.shape[0] - k:][::-1]]), np.concatenate([y, -y[len(y) - k:][::-1]])
    if y[0] != 0:
        return np.vstack([Z[:n_mirror][::-1], Z]), np.concatenate([-y[:n_mirror][::-1], y])
    return np.vstack([Z[1:n_mirror + 1][::-1], Z]), np.concatenate([-y[1:n_mirror + 1][::-1], y])


def recalc_levels(sigmin, ax, zstats=None):
    """src/utils.jl:21-73.  `zstats` = (min, max, min over values >= smallσ) when the device already reduced them
    (psa_gpu_postprocess); otherwise they are taken from `sigmin` like the reference does (utils.jl:24-31)."""
    if zstats is not None:
        smin, smax = float(zstats[0]), float(zstats[1])
    else:
        smin, smax = float(np.min(sigmin)), float(np.max(sigmin))
    if smax <= _SMALL_SIGMA:
        return np.zeros(0), -2
    if smin <= _SMALL_SIGMA:
        smin = float(zstats[2]) if zstats is not None else float(np.min(sigmin[sigmin >= _SMALL_SIGMA]))
    max_val, min_val = math.log10(smax), math.log10(smin)
    num_lines = 8
    last_lev = math.log10(0.03 * min(ax[3] - ax[2], ax[1] - ax[0]))
    if max_val >= last_lev:
        num_lines = int(math.ceil(num_lines * (last_lev - min_val) / (max_val - min_val)))
        max_val = last_lev
    if max_val < min_val:
        max_val = math.log10(smin) + 0.1 * math.log10(smax / smin)
        num_lines = 3
    levels = np.zeros(0)
    if num_lines > 0:
        rnd = lambda v: float(np.round(v))
        stepsize = max(rnd((max_val - min_val) / num_lines * 4) / 4, 0.25)
        if stepsize >= 0.75:
            stepsize = 1.0
        if (max_val - min_val) / stepsize < num_lines:
            stepsize = max(rnd((max_val - min_val) / num_lines * 10) / 10, 0.1)
            if stepsize == 0.3:
                stepsize = 0.2
            elif stepsize == 0.4 or 0.6 <= stepsize <= 0.8:
                stepsize = 0.5
            elif stepsize >= 0.9:
                stepsize = 1.0
        if (max_val - min_val) / stepsize < math.ceil(num_lines / 2):
            stepsize = (max_val - min_val) / num_lines
        else:
            min_val = math.ceil(min_val / stepsize) * stepsize
            max_val = math.floor(max_val / stepsize) * stepsize
        if stepsize > 0:
            levels = min_val + stepsize * np.arange(int(math.floor((max_val - min_val) / stepsize + 1e-12)) + 1)
        else:
            levels = np.array([min_val, max_val])
    stride = max(len(levels) // 9, 1)
    levels = levels[::-1][::stride][::-1]
    if len(levels) == 1:
        levels = levels * np.ones(2)
    return levels, 0


def _givens(f, g):
    """(c, s) of Julia's `givensAlgorithm(f, g)` (zlartg convention): [c s; -conj(s) c] [f; g] = [r; 0]."""
    if g == 0:
        return 1.0, 0.0 + 0.0j
    if f == 0:
        return 0.0, np.conj(g) / abs(g)
    af = abs(f)
    d = math.hypot(af, abs(g))
    return af / d, (f / af) * np.conj(g) / d


def _project_selection(m, factor, ax, eigA, thresholds=_default_thresholds):
    """Which eigenvalues the projection keeps (compute.jl:285-320): 0-based ascending positions, or None when every
    eigenvalue is selected (no projection)."""
    eigA = np.asarray(eigA)
    axis_w, axis_h = ax[1] - ax[0], ax[3] - ax[2]
    if factor >= 0:
        proj_w, proj_h = axis_w * factor, axis_h * factor
    else:
        proj_size = -1.0 / factor
    npick = 0
    ew_range = list(ax)
    selection = None
    if m > thresholds.minnev and eigA.size:
        while npick < thresholds.minnev:
            if factor >= 0:
                ew_range = [ew_range[0] - proj_w, ew_range[1] + proj_w, ew_range[2] - proj_h, ew_range[3] + proj_h]
                selection = np.flatnonzero((eigA.real > ew_range[0]) & (eigA.real < ew_range[1])
                                           & (eigA.imag > ew_range[2]) & (eigA.imag < ew_range[3]))
            else:
                selection = np.flatnonzero(np.abs(eigA) > proj_size)
                proj_size *= 0.5
            npick = len(selection)
            if factor == 0:
                break
    else:
        npick = m
    return None if npick == m else selection


def _maybe_project(Targ, factor, ax, eigA, thresholds=_default_thresholds, verbosity=0, ctx: PsaGpu | None = None):
    """Trefethen/Wright projection onto the invariant subspace of the eigenvalues near the window:
    `_maybe_project(Targ, factor, ax, eigA, thresholds, verbosity) -> (Tproj, eigAproj)`  [src/compute.jl:284-361].
    With the default proj_lev = Inf every eigenvalue is selected and T is returned untouched (compute.jl:322-328).
    The eigenvalue selection is host logic; the O(np n^2) Givens reordering runs on the device when `ctx` holds Targ
    (`psa_gpu_project`: the reference's swaps as a systolic wavefront), otherwise on the host exactly like the
    reference."""
    Targ = np.asarray(Targ)
    m = Targ.shape[0]
    eigA = np.asarray(eigA)
    if ctx is not None:
        selection = _project_selection(m, factor, ax, eigA, thresholds)
        if selection is None:
            ctx.project(np.arange(m), want_T=False)  # make sure the full matrix is the current one
            return Targ, eigA.copy()
        if verbosity > 1:
            print(f"projection reduces rank {m} -> {len(selection)}")
        if len(selection) == 0:
            return np.zeros((0, 0), dtype=np.complex128), eigA[selection]
        return ctx.project(selection), eigA[selection]
    selection = _project_selection(m, factor, ax, eigA, thresholds)
    if selection is None:
        return Targ, eigA.copy()
    npick = len(selection)
    if verbosity > 1:
        print(f"projection reduces rank {m} -> {npick}")
    eigAproj = eigA[selection]
    Tp = np.array(Targ, dtype=np.complex128)
    Tp = np.triu(Tp)  # UpperTriangular semantics: the strict lower part is not part of the matrix
    for i in range(npick):
        for k in range(selection[i] - 1, i - 1, -1):  # swap diagonal entries k, k+1 (0-based) by a Givens similarity
            c, sn = _givens(np.conj(Tp[k, k + 1]), np.conj(Tp[k, k] - Tp[k + 1, k + 1]))
            sn = -np.conj(sn)  # givens(f, g, k+1, k) with i1 > i2 flips the rotation (LinearAlgebra.givens)
            # rmul!(Tproj, adjoint(G)): columns k, k+1
            a1, a2 = Tp[:, k].copy(), Tp[:, k + 1].copy()
            Tp[:, k] = a1 * c + a2 * np.conj(sn)
            Tp[:, k + 1] = -a1 * sn + a2 * c
            # lmul!(G, Tproj): rows k, k+1
            r1, r2 = Tp[k, :].copy(), Tp[k + 1, :].copy()
            Tp[k, :] = c * r1 + sn * r2
            Tp[k + 1, :] = -np.conj(sn) * r1 + c * r2
    return np.asfortranarray(np.triu(Tp[:npick, :npick])), eigAproj


def psmode_inv_lanczos(A, q0, z, tol, num_its, *, fallbacksvd=False, ctx: PsaGpu | None = None, resident=False):
    """psmode_inv_lanczos(A,q,z,tol,num_its) -> Z,q          [src/modes.jl:185-274]

    Pseudomode at one point z: the same inverse-Lanczos recurrence as `_psa_lanczos!`, but the Lanczos basis is kept
    and the right singular vector for sigma_min(zI - A) is assembled from the Ritz vector (modes.jl:255-260).
    The whole recurrence, the basis and the final combination run on the GPU (`psa_gpu_pseudomode`); A must be upper
    triangular (the Schur factor, as `modeplot` passes it).  n < 200 uses a plain host SVD exactly like the reference
    (modes.jl:193-200).  Returns (nan, nan) when the iteration does not converge and fallbacksvd is False
    (modes.jl:262-272)."""
    A = np.asarray(A)
    m, n = A.shape
    if m != n:
        raise ValueError("Matrix must be square")  # modes.jl:188
    z = complex(z)

    def dense_svd():
        A1 = np.array(A, dtype=np.complex128)
        A1[np.arange(n), np.arange(n)] -= z
        U, S, Vh = np.linalg.svd(A1)
        return float(S[-1]), Vh[-1].conj()

    if n < 200:
        return dense_svd()
    if ctx is None:
        ctx = _context(1)
    key = _matrix_key(A, _default_thresholds, False)
    if not ((resident and ctx.shape == (n, n)) or getattr(ctx, "_matrix_key", None) == key):
        ctx._matrix_key = None
        ctx.set_thresholds(_default_thresholds.minlancs4psa, _default_thresholds.maxstdqr4hess)
        ctx.set_matrix(A)
        ctx._matrix_key = key
    Z, qv, l, info = ctx.pseudomode(z, q0, tol, num_its)
    if info != 0:  # `if l >= num_its` (modes.jl:262)
        if fallbacksvd:
            return dense_svd()
        return float("nan"), float("nan")
    return Z, qv


def tidyaxes(vmin, vmax, tol):
    """src/utils.jl:237-248"""
    round_to = (vmax - vmin) * tol
    expon = math.floor(math.log10(round_to))
    round_to = round(10 ** (math.log10(round_to) - expon)) * 10.0 ** expon
    return round_to * math.floor(vmin / round_to), round_to * math.ceil(vmax / round_to)


def vec2ax(dispvec):
    """Axis limits around the eigenvalues, as spectralportrait does (src/utils.jl:175-191, src/Pseudospectra.jl:729)."""
    v = np.asarray(dispvec)
    ca = [float(v.real.min()), float(v.real.max()), float(v.imag.min()), float(v.imag.max())]
    if ca[0] == ca[1]:
        ca[0], ca[1] = ca[0] - 0.5, ca[1] + 0.5
    if ca[2] == ca[3]:
        ca[2], ca[3] = ca[2] - 0.5, ca[3] + 0.5
    ext = max(ca[1] - ca[0], ca[3] - ca[2]) / 3
    ax = [ca[0] - ext, ca[1] + ext, ca[2] - ext, ca[3] + ext]
    ax[0], ax[1] = tidyaxes(ax[0], ax[1], 0.02)
    ax[2], ax[3] = tidyaxes(ax[2], ax[3], 0.02)
    return [float(a) for a in ax]


def psa_compute(Targ, npts, ax, eigA, opts=None, S=None, *, psatol=1e-5, thresholds=_default_thresholds,
                ctrlflag=None, q0=None, rng=None, ngpus=1, ctx: PsaGpu | None = None, per_batch_q=True,
                upper_wrapper=False):
    """psa_compute(T,npts,ax,eigA,opts,S=I) -> (Z,x,y,levels,info,Tproj,eigAproj,algo)   [src/compute.jl:87-268]

    Dense branch with the GPU option switched on.  `opts` keys honoured: levels, recompute_levels, real_matrix,
    scale_equal, proj_lev.  `ctrlflag` is a one-element int32 numpy array; setting it to 1 cancels and returns None
    like compute.jl:202-204.  Start vectors: the reference draws a new unit `q` of length n(Targ) for every batch of
    rows (compute.jl:199-208).  `q0` may be one vector (used for every batch), an (nbatch, n) array, or None: then
    they are drawn from `rng` in the reference's order, one per batch (per_batch_q=False: one for the whole grid).
    Un-mirroring, the clamp to ps_tiny and the min/max that feed `recalc_levels` run on the device
    (`psa_gpu_postprocess`); only the final array crosses PCIe."""
    opts = dict(opts or {})
    Targ = np.asarray(Targ)
    m, n = Targ.shape
    levels = opts.get("levels")
    re_calc = opts.get("recompute_levels", levels is None)
    if levels is None:
        levels = np.arange(-8, 0)
    elif np.ndim(levels) == 0 or len(levels) == 1:
        levels = np.ravel(levels)[0] * np.ones(2)
    ax = [float(a) for a in ax]
    real_matrix = opts.get("real_matrix", not np.iscomplexobj(Targ))
    x, y, n_mirror = _grid(ax, npts, real_matrix, opts.get("scale_equal", False))
    ly = len(y)
    eigAproj = np.array(eigA)
    Tproj = Targ
    if ctx is None:
        ctx = _context(ngpus)
    projected_on_device = False
    proj_lev = opts.get("proj_lev", math.inf)
    if m == n:  # compute.jl:147-153: projection only for square dense input
        on_device = (math.isfinite(proj_lev) and n >= thresholds.minlancs4psa and opts.get("gpu_project", True)
                     and _project_selection(m, proj_lev, ax, np.asarray(eigA), thresholds) is not None)
        if on_device:
            key = _matrix_key(Targ, thresholds, upper_wrapper)
            if getattr(ctx, "_matrix_key", None) != key:  # the FULL matrix stays resident across zooms
                ctx._matrix_key = None
                ctx.set_thresholds(thresholds.minlancs4psa, thresholds.maxstdqr4hess)
                ctx.set_matrix(Targ, upper_wrapper=upper_wrapper)
                ctx._matrix_key = key
            Tproj, eigAproj = _maybe_project(Targ, proj_lev, ax, np.asarray(eigA), thresholds, opts.get("verbosity", 1),
                                             ctx=ctx)
            projected_on_device = True
        else:
            Tproj, eigAproj = _maybe_project(Targ, proj_lev, ax, np.asarray(eigA), thresholds, opts.get("verbosity", 1))
        m = Tproj.shape[0]
    nproj = Tproj.shape[1]
    rb, nbatch = row_batches(m, ly)
    if q0 is None:
        rng = rng or np.random.default_rng()
        nq = nbatch if per_batch_q else 1
        q0 = np.empty((nq, n), dtype=np.complex128)
        for b in range(nq):  # `q = randn(n) + randn(n)*im; q = q / norm(q)` with n = size(Targ, 2) (compute.jl:207-208)
            v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
            q0[b] = v / np.linalg.norm(v)
    q0 = np.asarray(q0)
    batched = q0.ndim == 2 and q0.shape[0] > 1
    if q0.ndim == 2 and not batched:
        q0 = q0[0]
    # `q = q0[1:n]` (compute.jl:622): after a projection the vector is truncated, NOT renormalised
    q0 = q0[..., :nproj]
    try:
        _, algo, _ = psacore(Tproj, S, q0, x, y, m - nproj + 1, tol=psatol, thresholds=thresholds, ngpus=ngpus, ctx=ctx,
                             cancel=ctrlflag, row_batch=rb if batched else None, upper_wrapper=upper_wrapper,
                             keep_on_device=True, resident=projected_on_device)
    except PsaGpuCancelled:
        return None
    ps_tiny = 10 * math.sqrt(_FLOATMIN)  # compute.jl:236-238
    Z, y, zstats = ctx.postprocess(y, len(x), n_mirror, ps_tiny, _SMALL_SIGMA)
    err = 0
    if re_calc:
        levels, err = recalc_levels(Z, ax, zstats)
    elif np.min(levels) > math.log10(zstats[1]) or np.max(levels) < math.log10(zstats[0]):
        levels, err = recalc_levels(Z, ax, zstats)  # compute.jl:255-260
    return Z, x, y, levels, err, Tproj, eigAproj, algo
