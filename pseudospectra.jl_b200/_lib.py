"""ctypes binding of libpsagpu.so (the C ABI in include/psa_gpu.h).

The library is the product; this module only marshals numpy arrays / raw device pointers into it.
There is no CPU fallback: if the shared library or a B200 is missing, every call raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSA_LIB_PATH") or os.path.join(_HERE, "libpsagpu.so")  # env override: experiments only

PSA_OK = 0
PSA_ERR_ARG, PSA_ERR_CUDA, PSA_ERR_CANCELLED, PSA_ERR_UNSUPPORTED = -1, -2, -3, -4
PSA_ERR_NODEVICE, PSA_ERR_NOMATRIX, PSA_ERR_NCCL, PSA_ERR_NOMEM = -5, -6, -7, -8
ALGO_NAMES = {0: None, 1: "sq_lanc", 2: "HessQR", 3: "rect_qr", 4: "SVD"}
NCOUNTS = 4
NSTATS = 9
STAT_NAMES = ("solve_ms", "solve_launches", "total_launches", "grid_ms", "ticks", "flops", "bytes", "pack_ms",
              "max_device_ms")
MATRIX_DEFAULT, MATRIX_UPPER_WRAPPER = 0, 1
EXPORTS = ("psa_gpu_init", "psa_gpu_set_matrix", "psa_gpu_set_matrix_ex", "psa_gpu_set_matrix_device",
           "psa_gpu_set_thresholds", "psa_gpu_project", "psa_gpu_grid", "psa_gpu_grid_batched", "psa_gpu_grid_device",
           "psa_gpu_postprocess", "psa_gpu_pseudomode", "psa_gpu_apply_resolvent", "psa_gpu_last_stats", "psa_gpu_free",
           "psa_gpu_strerror", "psa_gpu_last_error", "psa_gpu_version", "psa_gpu_plan")

_lib = None


class PsaGpuError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str = ""):
        self.status = status
        super().__init__(f"{what}: status {status} ({_strerror(status)}){(' -- ' + detail) if detail else ''}")


class PsaGpuUnsupported(PsaGpuError):
    """Shape not covered by the GPU path; the reference shim falls through to the host algorithm."""


class PsaGpuCancelled(PsaGpuError):
    pass


def load():
    """dlopen libpsagpu.so and declare prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not built: run `make` (or __graft_entry__.build()) first; "
                                "there is no CPU fallback for this path")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    lib.psa_gpu_init.argtypes = [i32, vp, C.POINTER(vp)]
    lib.psa_gpu_set_matrix.argtypes = [vp, vp, i64, i32, i32]
    lib.psa_gpu_set_matrix_ex.argtypes = [vp, vp, i64, i32, i32, i32]
    lib.psa_gpu_set_matrix_device.argtypes = [vp, vp, i64, i32, i32]
    lib.psa_gpu_set_thresholds.argtypes = [vp, i32, i32]
    lib.psa_gpu_project.argtypes = [vp, vp, i32, vp, i64]
    lib.psa_gpu_grid.argtypes = [vp, vp, i32, vp, i32, vp, dbl, i32, dbl, vp, vp, vp, vp, vp]
    lib.psa_gpu_grid_batched.argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, dbl, i32, dbl, vp, vp, vp, vp, vp]
    lib.psa_gpu_postprocess.argtypes = [vp, vp, i32, i32, i32, dbl, dbl, vp, i64, vp, vp]
    lib.psa_gpu_pseudomode.argtypes = [vp, dbl, dbl, vp, dbl, i32, vp, vp, vp, vp]
    lib.psa_gpu_grid_device.argtypes = [vp, vp, i32, vp, i32, vp, dbl, i32, dbl, vp, vp, vp, vp, vp]
    lib.psa_gpu_apply_resolvent.argtypes = [vp, vp, i32, vp, vp]
    lib.psa_gpu_last_stats.argtypes = [vp, vp]
    lib.psa_gpu_free.argtypes = [vp]
    for name in EXPORTS:
        getattr(lib, name).restype = i32
    lib.psa_gpu_strerror.argtypes = [i32]
    lib.psa_gpu_strerror.restype = C.c_char_p
    lib.psa_gpu_last_error.argtypes = [vp]
    lib.psa_gpu_last_error.restype = C.c_char_p
    lib.psa_gpu_version.argtypes = []
    lib.psa_gpu_version.restype = C.c_char_p
    lib.psa_gpu_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.psa_gpu_plan.restype = C.c_int
    _lib = lib
    return lib


def _strerror(status: int) -> str:
    try:
        return load().psa_gpu_strerror(status).decode()
    except Exception:  # library missing: still produce a message
        return "?"


class PsaGpu:
    """A libpsagpu context: device-resident matrix + slot buffers on `ngpus` devices of this process."""

    def __init__(self, ngpus: int = 1, device_ids=None):
        self._lib = load()
        self._ctx = C.c_void_p()
        ids = None
        if device_ids is not None:
            ids = (C.c_int * len(device_ids))(*device_ids)
            ngpus = len(device_ids)
        st = self._lib.psa_gpu_init(ngpus, ids, C.byref(self._ctx))
        if st != PSA_OK:
            self._ctx = C.c_void_p()
            raise PsaGpuError(st, "psa_gpu_init")
        self.ngpus = ngpus
        self.shape = None

    # -- helpers ----------------------------------------------------------------------------------
    def _check(self, st: int, what: str):
        if st == PSA_OK:
            return
        detail = self._lib.psa_gpu_last_error(self._ctx).decode() if self._ctx else ""
        if st == PSA_ERR_UNSUPPORTED:
            raise PsaGpuUnsupported(st, what, detail)
        if st == PSA_ERR_CANCELLED:
            raise PsaGpuCancelled(st, what, detail)
        raise PsaGpuError(st, what, detail)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.psa_gpu_free(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- API ----------------------------------------------------------------------------------------
    def set_matrix(self, T: np.ndarray, upper_wrapper: bool = False):
        """Host matrix (m x n, any numeric dtype; complexified like compute.jl:450-454).  `upper_wrapper=True` says T is
        the parent array of an `UpperTriangular` wrapper (strict lower part ignored without looking); otherwise a
        square T that is not upper triangular is PsaGpuUnsupported (the reference factorises it, compute.jl:597-601)."""
        T = np.asfortranarray(T, dtype=np.complex128)
        m, n = T.shape
        self.shape = None
        self._check(self._lib.psa_gpu_set_matrix_ex(self._ctx, T.ctypes.data, m, m, n,
                                                    MATRIX_UPPER_WRAPPER if upper_wrapper else MATRIX_DEFAULT),
                    "psa_gpu_set_matrix_ex")
        self.shape = (m, n)

    def set_thresholds(self, minlancs4psa: int = 55, maxstdqr4hess: int = 200):
        """The reference's mutable ComputeThresholds (compute.jl:18-30); takes effect at the next set_matrix."""
        self._check(self._lib.psa_gpu_set_thresholds(self._ctx, minlancs4psa, maxstdqr4hess), "psa_gpu_set_thresholds")

    def set_matrix_device(self, dptr: int, ldt: int, m: int, n: int):
        """Device-resident column-major complex128 matrix (raw device pointer, e.g. torch .data_ptr())."""
        self.shape = None
        self._check(self._lib.psa_gpu_set_matrix_device(self._ctx, dptr, ldt, m, n), "psa_gpu_set_matrix_device")
        self.shape = (m, n)

    def project(self, selection, want_T: bool = True):
        """Device-side Schur reordering + truncation (compute.jl:339-357): move the eigenvalues at the 0-based ascending
        diagonal positions `selection` to the top, keep the leading len(selection) x len(selection) block as the
        current matrix.  Returns Tproj (host copy) or None.  The full matrix stays resident for the next projection."""
        sel = np.ascontiguousarray(selection, dtype=np.int32)
        k = len(sel)
        Tp = np.empty((k, k), dtype=np.complex128, order="F") if want_T else None
        self._check(self._lib.psa_gpu_project(self._ctx, sel.ctypes.data, k, Tp.ctypes.data if want_T else None, k),
                    "psa_gpu_project")
        self.shape = (k, k)
        return Tp

    def grid(self, x, y, q0, tol=1e-5, maxit=99, bigsigma=0.1 * np.finfo(np.float64).max, want_iters=True,
             cancel=None, row_batch=None, want_Z=True):
        """psacore replacement on host arrays.  Returns (Z[ly,lx] | None, iters[ly,lx] | None, counts[4], algo).
        q0: one start vector (length >= n), or -- with `row_batch` (one int per grid row) -- a (nbatch, >= n) array of
        per-row-batch start vectors as the reference draws them (compute.jl:199-208).  want_Z=False keeps the result
        on the device for `postprocess`."""
        if self.shape is None:
            self._check(PSA_ERR_NOMATRIX, "psa_gpu_grid")
        n = self.shape[1]
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        q0 = np.asarray(q0)
        if q0.shape[-1] < n:
            raise ValueError("q0 shorter than n")
        lx, ly = len(x), len(y)
        Z = np.empty((ly, lx), order="F") if want_Z else None
        it = np.empty((ly, lx), dtype=np.int32, order="F") if want_iters else None
        counts = np.zeros(NCOUNTS, dtype=np.int64)
        algo = C.c_int(0)
        cptr = cancel.ctypes.data if cancel is not None else None
        if row_batch is None:
            q0 = np.ascontiguousarray(q0[:n], dtype=np.complex128)
            st = self._lib.psa_gpu_grid(self._ctx, x.ctypes.data, lx, y.ctypes.data, ly, q0.ctypes.data, tol, maxit,
                                        bigsigma, Z.ctypes.data if want_Z else None, it.ctypes.data if want_iters else None,
                                        counts.ctypes.data, C.addressof(algo), cptr)
        else:
            q0 = np.ascontiguousarray(np.atleast_2d(q0)[:, :n], dtype=np.complex128)
            rb = np.ascontiguousarray(row_batch, dtype=np.int32)
            if rb.shape != (ly,):
                raise ValueError("row_batch needs one entry per grid row")
            st = self._lib.psa_gpu_grid_batched(self._ctx, x.ctypes.data, lx, y.ctypes.data, ly, q0.ctypes.data,
                                                q0.shape[0], rb.ctypes.data, tol, maxit, bigsigma,
                                                Z.ctypes.data if want_Z else None, it.ctypes.data if want_iters else None,
                                                counts.ctypes.data, C.addressof(algo), cptr)
        self._check(st, "psa_gpu_grid")
        return Z, it, counts, ALGO_NAMES[algo.value]

    def postprocess(self, y, lx, n_mirror, ps_tiny, small_sigma=1e-150):
        """Device-side un-mirror + clamp + level statistics of the last grid result (compute.jl:222-238,
        utils.jl:24-31).  Returns (Zfull[lyf,lx], yfull[lyf], (min, max, min over >= small_sigma))."""
        y = np.ascontiguousarray(y, dtype=np.float64)
        ly = len(y)
        lyf = ly + abs(int(n_mirror))
        Z = np.empty((lyf, lx), order="F")
        yf = np.empty(lyf)
        zs = np.zeros(3)
        st = self._lib.psa_gpu_postprocess(self._ctx, y.ctypes.data, ly, lx, int(n_mirror), ps_tiny, small_sigma,
                                           Z.ctypes.data, lyf, yf.ctypes.data, zs.ctypes.data)
        self._check(st, "psa_gpu_postprocess")
        return Z, yf, tuple(zs.tolist())

    def pseudomode(self, z, q0, tol, num_its):
        """psmode_inv_lanczos on the resident square matrix (modes.jl:185-274).  Returns (Z, q, iters, info)."""
        if self.shape is None:
            self._check(PSA_ERR_NOMATRIX, "psa_gpu_pseudomode")
        n = self.shape[1]
        q0 = np.ascontiguousarray(np.asarray(q0)[:n], dtype=np.complex128)
        mode = np.empty(n, dtype=np.complex128)
        sig, it, info = C.c_double(0), C.c_int(0), C.c_int(0)
        z = complex(z)
        st = self._lib.psa_gpu_pseudomode(self._ctx, z.real, z.imag, q0.ctypes.data, tol, num_its, C.addressof(sig),
                                          mode.ctypes.data, C.addressof(it), C.addressof(info))
        self._check(st, "psa_gpu_pseudomode")
        return sig.value, mode, it.value, info.value

    def grid_device(self, d_x: int, lx: int, d_y: int, ly: int, d_q0: int, d_Z: int, d_iters: int = 0, tol=1e-5,
                    maxit=99, bigsigma=0.1 * np.finfo(np.float64).max):
        """Same computation on raw device pointers (single-device context).  Returns (counts, algo)."""
        counts = np.zeros(NCOUNTS, dtype=np.int64)
        algo = C.c_int(0)
        st = self._lib.psa_gpu_grid_device(self._ctx, d_x, lx, d_y, ly, d_q0, tol, maxit, bigsigma, d_Z,
                                           d_iters or None, counts.ctypes.data, C.addressof(algo), None)
        self._check(st, "psa_gpu_grid_device")
        return counts, ALGO_NAMES[algo.value]

    def apply_resolvent(self, z, Q):
        """V[s] = (T - z_s)^-1 (T - z_s)^-H Q[s]  (diagnostic hook; Q is (ns, n))."""
        z = np.ascontiguousarray(z, dtype=np.complex128)
        Q = np.ascontiguousarray(Q, dtype=np.complex128)
        V = np.empty_like(Q)
        self._check(self._lib.psa_gpu_apply_resolvent(self._ctx, z.ctypes.data, len(z), Q.ctypes.data, V.ctypes.data),
                    "psa_gpu_apply_resolvent")
        return V

    def last_stats(self) -> dict:
        s = np.zeros(NSTATS)
        self._check(self._lib.psa_gpu_last_stats(self._ctx, s.ctypes.data), "psa_gpu_last_stats")
        return dict(zip(STAT_NAMES, s.tolist()))


def version() -> str:
    return load().psa_gpu_version().decode()


def plan(m: int, n: int, maxit: int = 99):
    """(algo name | None, grid points in flight per SM in the point-resident kernel or 0): which machine a grid call
    would use for an m x n matrix -- host logic only, works without a GPU."""
    w = C.c_int(0)
    algo = load().psa_gpu_plan(int(m), int(n), int(maxit), C.byref(w))
    return ALGO_NAMES[algo], w.value
