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
NSTATS = 8
STAT_NAMES = ("solve_ms", "solve_launches", "total_launches", "grid_ms", "ticks", "flops", "bytes", "pack_ms")
EXPORTS = ("psa_gpu_init", "psa_gpu_set_matrix", "psa_gpu_set_matrix_device", "psa_gpu_grid", "psa_gpu_grid_device",
           "psa_gpu_apply_resolvent", "psa_gpu_last_stats", "psa_gpu_free", "psa_gpu_strerror", "psa_gpu_last_error",
           "psa_gpu_version")

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
    lib.psa_gpu_set_matrix_device.argtypes = [vp, vp, i64, i32, i32]
    lib.psa_gpu_grid.argtypes = [vp, vp, i32, vp, i32, vp, dbl, i32, dbl, vp, vp, vp, vp, vp]
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
    def set_matrix(self, T: np.ndarray):
        """Host matrix (m x n, any numeric dtype; complexified like compute.jl:450-454)."""
        T = np.asfortranarray(T, dtype=np.complex128)
        m, n = T.shape
        self._check(self._lib.psa_gpu_set_matrix(self._ctx, T.ctypes.data, m, m, n), "psa_gpu_set_matrix")
        self.shape = (m, n)

    def set_matrix_device(self, dptr: int, ldt: int, m: int, n: int):
        """Device-resident column-major complex128 matrix (raw device pointer, e.g. torch .data_ptr())."""
        self._check(self._lib.psa_gpu_set_matrix_device(self._ctx, dptr, ldt, m, n), "psa_gpu_set_matrix_device")
        self.shape = (m, n)

    def grid(self, x, y, q0, tol=1e-5, maxit=99, bigsigma=0.1 * np.finfo(np.float64).max, want_iters=True,
             cancel=None):
        """psacore replacement on host arrays.  Returns (Z[ly,lx], iters[ly,lx] | None, counts[4], algo)."""
        if self.shape is None:
            self._check(PSA_ERR_NOMATRIX, "psa_gpu_grid")
        n = self.shape[1]
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        q0 = np.ascontiguousarray(np.asarray(q0)[:n], dtype=np.complex128)
        if q0.shape[0] < n:
            raise ValueError("q0 shorter than n")
        lx, ly = len(x), len(y)
        Z = np.empty((ly, lx), order="F")
        it = np.empty((ly, lx), dtype=np.int32, order="F") if want_iters else None
        counts = np.zeros(NCOUNTS, dtype=np.int64)
        algo = C.c_int(0)
        cptr = cancel.ctypes.data if cancel is not None else None
        st = self._lib.psa_gpu_grid(self._ctx, x.ctypes.data, lx, y.ctypes.data, ly, q0.ctypes.data, tol, maxit,
                                    bigsigma, Z.ctypes.data, it.ctypes.data if want_iters else None,
                                    counts.ctypes.data, C.addressof(algo), cptr)
        self._check(st, "psa_gpu_grid")
        return Z, it, counts, ALGO_NAMES[algo.value]

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
