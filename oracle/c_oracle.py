"""ctypes binding of oracle/libpsa_oracle.so (the C restatement; TEST INFRASTRUCTURE ONLY).
Used by tests/ and by bench.py's cpu_baseline / --impl reference legs as the threaded CPU arm."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpsa_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.psa_oracle_grid.restype = C.c_int
        _LIB.psa_oracle_grid.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
    return _LIB


def grid(T, x, y, q0, tol=1e-5, maxit=99, bigsigma=0.1 * np.finfo(np.float64).max, nthreads=1):
    """Threaded psacore restatement (src/compute.jl:534-562 for m == n; 571-592 HessQR for m == n+1).
    Returns (Z[ly,lx], iters[ly,lx], flags[ly,lx]); flags bit0 = not converged, bit1 = eig failure."""
    T = np.asfortranarray(T, dtype=np.complex128)
    m, n = T.shape
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    q0 = np.ascontiguousarray(q0[:n], dtype=np.complex128)
    lx, ly = len(x), len(y)
    Z = np.zeros((ly, lx), order="F")
    it = np.zeros((ly, lx), dtype=np.int32, order="F")
    fl = np.zeros((ly, lx), dtype=np.int32, order="F")
    lib().psa_oracle_grid(m, n, T.ctypes.data, m, x.ctypes.data, lx, y.ctypes.data, ly, q0.ctypes.data, tol, maxit,
                          bigsigma, nthreads, Z.ctypes.data, it.ctypes.data, fl.ctypes.data)
    return Z, it, fl
