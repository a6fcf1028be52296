"""CPU arm built the way BASELINE.md §4 describes the reference's Threads path: a queue of grid points consumed by one
worker per host core, each worker holding a PRIVATE copy of T whose diagonal it shifts in place
(`_Sq_Lancs_Server`, /root/reference/src/compute.jl:388-417, 534-562) and running `_psa_lanczos!` (619-665) with the
same library routines Julia dispatches to: OpenBLAS `ztrsv('U','C')` / `ztrsv('U','N')` for `F1 \\ (F1' \\ q)` and LAPACK
`dsyevr` for `eigvals(H[1:l,1:l])`, BLAS threads pinned to 1 per worker (the reference warns about oversubscription,
compute.jl:439-441).  TEST INFRASTRUCTURE ONLY (bench.py's cpu_baseline / --impl reference legs and tests/).

Workers are processes (fork), not threads: scipy's BLAS wrappers release the GIL only inside the call, and a process per
core is what gives every worker its own T copy anyway.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

import psa_oracle as po

_G = {}


def _init(T, x, y, q0, tol, maxit):
    try:
        from threadpoolctl import threadpool_limits
        _G["limit"] = threadpool_limits(limits=1)
    except Exception:  # pragma: no cover
        os.environ["OPENBLAS_NUM_THREADS"] = "1"
    _G["T1"] = np.array(T, dtype=np.complex128, order="F")  # `copy(Twork)` per server, compute.jl:537
    _G["diag"] = np.diagonal(T).copy()
    _G["x"], _G["y"], _G["q0"], _G["tol"], _G["maxit"] = x, y, q0, tol, maxit
    _G["idx"] = np.arange(T.shape[0])


def _point(p):
    j, k = p
    T1 = _G["T1"]
    T1[_G["idx"], _G["idx"]] = _G["diag"] - (_G["x"][k] + 1j * _G["y"][j])  # compute.jl:405
    sig, conv, fail, it = po.psa_lanczos(T1, _G["q0"], _G["tol"], _G["maxit"], po.BIGSIGMA)
    with np.errstate(all="ignore"):
        return j, k, 1.0 / np.sqrt(sig), it, (0 if conv else 1) | (2 if fail else 0)


def grid(T, x, y, q0, tol=1e-5, maxit=99, nworkers=None):
    """Square (:sq_lanc) grid with one forked worker per core.  Returns (Z[ly,lx], iters, flags, seconds) where seconds
    excludes process start-up and the per-worker copy of T (the reference re-creates its servers per row batch, but that
    is host set-up, not the path being timed)."""
    nworkers = nworkers or os.cpu_count() or 1
    T = np.asfortranarray(T, dtype=np.complex128)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    q0 = np.ascontiguousarray(q0[:T.shape[1]], dtype=np.complex128)
    pts = [(j, k) for j in range(len(y)) for k in range(len(x))]
    Z = np.zeros((len(y), len(x)))
    it = np.zeros((len(y), len(x)), dtype=np.int32)
    fl = np.zeros((len(y), len(x)), dtype=np.int32)
    ctx = mp.get_context("fork")
    with ctx.Pool(nworkers, initializer=_init, initargs=(T, x, y, q0, tol, maxit)) as pool:
        pool.map(_noop, range(nworkers * 4))  # all workers forked and initialised before the clock starts
        t0 = time.time()
        for j, k, z, i, f in pool.imap_unordered(_point, pts, chunksize=1):
            Z[j, k], it[j, k], fl[j, k] = z, i, f
        dt = time.time() - t0
    return Z, it, fl, dt


def _noop(_):
    time.sleep(0.01)
    return 0
