"""Generates the committed golden fixtures from the CPU oracle (oracle/psa_oracle.py) and dense SVD.

The reference (pure Julia) cannot run in the build container and ships no golden Z arrays, so these vectors
are produced by the oracle restatement and cross-checked against `svdvals(T - zI)[-1]` -- the quantity the
reference's own SVD branch returns (src/compute.jl:510-518).  Run:  python tests/golden/make_golden.py"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import psa_oracle as po


def case_readme():
    T, eigA = po.complex_schur(po.readme_matrix(150))
    ax = po.vec2ax(eigA)
    x, y, _ = po.make_grid(ax, 100, False)
    xs, ys = x[::9], y[::9]
    q0 = po.start_vector(150, 0)
    Z, algo, warned, it = po.psacore(T, q0, xs, ys, 1, want_iters=True)
    svd = np.array([[po.sigmin_svd(T, xx + 1j * yy) for xx in xs] for yy in ys])
    return dict(T=T, ax=np.array(ax), x=xs, y=ys, q0=q0, Z=Z, iters=it, svd=svd, algo=algo)


def case_grcar_real():
    T, eigA = po.complex_schur(po.grcar(120))
    ax = po.vec2ax(eigA)
    q0 = po.start_vector(120, 3)
    out = po.psa_compute(T, 24, ax, eigA, {"real_matrix": True}, q_for_batch=lambda b, n: q0)
    return dict(T=T, ax=np.array(ax), q0=q0, Z=out[0], x=out[1], y=out[2], levels=out[3], algo=out[7])


def case_hessqr():
    H = np.asfortranarray(po.grcar(210)[:, :-1]).astype(np.complex128)  # test/medrect.jl:8-11
    xs, ys = np.linspace(-1.0, 3.0, 9), np.linspace(-3.0, 3.0, 8)
    q0 = po.start_vector(209, 1)
    Z, algo, warned, it = po.psacore(H, q0, xs, ys, 2, want_iters=True)
    return dict(H=H, x=xs, y=ys, q0=q0, Z=Z, iters=it, algo=algo)


if __name__ == "__main__":
    for name, fn in [("readme150", case_readme), ("grcar120_real", case_grcar_real), ("hessqr210", case_hessqr)]:
        d = fn()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, {k: getattr(v, "shape", v) for k, v in d.items()})
