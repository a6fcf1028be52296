"""Device-side projection (psa_gpu_project, the Givens reordering of _maybe_project, compute.jl:339-357) at scale:
time on the device vs the sequential host restatement, plus invariants of the result.

    python tools/project_bench.py [n] [fraction of eigenvalues kept]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
from pseudospectra_jl_b200 import compute as hc
import psa_oracle as po

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
rng = np.random.default_rng(3)
# synthetic Schur factor with Ginibre statistics (no O(n^3) host reduction needed for a timing tool)
T = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)), 1)
ev = np.sqrt(2 * n) * np.sqrt(rng.random(n)) * np.exp(2j * np.pi * rng.random(n))
T[np.arange(n), np.arange(n)] = ev
T = np.asfortranarray(T)
# window: the eigenvalues inside a centred square holding about `frac` of them
half = np.sqrt(2 * n) * np.sqrt(frac * np.pi) / 2
sel = np.flatnonzero((np.abs(ev.real) < half) & (np.abs(ev.imag) < half))
print(f"n {n}, selected {len(sel)} eigenvalues, swaps {int(np.sum(sel - np.arange(len(sel))))}", flush=True)
g = P.PsaGpu(1)
g.set_matrix(T)
g.project(sel[:32])  # warm-up (allocations)
t0 = time.time(); Tp = g.project(sel); dt = time.time() - t0
print(f"device projection: {dt:.3f} s", flush=True)
assert np.abs(np.tril(Tp, -1)).max() == 0
assert np.allclose(np.diagonal(Tp), ev[sel], rtol=0, atol=1e-9 * np.abs(ev).max()), "eigenvalue order"
# unitary similarity restricted to an invariant subspace: singular values of the resolvent can only shrink
for z in (0.3 + 0.2j, -1.0 + 0.5j):
    a, b = po.sigmin_svd(Tp, z), po.sigmin_svd(T, z)
    assert a >= b * (1 - 1e-8), (a, b)
    print(f"  sigma_min at {z}: projected {a:.6e} >= full {b:.6e}")
if n <= 1200:
    class Th:  # force exactly this selection through the host restatement
        pass
    t0 = time.time()
    Tw = np.triu(np.array(T))
    for i in range(len(sel)):
        for k in range(sel[i] - 1, i - 1, -1):
            c, sn = hc._givens(np.conj(Tw[k, k + 1]), np.conj(Tw[k, k] - Tw[k + 1, k + 1]))
            sn = -np.conj(sn)
            a1, a2 = Tw[:, k].copy(), Tw[:, k + 1].copy()
            Tw[:, k] = a1 * c + a2 * np.conj(sn); Tw[:, k + 1] = -a1 * sn + a2 * c
            r1, r2 = Tw[k, :].copy(), Tw[k + 1, :].copy()
            Tw[k, :] = c * r1 + sn * r2; Tw[k + 1, :] = -np.conj(sn) * r1 + c * r2
    th = time.time() - t0
    Th_ = np.triu(Tw[:len(sel), :len(sel)])
    print(f"host restatement (numpy, sequential swaps): {th:.1f} s; |Tp - Th| / |Th| = {np.linalg.norm(Tp - Th_) / np.linalg.norm(Th_):.2e}")
x, y = np.linspace(-half / 2, half / 2, 24), np.linspace(-half / 2, half / 2, 20)
q0 = po.start_vector(n, 0)[:len(sel)]
t0 = time.time(); Z, it, c, algo = g.grid(x, y, q0); dg = time.time() - t0
print(f"grid on the projected matrix ({len(sel)} x {len(sel)}, {algo}): {dg*1e3:.1f} ms, mean iters {c[2]/c[3]:.1f}")
g.project(np.arange(n), want_T=False)  # npick == n restores the full matrix (it stayed resident)
t0 = time.time(); Zf, itf, cf, _ = g.grid(x, y, po.start_vector(n, 0)); df = time.time() - t0
print(f"grid on the full matrix again (no upload): {df*1e3:.1f} ms")
