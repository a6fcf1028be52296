"""Correctness (bitwise vs the single-CTA kernels) and latency of the cluster-pipelined solve kernel for sparse launches.
PSA_CLUSTER caps the cluster size (1 = off, 2, 4)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
g = P.PsaGpu(1)
rng = np.random.default_rng(0)
ok = True
for n, ns in ((64, 1), (150, 9), (333, 20), (1000, 37), (4000, 8), (4000, 290)):
    T = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n) + np.diag(3 * rng.standard_normal(n))
    z = 5 * (rng.standard_normal(ns) + 1j * rng.standard_normal(ns))
    Q = rng.standard_normal((ns, n)) + 1j * rng.standard_normal((ns, n))
    g.set_matrix(T)
    os.environ["PSA_SOLVE_FLAVOR"] = "shallow"; os.environ.pop("PSA_PANEL_WIDTH", None)
    V0 = g.apply_resolvent(z, Q)
    os.environ["PSA_SOLVE_FLAVOR"] = "deep"
    for w in (8, 16, 32):
        os.environ["PSA_PANEL_WIDTH"] = str(w)
        for cl in (1, 2, 4):
            os.environ["PSA_CLUSTER"] = str(cl)
            V = g.apply_resolvent(z, Q)
            ms = []
            for _ in range(2):
                g.apply_resolvent(z, Q); ms.append(g.last_stats()["solve_ms"])
            same = bool(np.array_equal(V, V0)); ok &= same
            print(f"n={n:5d} ns={ns:4d} width={w:2d} cluster<={cl}: {min(ms):8.3f} ms  bitwise equal {same}", flush=True)
print("ALL BITWISE EQUAL" if ok else "MISMATCH")
