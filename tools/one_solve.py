import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
n = 4000; ns = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
rng = np.random.default_rng(0)
T = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n) + np.diag(3 * rng.standard_normal(n))
g = P.PsaGpu(1); g.set_matrix(T)
z = 5 * (rng.standard_normal(ns) + 1j * rng.standard_normal(ns))
Q = rng.standard_normal((ns, n)) + 1j * rng.standard_normal((ns, n))
for _ in range(2):
    g.apply_resolvent(z, Q); print(g.last_stats()["solve_ms"], flush=True)
