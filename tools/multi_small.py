"""Smallest exercise of the multi-panel solve kernel (PSA_MULTI=2|3): bitwise check against the round-1 kernel.
usage: multi_small.py n ns"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
g = P.PsaGpu(1)
rng = np.random.default_rng(0)
cases = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(64, 1), (150, 70), (333, 130), (1000, 300)]
for n, ns in cases:
    T = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) + np.diag(3 * rng.standard_normal(n))
    z = 5 * (rng.standard_normal(ns) + 1j * rng.standard_normal(ns))
    Q = rng.standard_normal((ns, n)) + 1j * rng.standard_normal((ns, n))
    g.set_matrix(T)
    os.environ["PSA_SOLVE_FLAVOR"] = "shallow"; os.environ.pop("PSA_PANEL_WIDTH", None)
    V0 = g.apply_resolvent(z, Q)
    os.environ["PSA_SOLVE_FLAVOR"] = "deep"; os.environ["PSA_DEEP_CW"] = "4"; os.environ["PSA_PANEL_WIDTH"] = "32"; os.environ["PSA_CLUSTER"] = "1"
    V1 = g.apply_resolvent(z, Q)
    print(n, ns, "equal", np.array_equal(V0, V1), flush=True)
