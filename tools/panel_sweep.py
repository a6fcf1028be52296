"""Solve-kernel time vs number of active shifts and panel width (n = 4000 random triangular matrix)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
WIDTHS = (32, 16, 8)
rng = np.random.default_rng(0)
T = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n) + np.diag(3 * rng.standard_normal(n))
g = P.PsaGpu(1)
g.set_matrix(T)
print("n", n)
for ns in (18944, 14208, 9472, 7104, 4736, 3552, 2368, 1184, 592, 148):
    z = 5 * (rng.standard_normal(ns) + 1j * rng.standard_normal(ns))
    Q = rng.standard_normal((ns, n)) + 1j * rng.standard_normal((ns, n))
    for w in WIDTHS:
        os.environ["PSA_PANEL_WIDTH"] = str(w)
        g.apply_resolvent(z, Q)
        ms = []
        for _ in range(2):
            g.apply_resolvent(z, Q)
            ms.append(g.last_stats()["solve_ms"])
        st = g.last_stats()
        print(f"ns={ns:6d} width={w:2d} panels={-(-ns//w):5d}  {min(ms):8.2f} ms  {st['flops']/min(ms)*1e-9:6.2f} TF/s", flush=True)
