"""Longest-job-first queue order vs natural order on several matrix classes (seconds per grid, same results)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
from pseudospectra_jl_b200 import compute as hc
import psa_oracle as po
g = P.PsaGpu(1)
cases = []
T, e = po.complex_schur(po.toeplitz_from_symbol([0.0, 1.0], [0.0, 0.0, 0.25], 2000)); ax = P.vec2ax(e)
cases.append(("toeplitz2000 full", T, ax, True))
cases.append(("toeplitz2000 zoom", T, [ax[0] + 0.25 * (ax[1] - ax[0]), ax[0] + 0.5 * (ax[1] - ax[0]), 0.25 * ax[3], 0.5 * ax[3]], False))
T2, e2 = po.complex_schur(po.grcar(1000)); cases.append(("grcar1000", T2, P.vec2ax(e2), True))
T3, e3 = po.complex_schur(po.random_complex(1500, 2)); cases.append(("random1500", T3, P.vec2ax(e3), False))
cases.append(("random1500 zoom", T3, [-10.0, 10.0, -10.0, 10.0], False))
for name, T, ax, real in cases:
    x, y, nm = hc._grid(ax, 256, real, False)
    q0 = po.start_vector(T.shape[0], 0)
    g.set_matrix(T)
    res = {}
    for mode in ("lpt", "natural"):
        if mode == "natural": os.environ["PSA_NATURAL_ORDER"] = "1"
        else: os.environ.pop("PSA_NATURAL_ORDER", None)
        g.grid(x, y, q0, want_iters=False)
        t0 = time.time(); Z, it, c, a = g.grid(x, y, q0); dt = time.time() - t0
        res[mode] = (dt, g.last_stats()["ticks"], Z)
    same = np.array_equal(res["lpt"][2], res["natural"][2])
    print(f"{name:20s} points {len(x)*len(y):6d} mean iters {c[2]/c[3]:5.1f} max {it.max():3d} | lpt {res['lpt'][0]:.3f}s ({res['lpt'][1]:.0f} ticks) "
          f"natural {res['natural'][0]:.3f}s ({res['natural'][1]:.0f} ticks) bitwise equal {same}", flush=True)
