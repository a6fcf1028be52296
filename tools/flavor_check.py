"""Correctness + timing of the two solve-kernel choreographies (PSA_SOLVE_FLAVOR=shallow|deep) in one process:
bitwise equality of apply_resolvent and of a whole grid, then one solve launch vs number of active shifts and width."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
import psa_oracle as po, c_oracle as co

def flavor(f, w=None):
    if f.startswith("deep") and len(f) > 4:
        os.environ["PSA_DEEP_CW"] = f[4:]
        f = "deep"
    os.environ["PSA_SOLVE_FLAVOR"] = f
    if w: os.environ["PSA_PANEL_WIDTH"] = str(w)
    else: os.environ.pop("PSA_PANEL_WIDTH", None)

g = P.PsaGpu(1)
rng = np.random.default_rng(0)
ok = True
for n, ns in ((64, 1), (150, 70), (333, 130), (1000, 300)):
    T = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) + np.diag(3 * rng.standard_normal(n))
    z = 5 * (rng.standard_normal(ns) + 1j * rng.standard_normal(ns))
    Q = rng.standard_normal((ns, n)) + 1j * rng.standard_normal((ns, n))
    g.set_matrix(T)
    flavor("shallow"); V0 = g.apply_resolvent(z, Q)
    for w in (8, 16, 32):
        for f in ("deep4", "deep8"):
            flavor(f, w); V1 = g.apply_resolvent(z, Q)
            same = np.array_equal(V0, V1)
            ok &= same
            print(f"apply n={n} ns={ns} {f} width {w}: bitwise equal {same}" + ("" if same else f"  max rel diff {np.max(np.abs(V0 - V1)) / np.max(np.abs(V0)):.3e}"), flush=True)
T, eigA = po.complex_schur(po.readme_matrix(150))
x, y, _ = po.make_grid(po.vec2ax(eigA), 100, False)
q0 = po.start_vector(150, 0)
g.set_matrix(T)
for f in ("shallow", "deep4", "deep8"):
    flavor(f)
    g.grid(x, y, q0)
    t0 = time.time(); Z, it, c, _ = g.grid(x, y, q0); dt = time.time() - t0
    st = g.last_stats()
    print(f"C1 {f}: {dt*1e3:.2f} ms, solve {st['solve_ms']:.2f} ms over {st['ticks']:.0f} ticks", flush=True)
    if f == "shallow": Z0, it0 = Z, it
    else:
        same = np.array_equal(Z, Z0) and np.array_equal(it, it0); ok &= same
        print("C1 grid bitwise equal:", same)
Zc, itc, _ = co.grid(T, x, y, q0, nthreads=8)
print("C1 vs oracle max rel", float(np.max(np.abs(Z - Zc) / Zc)), "iters equal", float(np.mean(it == itc)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
T = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n) + np.diag(3 * rng.standard_normal(n))
g.set_matrix(T)
print("n", n)
NS_LIST = (18944, 9472, 8) if len(sys.argv) > 2 and sys.argv[2] == "quick" else (28416, 18944, 14208, 9472, 4736, 2368, 1184, 592, 148, 8)
for ns in NS_LIST:
    z = 5 * (rng.standard_normal(ns) + 1j * rng.standard_normal(ns))
    Q = rng.standard_normal((ns, n)) + 1j * rng.standard_normal((ns, n))
    ref = None
    for f in ("shallow", "deep4", "deep8"):
        for w in (32, 16, 8):
            if ns < 148 and w != 8: continue
            if f == "shallow" and ns not in (18944, 9472, 8): continue
            if f == "deep8" and ns > 4736: continue  # the sparse instantiation (one CTA per SM) is for tails
            flavor(f, w)
            V = g.apply_resolvent(z, Q)
            ms = []
            for _ in range(2):
                g.apply_resolvent(z, Q); ms.append(g.last_stats()["solve_ms"])
            if ref is None: ref = V
            same = np.array_equal(V, ref); ok &= same
            if not same: print(f"   max rel diff {np.max(np.abs(V - ref)) / np.max(np.abs(ref)):.3e}")
            print(f"ns={ns:6d} {f:8s} width={w:2d} panels={-(-ns//w):5d} {min(ms):8.2f} ms {8.0*n*n*ns/min(ms)*1e-9:6.2f} TF/s equal={same}", flush=True)
print("ALL BITWISE EQUAL" if ok else "MISMATCH")
