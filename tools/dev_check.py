"""Developer GPU check (not a test): solve kernel vs numpy, grid vs oracle, quick timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import scipy.linalg as sla
import psa_b200_loader
P = psa_b200_loader.load_package()
import psa_oracle as po, c_oracle as co

def check_apply(n, ns, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    T = np.triu(A) + np.diag(rng.standard_normal(n) * 3)
    z = rng.standard_normal(ns) * 5 + 1j * rng.standard_normal(ns) * 5
    Q = rng.standard_normal((ns, n)) + 1j * rng.standard_normal((ns, n))
    g = P.PsaGpu(1)
    g.set_matrix(T)
    V = g.apply_resolvent(z, Q)
    err = 0
    for s in range(ns):
        F = T - z[s] * np.eye(n)
        w = sla.solve_triangular(F, Q[s], trans='C')
        v = sla.solve_triangular(F, w)
        err = max(err, np.linalg.norm(V[s] - v) / np.linalg.norm(v))
    print(f"apply n={n} ns={ns}: max rel err {err:.2e}", flush=True)
    g.close()
    return err

for n, ns in [(64, 3), (100, 64), (150, 70), (200, 5), (333, 130)]:
    check_apply(n, ns)

# config 1 subsampled
B = po.readme_matrix(150)
T, eigA = po.complex_schur(B)
ax = po.vec2ax(eigA)
x, y, _ = po.make_grid(ax, 100, False)
q0 = po.start_vector(150, 0)
xs, ys = x[::3], y[::3]
g = P.PsaGpu(1)
g.set_matrix(T)
t = time.time(); Z, it, counts, algo = g.grid(xs, ys, q0); t1 = time.time() - t
Zc, itc, flc = co.grid(T, xs, ys, q0, nthreads=8)
print("config1 sub:", algo, "rel", np.max(np.abs(Z - Zc) / Zc), "iters mismatch", int((it != itc).sum()), "of", it.size,
      "counts", counts, f"{t1*1e3:.1f} ms", g.last_stats(), flush=True)
t = time.time(); Z, it, counts, algo = g.grid(x, y, q0); t1 = time.time() - t
print("config1 full 100x100:", f"{t1*1e3:.1f} ms", counts, g.last_stats(), flush=True)
g.close()

# HessQR
G = po.grcar(260)[:, :-1]
G = np.asfortranarray(G[:260, :259])
xs = np.linspace(-1, 3, 12); ys = np.linspace(-3, 3, 11)
q0 = po.start_vector(259, 0)
g = P.PsaGpu(1)
g.set_matrix(G)
Z, it, counts, algo = g.grid(xs, ys, q0)
Zc, itc, flc = co.grid(G, xs, ys, q0, nthreads=8)
ok = Zc > 1e-13
print("hessqr:", algo, "rel", np.max(np.abs(Z - Zc)[ok] / Zc[ok]), "iters mismatch", int((it != itc).sum()), "of", it.size, counts,
      g.last_stats(), flush=True)
g.close()

# mid-size timing
for n, npts in [(1000, 64), (2000, 96)]:
    A = po.random_complex(n, 1)
    t = time.time(); T, eigA = po.complex_schur(A); ts = time.time() - t
    ax = po.vec2ax(eigA)
    x, y, _ = po.make_grid(ax, npts, False)
    q0 = po.start_vector(n, 0)
    g = P.PsaGpu(1)
    t = time.time(); g.set_matrix(T); tm = time.time() - t
    t = time.time(); Z, it, counts, algo = g.grid(x, y, q0); t1 = time.time() - t
    st = g.last_stats()
    print(f"n={n} grid {npts}^2: schur {ts:.1f}s set_matrix {tm*1e3:.0f} ms grid {t1:.3f}s  pts/s {npts*npts/t1:.0f} mean iters {counts[2]/counts[3]:.1f} "
          f"solve TF/s {st['flops']/st['solve_ms']*1e-9:.2f} overall TF/s {st['flops']/(t1*1e3)*1e-9:.2f}", st, flush=True)
    # spot check vs oracle
    idx = [(0, 0), (npts // 2, npts // 2), (npts - 1, 3)]
    for (j, k) in idx:
        Zc, itc, _ = co.grid(T, x[k:k+1], y[j:j+1], q0)
        print("   spot", j, k, Z[j, k], Zc[0, 0], abs(Z[j, k] - Zc[0, 0]) / Zc[0, 0], it[j, k], itc[0, 0], flush=True)
    g.close()
