""":HessQR with 256 < n <= 1024 (hess_solve_big_kernel<16|32>): algorithmic GB/s of the per-tick kernels against the
CTA-per-slot kernel (PSA_HESS_CTA=1), and agreement of the two.  usage: hess_big.py [n ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
import psa_oracle as po
g = P.PsaGpu(1)
rng = np.random.default_rng(3)
for n in [int(a) for a in sys.argv[1:]] or [300, 512, 1000]:
    H = np.triu(rng.standard_normal((n + 1, n)) + 1j * rng.standard_normal((n + 1, n)), -1) / np.sqrt(n)
    H[np.arange(1, n + 1), np.arange(n)] = 1.0
    npts = 64 if n <= 512 else 40
    x, y = np.linspace(-1.5, 1.5, npts), np.linspace(-1.5, 1.5, npts)
    q0 = po.start_vector(n, 0)
    g.set_matrix(np.asfortranarray(H))
    ref = None
    for mode in ("stream", "cta"):
        if mode == "cta": os.environ["PSA_HESS_CTA"] = "1"
        else: os.environ.pop("PSA_HESS_CTA", None)
        for _ in range(2):
            t0 = time.time(); Z, it, c, algo = g.grid(x, y, q0); dt = time.time() - t0
        st = g.last_stats()
        line = f"n={n} {mode:6s} {algo}: {dt*1e3:8.1f} ms, ticks {st['ticks']:.0f}, tick kernels {st['bytes']/st['solve_ms']*1e-6:6.0f} GB/s algorithmic, mean iters {c[2]/c[3]:.1f}"
        if ref is None: ref = Z
        else: line += f", max rel diff to stream {np.max(np.abs(Z - ref) / ref):.2e}"
        print(line, flush=True)
