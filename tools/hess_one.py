"""A short :HessQR run (Arnoldi-projected 201x200 Hessenberg, BASELINE config 4 in miniature: 160x160 grid) -- the
command profiled with ncu for the Hessenberg kernels (dram bytes + duration of hess_qr_kernel / hess_solve_warp_kernel)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
import psa_oracle as po
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 160
H = po.arnoldi_hessenberg(po.convdiff_fd(50), 200, seed=0)
ev = np.linalg.eigvals(H[:200, :])
x, y, _ = po.make_grid(po.vec2ax(ev), npts, False)
g = P.PsaGpu(1)
g.set_matrix(np.asfortranarray(H, dtype=complex))
for _ in range(2):
    t0 = time.time(); Z, it, c, algo = g.grid(x, y, po.start_vector(200, 0)); dt = time.time() - t0
    st = g.last_stats()
    print(algo, f"{dt*1e3:.1f} ms, ticks {st['ticks']:.0f}, solve kernels {st['bytes']/st['solve_ms']*1e-6:.0f} GB/s algorithmic, mean iters {c[2]/c[3]:.1f}", flush=True)
