"""One grid of the README matrix (n = 150) through the point-resident kernel: the ncu target for csrc/psa_point.cuh.
usage: python tools/point_one.py [npts=400] [repeat=2]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
from pseudospectra_jl_b200 import compute as hc
import psa_oracle as po

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 400
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T, eigA = po.complex_schur(po.readme_matrix(150))
x, y, _ = hc._grid(P.vec2ax(eigA), npts, False, False)
g = P.PsaGpu(1)
g.set_matrix(T)
for _ in range(rep):
    t0 = time.perf_counter(); Z, it, c, algo = g.grid(x, y, po.start_vector(150, 0)); dt = time.perf_counter() - t0
    st = g.last_stats()
    print(f"{algo} {Z.size} points {dt * 1e3:.2f} ms wall, {st['grid_ms']:.2f} ms device, {st['flops'] / st['grid_ms'] / 1e9:.2f} TFLOP/s, "
          f"mean iters {c[2] / c[3]:.2f}, ticks {st['ticks']:.0f}", flush=True)
