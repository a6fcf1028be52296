"""Does a row shard computed on its own reproduce the rows of the full grid bit for bit (Z and iteration counts)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
pkg = __import__("psa_b200_loader").load_package()
n, npts = int(sys.argv[1]), int(sys.argv[2])
T, _, _ = bench.build_workload(n, npts)
ax, x, y = bench.grid_for(T, npts, pkg)
q0 = bench.start_vector(n)
g = pkg.PsaGpu(1); g.set_matrix(T)
Z, it, c, a = g.grid(x, y, q0)
print("full: mean iters", it.mean(), "counts", c)
for w in (2, 4):
    tot = 0
    for r in range(w):
        Zr, itr, cr, _ = g.grid(x, y[r::w], q0)
        same = np.array_equal(Zr, Z[r::w]) and np.array_equal(itr, it[r::w])
        tot += cr[2]
        print(f"world {w} rank {r}: rows {len(y[r::w])} bitwise equal {same} iters {cr[2]} expected {it[r::w].sum()} maxdiff {np.abs(Zr - Z[r::w]).max():.3e}")
    print(f"world {w}: mean iters {tot / it.size}")
