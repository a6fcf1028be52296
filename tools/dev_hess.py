import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
import psa_oracle as po, c_oracle as co
np.set_printoptions(linewidth=200, precision=3)
A = po.convdiff_fd(6)
H = po.arnoldi_hessenberg(A, 208, seed=0)
ev = np.linalg.eigvals(H[:208, :])
x, y, _ = po.make_grid(po.vec2ax(ev), 14, False)
q0 = po.start_vector(208, 2)
g = P.PsaGpu(1)
g.set_matrix(H)
Z, it, counts, algo = g.grid(x, y, q0)
Zc, itc, flc = co.grid(np.asarray(H, dtype=complex), x, y, q0, nthreads=8)
print(np.abs(Z - Zc) / Zc)
print(it); print(itc); print(flc)
print(Zc)
rng = np.random.default_rng(0)
for n in (208, 209, 224, 230):
    Hd = np.triu(rng.standard_normal((n + 1, n)) + 1j * rng.standard_normal((n + 1, n)), -1)
    g.set_matrix(Hd)
    xs, ys = np.linspace(-3, 3, 5), np.linspace(-3, 3, 4)
    q = po.start_vector(n, 1)
    Z, it, counts, algo = g.grid(xs, ys, q)
    Zc, itc, flc = co.grid(Hd, xs, ys, q, nthreads=8)
    print(n, algo, np.max(np.abs(Z - Zc) / Zc), (it != itc).sum())
