import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
import psa_oracle as po, c_oracle as co
np.set_printoptions(linewidth=220, precision=3)
T, eigA = po.complex_schur(po.toeplitz_from_symbol([0.0, 1.0], [0.0, 0.0, 0.25], 1024))
x = np.linspace(-0.2, 1.4, 12); y = np.array([0.31, 0.77, 1.3])
q0 = po.start_vector(1024, 0)
g = P.PsaGpu(1); g.set_matrix(T)
Zp, itp, cp, _ = g.grid(x, y, q0)
Zm, itm, cm, _ = g.grid(x, -y, q0)
Zc, itc, flc = co.grid(T, x, y, q0, nthreads=8)
print(Zp); print(np.abs(Zp - Zm) / Zp); print(np.abs(Zp - Zc) / Zc); print(itp); print(itm); print(cp, cm, np.linalg.norm(T))
