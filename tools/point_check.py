"""Point-resident Lanczos kernel (csrc/psa_point.cuh) against the tick engine and the C oracle, with timings.

For each case the same grid is computed twice through the C ABI -- once with the point kernel (default for small n) and
once with PSA_NO_POINT_KERNEL=1 (the DMMA tick engine) -- and both are compared with the oracle on every point (small
grids) or a strided sample (large grids).  Writes gpurun_out/point_check.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
from pseudospectra_jl_b200 import compute as hc
import psa_oracle as po, c_oracle as co

g = P.PsaGpu(1)
out = {}


def timed_grid(x, y, q0, repeat, **kw):
    best, res, st = None, None, None
    for _ in range(repeat):
        t0 = time.perf_counter(); r = g.grid(x, y, q0, **kw); dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, res, st = dt, r, g.last_stats()
    return best, res, st


def case(name, T, x, y, q0, oracle_stride=1, repeat=5, **kw):
    g.set_matrix(T)
    os.environ.pop("PSA_NO_POINT_KERNEL", None)
    tp, (Zp, itp, cp, algo), stp = timed_grid(x, y, q0, repeat, **kw)
    os.environ["PSA_NO_POINT_KERNEL"] = "1"
    tt, (Zt, itt, ct, _), stt = timed_grid(x, y, q0, repeat, **kw)
    os.environ.pop("PSA_NO_POINT_KERNEL", None)
    floor = 1e-9 * np.linalg.norm(T)
    ok = Zt > floor
    r = {"algo": algo, "n": int(T.shape[0]), "points": int(Zp.size), "point_kernel_ms": tp * 1e3, "tick_engine_ms": tt * 1e3,
         "point_kernel_device_ms": stp["grid_ms"], "tick_engine_device_ms": stt["grid_ms"],
         "point_kernel_ticks": stp["ticks"], "tick_engine_ticks": stt["ticks"], "launches_point": stp["total_launches"],
         "launches_tick": stt["total_launches"], "tflops_point": stp["flops"] / (stp["grid_ms"] * 1e9),
         "tflops_tick": stt["flops"] / (stt["grid_ms"] * 1e9),
         "max_rel_point_vs_tick": float(np.max(np.abs(Zp[ok] - Zt[ok]) / Zt[ok])) if ok.any() else None,
         "iters_equal_point_vs_tick": float(np.mean(itp == itt)), "counts_point": [int(c) for c in cp],
         "counts_tick": [int(c) for c in ct], "mean_iters": float(cp[2] / max(cp[3], 1)), "max_iters": int(itp.max())}
    ys = np.arange(0, len(y), oracle_stride)
    xs = np.arange(0, len(x), oracle_stride)
    Zc, itc, flc = co.grid(T, x[xs], y[ys], q0, nthreads=8, **{k: v for k, v in kw.items() if k in ("maxit", "tol")})
    Zs, its = Zp[np.ix_(ys, xs)], itp[np.ix_(ys, xs)]
    okc = Zc > floor
    r.update({"oracle_points": int(Zc.size), "oracle_above_floor": int(okc.sum()),
              "max_rel_point_vs_oracle": float(np.max(np.abs(Zs[okc] - Zc[okc]) / Zc[okc])) if okc.any() else None,
              "iters_equal_point_vs_oracle": float(np.mean(its == itc)),
              "fallback_points_oracle": int(((flc & 2) != 0).sum()),
              "fallback_equal": bool(np.array_equal(Zs < po.PS_TINY, (flc & 2) != 0))})
    out[name] = r
    print(name, json.dumps(r), flush=True)


# C1: the README example, 100 x 100 (BASELINE configs[0]) -- every point against the oracle
T, eigA = po.complex_schur(po.readme_matrix(150))
x, y, nm = hc._grid(P.vec2ax(eigA), 100, False, False)
case("c1_readme150_100x100", T, x, y, po.start_vector(150, 0))
# the same matrix on a 400 x 400 grid: throughput regime
x4, y4, _ = hc._grid(P.vec2ax(eigA), 400, False, False)
case("readme150_400x400", T, x4, y4, po.start_vector(150, 0), oracle_stride=8, repeat=3)
# register-block boundaries of the kernel (NPL = 2..6) and the largest n that fits beside 8 warps' tridiagonals
for n in (55, 64, 65, 96, 97, 128, 129, 160, 161, 163):
    Tn, en = po.complex_schur(po.random_complex(n, n))
    xn, yn, _ = hc._grid(P.vec2ax(en), 24, False, False)
    case("random%d_24x24" % n, Tn, xn, yn, po.start_vector(n, 1), repeat=2)
# non-convergence path (maxit = 3) and a non-normal matrix with overflow points
Tn, en = po.complex_schur(po.random_complex(120, 3))
xn, yn, _ = hc._grid(P.vec2ax(en), 16, False, False)
case("random120_maxit3", Tn, xn, yn, po.start_vector(120, 0), repeat=1, maxit=3)
Tg, eg = po.complex_schur(po.grcar(120))
xg, yg, _ = hc._grid(P.vec2ax(eg), 40, False, False)
case("grcar120_40x40", Tg, xg, yg, po.start_vector(120, 0), repeat=2)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "point_check.json"), "w"), indent=1)
