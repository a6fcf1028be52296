"""Runs the BASELINE.json configs through the C ABI on one B200: parity on sampled points against the C oracle
(and dense SVD where cheap), throughput with the same accounting as bench.py.  Writes gpurun_out/configs.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import psa_b200_loader
P = psa_b200_loader.load_package()
from pseudospectra_jl_b200 import compute as hc
import psa_oracle as po, c_oracle as co

which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
out = {}
g = P.PsaGpu(1)
HBM = 6552.3


def sample_check(T, x, y, q0, Z, it, nsamp=24, seed=0, svd=False, nsvd=None):
    """Oracle (and optionally dense-SVD) check on sampled points.  Half of the sample is drawn among the points whose
    GPU result is above the rounding floor n*eps*|T| (so that a window deep inside a non-normal spectrum, where almost
    every point sits on the floor, still gets a meaningful relative comparison), half uniformly."""
    rng = np.random.default_rng(seed)
    floor = 1e-9 * np.linalg.norm(T)
    cand = np.argwhere(Z > 10 * floor)
    pts = [tuple(p) for p in cand[rng.permutation(len(cand))[: nsamp // 2]]] if len(cand) else []
    while len(pts) < nsamp:
        pts.append((int(rng.integers(0, len(y))), int(rng.integers(0, len(x)))))
    rel, itm, svdrel = [], 0, []
    nsvd = nsamp if nsvd is None else nsvd
    for j, k in pts:
        Zc, itc, flc = co.grid(T, x[k:k + 1], y[j:j + 1], q0)
        if Zc[0, 0] > floor:
            rel.append(abs(Z[j, k] - Zc[0, 0]) / Zc[0, 0])
            if svd and len(svdrel) < nsvd:
                s = po.sigmin_svd(T, x[k] + 1j * y[j])
                svdrel.append((abs(Z[j, k] - s) / s, abs(Zc[0, 0] - s) / s))
        itm += int(it[j, k] != itc[0, 0])
    r = {"sampled": int(nsamp), "above_floor": len(rel), "max_rel_vs_oracle": float(max(rel)) if rel else None,
         "iter_mismatches": itm}
    if svdrel:
        r["max_rel_vs_svd_gpu"] = float(max(a for a, b in svdrel)); r["max_rel_vs_svd_oracle"] = float(max(b for a, b in svdrel))
    return r


def run(name, T, x, y, q0, n_mirror=0, nsamp=24, svd=False, repeat=2, nsvd=None):
    t0 = time.time(); g.set_matrix(T); t_set = time.time() - t0
    best = None
    for _ in range(repeat):
        t0 = time.time(); Z, it, counts, algo = g.grid(x, y, q0); dt = time.time() - t0
        st = g.last_stats()
        if best is None or dt < best[0]:
            best = (dt, st)
    dt, st = best
    pts = len(x) * len(y)
    r = {"algo": algo, "shape": list(T.shape), "grid_computed": [len(y), len(x)], "n_mirror": n_mirror, "seconds": dt,
         "points_per_s": pts / dt, "set_matrix_s": t_set, "mean_iters": float(counts[2] / counts[3]), "max_iters": int(it.max()),
         "nonconverged": int(counts[0]), "fallback_points": int(counts[1]), "ticks": st["ticks"],
         "solve_share": st["solve_ms"] / st["grid_ms"], "tflops_whole": st["flops"] / (dt * 1e12),
         "tflops_solve_kernel": st["flops"] / (st["solve_ms"] * 1e9)}
    if st["bytes"] > 0:
        r["gbs_solve_kernels"] = st["bytes"] / (st["solve_ms"] * 1e6); r["hbm_frac"] = r["gbs_solve_kernels"] / HBM
    r.update(sample_check(T, x, y, q0, Z, it, nsamp=nsamp, svd=svd, nsvd=nsvd))
    out[name] = r
    print(name, json.dumps(r), flush=True)
    return Z, it


if "c1" in which:  # README 150x150, npts=100
    T, eigA = po.complex_schur(po.readme_matrix(150))
    x, y, nm = hc._grid(P.vec2ax(eigA), 100, False, False)
    run("c1_readme150_100x100", T, x, y, po.start_vector(150, 0), nm, nsamp=60, svd=True)
if "c2" in which:  # Grcar 1000, 200x200, real -> mirror
    T, eigA = po.complex_schur(po.grcar(1000))
    ax = P.vec2ax(eigA)
    x, y, nm = hc._grid(ax, 200, True, False)
    print("c2 ax", ax, "rows computed", len(y), "mirror", nm)
    run("c2_grcar1000_200x200_mirror", T, x, y, po.start_vector(1000, 0), nm, nsamp=40)
if "c3" in which:  # the headline workload: dense random complex n=4000, 400x400 (same inputs as bench.py)
    import bench
    T, schur_s, how = bench.build_workload(4000, 400)
    ax, x, y = bench.grid_for(T, 400, P)
    print("c3 schur", how, schur_s, "s ax", ax)
    run("c3_random4000_400x400", T, x, y, bench.start_vector(4000), 0, nsamp=64, svd=True, repeat=1, nsvd=8)
if "c4" in which:  # convdiff_fd(50) projected by 200-step Arnoldi -> 201x200 Hessenberg, 500x500
    A = po.convdiff_fd(50)
    t0 = time.time(); H = po.arnoldi_hessenberg(A, 200, seed=0); print("arnoldi", time.time() - t0, "s")
    ev = np.linalg.eigvals(H[:200, :])
    x, y, nm = hc._grid(P.vec2ax(ev), 500, False, False)
    Hc = np.asfortranarray(H, dtype=complex)
    Zg, itg = run("c4_convdiff_arnoldi201x200_500x500", Hc, x, y, po.start_vector(200, 0), nm, nsamp=60)
    # fast path: one Schur factorisation of the square part, then the tensor-core square kernels (compute.py psacore)
    g._matrix_key = None
    best = None
    for _ in range(2):
        t0 = time.time()
        Zf, algo, _, itf = hc.psacore(Hc, None, po.start_vector(200, 0), x, y, 2, ctx=g, want_iters=True, hess_via_schur=True)
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    st = g.last_stats()
    floor = 1e-9 * np.linalg.norm(Hc)
    okf = Zg > floor
    out["c4_fast_path_schur"] = {"seconds": best, "points_per_s": len(x) * len(y) / best, "ticks": st["ticks"],
                                 "tflops_solve_kernel": st["flops"] / (st["solve_ms"] * 1e9),
                                 "max_rel_vs_givens_path": float(np.max(np.abs(Zf[okf] - Zg[okf]) / Zg[okf])),
                                 "iters_equal_frac": float(np.mean(itf == itg)), "points_above_floor": int(okf.sum())}
    print("c4_fast", json.dumps(out["c4_fast_path_schur"]), flush=True)
if "c5" in which:  # real Toeplitz n=8000 'triangle' symbol, 512x512 + zoom on the resident matrix
    n = 8000
    t0 = time.time(); T, eigA = po.complex_schur(po.toeplitz_from_symbol([0.0, 1.0], [0.0, 0.0, 0.25], n)); ts = time.time() - t0
    ax = P.vec2ax(eigA)
    x, y, nm = hc._grid(ax, 512, True, False)
    print("c5 schur", ts, "s ax", ax, "rows", len(y), "mirror", nm)
    q0 = po.start_vector(n, 0)
    run("c5_toeplitz8000_512x512_mirror", T, x, y, q0, nm, nsamp=6, repeat=1)
    # zoom: quarter window, same resident T (no set_matrix)
    axz = [ax[0] + 0.25 * (ax[1] - ax[0]), ax[0] + 0.5 * (ax[1] - ax[0]), 0.25 * ax[3], 0.5 * ax[3]]
    xz, yz, nmz = hc._grid(axz, 512, True, False)
    t0 = time.time(); Zz, itz, cz, _ = g.grid(xz, yz, q0); dz = time.time() - t0
    stz = g.last_stats()
    out["c5_zoom_recompute"] = {"ax": axz, "seconds": dz, "points_per_s": len(xz) * len(yz) / dz, "mean_iters": float(cz[2] / cz[3]),
                                "max_iters": int(itz.max()), "ticks": stz["ticks"], "tflops_whole": stz["flops"] / (dz * 1e12),
                                "tflops_solve_kernel": stz["flops"] / (stz["solve_ms"] * 1e9),
                                "fallback_points": int(cz[1]), **sample_check(T, xz, yz, q0, Zz, itz, nsamp=8)}
    print("c5_zoom", json.dumps(out["c5_zoom_recompute"]), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_" + "_".join(which) + ".json"), "w"), indent=1)
