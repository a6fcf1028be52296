"""The library's OWN multi-GPU path (the one a Julia `ccall` host takes): ONE process drives N devices through
psa_gpu_init(N) + psa_gpu_set_matrix + psa_gpu_grid on the bench workload (n = 4000 Schur factor, 400x400 grid).
Reports points/s per N (wall clock around the blocking call, which includes H2D of x/y/q0, the NCCL gather of the
sigma_min tiles on device 0 and the D2H of Z) and checks bitwise equality with the 1-device result.

    python tools/inproc_scale.py [--n 4000] [--npts 400] [--gpus 1,2,4,8] [--steps 2]
Writes gpurun_out/inproc_scale.json."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import psa_b200_loader

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4000)
ap.add_argument("--npts", type=int, default=400)
ap.add_argument("--gpus", default="1,2")
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
P = psa_b200_loader.load_package()
import torch
avail = torch.cuda.device_count()
T, schur_s, how = bench.build_workload(a.n, a.npts)
ax, x, y = bench.grid_for(T, a.npts, P)
q0 = bench.start_vector(a.n)
out = {"n": a.n, "grid": [a.npts, a.npts], "schur": how, "runs": []}
Zref = itref = None
for ng in [int(v) for v in a.gpus.split(",")]:
    if ng > avail:
        continue
    with P.PsaGpu(ng) as g:
        t0 = time.time(); g.set_matrix(T); t_set = time.time() - t0
        g.grid(x, y, q0)  # warm-up (allocations, NCCL communicator set-up)
        ts = []
        for _ in range(a.steps):
            t0 = time.time(); Z, it, counts, algo = g.grid(x, y, q0); ts.append(time.time() - t0)
        st = g.last_stats()
    if Zref is None:
        Zref, itref = Z, it
    r = {"ngpus": ng, "points_per_s": len(x) * len(y) / min(ts), "seconds": min(ts), "set_matrix_s": t_set,
         "bitwise_equal_to_first": bool(np.array_equal(Z, Zref) and np.array_equal(it, itref)),
         "ticks_dev0": st["ticks"], "grid_ms_dev0": st["grid_ms"], "max_device_ms": st["max_device_ms"],
         "mean_iters": float(counts[2] / counts[3])}
    out["runs"].append(r)
    print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "inproc_scale.json"), "w"), indent=1)
