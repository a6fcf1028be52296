"""Scheduler model used for the scaling analysis in DESIGN.md: replays the tick loop on the measured per-point
Lanczos iteration counts of the bench workload (profiles/r01_iters_n4000.npz) with the measured launch times of the
round-2 solve kernels (profiles/r02_flavor_check_v3.log / _v4.log), for 1/2/4/8 row-cyclic shards.  CPU only.

    python tools/sim_ticks.py [sparse_ms_w8]      # optional: what-if lone-panel latency of the 8-wide sparse launch
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
it = np.load(os.path.join(ROOT, "profiles", "r01_iters_n4000.npz"))["iters"].astype(int)
SMS, W = 148, 9472  # one round of 2 CTAs per SM x 32 shifts (slot_capacity in psa_api.cu)
# ms per launch at n = 4000 by panel width: [sparse: <= 1 panel per SM (8 consumer warps), full round: 2 CTAs per SM]
COST = {8: [6.17, 10.78], 16: [10.44, 19.32], 32: [19.0, 36.06]}
if len(sys.argv) > 1:
    s = float(sys.argv[1]) / COST[8][0]
    for w in COST:
        COST[w][0] *= s
FULL_RATE = 2 * SMS * 32 / COST[32][1]  # shift-iterations per ms at full occupancy


# cluster pipeline over P SMs per panel (profiles/r02_cluster_check.log): time relative to the sparse launch, and how
# many clusters fit on the device at once (cudaOccupancyMaxActiveClusters; clusters must sit inside one GPC)
CLUSTER = {2: (0.55, 66), 4: (0.32, 32)}
USE_CLUSTER = True


def tick_ms(n):
    best = 1e9
    for w, (c1, c2) in COST.items():
        panels = -(-n // w)
        rounds, rest = divmod(panels, 2 * SMS)
        best = min(best, rounds * c2 + (0 if rest == 0 else (c1 if rest <= SMS else c2)))
        if USE_CLUSTER:
            for P, (f, cap) in CLUSTER.items():
                if panels <= cap:
                    best = min(best, c1 * f)
    return best


def simulate(iters, order="ljf"):
    iters = np.sort(iters)[::-1] if order == "ljf" else iters  # longest-job-first queue (ideal predictor)
    cap = W
    head, rem, t, ticks = 0, np.zeros(0, int), 0.0, 0
    while head < len(iters) or len(rem):
        take = min(cap - len(rem), len(iters) - head)
        if take > 0:
            rem = np.concatenate([rem, iters[head:head + take]])
            head += take
        t += tick_ms(len(rem))
        ticks += 1
        rem = rem[rem > 1] - 1
    return t, ticks


if __name__ == "__main__":
    if "--no-cluster" in sys.argv:
        USE_CLUSTER = False
        sys.argv.remove("--no-cluster")
    base = None
    for ng in (1, 2, 4, 8):
        t, ticks = simulate(it[0::ng].ravel())
        pts = it.size / (t * 1e-3)
        base = base or pts
        ideal = it[0::ng].sum() / FULL_RATE
        print(f"{ng} GPU: {t / 1e3:6.2f} s per grid, {ticks:3d} ticks, {pts:9.0f} points/s, scaling efficiency {pts / ng / base:.2f}, "
              f"tick-quantisation + tail loss {100 * (1 - ideal / t):.0f} %")
