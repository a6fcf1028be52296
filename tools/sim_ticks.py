"""Scheduler model used for the scaling analysis in DESIGN.md: replays the tick loop on the measured per-point
Lanczos iteration counts of the bench workload (profiles/r01_iters_n4000.npz) with the measured launch times of
profiles/r01_panel_sweep.txt, for 1/2/4/8 row-cyclic shards.  CPU only.

    python tools/sim_ticks.py
"""
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
it = np.load(os.path.join(ROOT, "profiles", "r01_iters_n4000.npz"))["iters"].astype(int)
SMS, W = 148, 18944
# ms per launch with k = 1..4 co-resident CTAs per SM, by panel width (n = 4000)
COST = {8: [0, 9.7, 11.98, 16.5, 20.65], 16: [0, 15.3, 21.45, 30.0, 38.3], 32: [0, 25.8, 37.6, 57.0, 73.85]}
FULL_RATE = W / COST[32][4]  # shift-iterations per ms at full occupancy


def tick_ms(n):
    best = 1e9
    for w, c in COST.items():
        per_sm = -(-(-(-n // w)) // SMS)
        best = min(best, (per_sm // 4) * c[4] + c[per_sm % 4])
    return best


def simulate(iters):
    iters = np.sort(iters)[::-1]  # longest-job-first queue (ideal predictor)
    head, rem, t, ticks = 0, np.zeros(0, int), 0.0, 0
    while head < len(iters) or len(rem):
        take = min(W - len(rem), len(iters) - head)
        if take > 0:
            rem = np.concatenate([rem, iters[head:head + take]])
            head += take
        t += tick_ms(len(rem))
        ticks += 1
        rem = rem[rem > 1] - 1
    return t, ticks


if __name__ == "__main__":
    base = None
    for ng in (1, 2, 4, 8):
        t, ticks = simulate(it[0::ng].ravel())
        pts = it.size / (t * 1e-3)
        base = base or pts
        ideal = it[0::ng].sum() / FULL_RATE
        print(f"{ng} GPU: {t / 1e3:6.2f} s per grid, {ticks:3d} ticks, {pts:9.0f} points/s, scaling efficiency {pts / ng / base:.2f}, "
              f"tick-quantisation + tail loss {100 * (1 - ideal / t):.0f} %")
