"""Host-only statistics of the tuple distribution and of the slice fetch schedule (no GPU needed):
what the reference's bench/tuples-distribution.cxx reports for its slice database, here for the
engine's group-and-sort lists (GPU == node), ownership map and per-batch fetch plan.

  python tools/distribution_stats.py [c2 c3 c4 c5]      -> table on stdout (kept in profiles/)

Per config and rank count: list length per rank, fake tuples (imbalance), and -- over batches
sampled at the start, middle and end of rank 0's and the last rank's list -- remote slices fetched
per tuple by store, messages (coalesced slot ranges) per batch and bytes per tuple, converted to
GB/s at the measured per-GPU tuple rate of round 1 (~33 TFLOP/s per GPU)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atrip_b200 import capi

CONFIGS = {"c2": (40, 400), "c3": (64, 640), "c4": (100, 1000), "c5": (32, 1200)}
NSM = 148


def default_batch(No):
    """engine.cu: create_impl batch sizing (real field)"""
    p = capi.host_plan(No)
    per_tuple = 3 * p["row_tiles"] * p["col_tiles"]
    nb = (No + 7) // 8
    cap = max(16, (1 << 30) // (3 * nb ** 3 * 512 * 8))
    b = max(NSM, (NSM * 24 + per_tuple - 1) // per_tuple * 4)
    b = min(b, cap)
    b = (b + NSM - 1) // NSM * NSM
    return min(b, cap)


def main(names):
    print("config ranks rank   tuples/rank  fakes   batch  A/tuple  B/tuple  V/tuple  msgs/batch  MB/tuple  GB/s@33TF")
    for name in names:
        No, Nv = CONFIGS[name]
        Kp = (No + Nv + 15) // 16 * 16
        sz = [No * No * Kp * 8, No * Kp * 8, No * No * 8]  # bytes per A, B, V slice
        flops = 12.0 * No ** 3 * (No + Nv)
        t_tuple = flops / 33e12
        batch = default_batch(No)
        for n in (2, 4, 8):
            for rank in sorted({0, n - 1}):
                t0 = time.time()
                tl = capi.host_tuples(capi.GROUP_AND_SORT, Nv, rank, n, pad=True)
                fakes = int((tl.sum(axis=1) == 0).sum())
                owned = capi.shard_sizes(Nv, rank, n)
                nb = len(tl) // batch
                picks = sorted({0, 1, nb // 4, nb // 2, (3 * nb) // 4, max(nb - 2, 0)})
                cnt = np.zeros(3)
                msgs, tuples = 0, 0
                for k in picks:
                    abc = tl[k * batch:(k + 1) * batch]
                    recs, ranges = capi.plan_batch(Nv, rank, n, abc, owned)
                    for kind in range(3):
                        cnt[kind] += ranges[ranges[:, 1] == kind][:, 3].sum()
                    msgs += len(ranges)
                    tuples += len(abc)
                per = cnt / tuples
                mb = float((per * sz).sum()) / 1e6
                print(f"{name:5s} {n:5d} {rank:4d} {len(tl):13d} {fakes:6d} {batch:7d} {per[0]:8.4f} {per[1]:8.3f} "
                      f"{per[2]:8.3f} {msgs / len(picks):11.1f} {mb:9.3f} {mb / 1e3 / t_tuple:10.1f}"
                      f"   ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["c2", "c3", "c4", "c5"])
