"""Host-only dry run of the tuple distribution and the slice-fetch schedule: what the reference's
bench/tuples-distribution.cxx does with a fake ClusterInfo (bench/tuples-distribution.cxx:62-76) and 16-byte
dummy slices (ATRIP_DRY) -- run the real slice protocol for `ranks` ranks without a cluster and count, per rank,
the slices it has to fetch (:560-569).  Here: the engine's group-and-sort lists (GPU == node), ownership map and
PERSISTENT fetch cache, replayed batch by batch exactly as atrip_b200_run walks them
(atrip_b200_host_check_schedule: every record checked against an independent model of the cache).

  python tools/tuples_distribution.py --no 32 --nv 1200 --ranks 8 [--batch B] [--calls C]

Columns: tuples in the rank's list (fakes = FAKE_TUPLE padding), slices fetched per store over the WHOLE list
(A = TAPHH+HHHA, B = ABPH+TABHH ordered pairs, V = ABHH), remote slices found in the cache instead (the
reference's Recycled / exact-match cases), copy ranges (messages), GB fetched, and the average fetch bandwidth the
rank needs at 33 TFLOP/s per GPU."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atrip_b200 import capi
from tools.distribution_stats import default_batch


def engine_caps(No, Nv, need):
    """engine.cu: ensure_caches -- 3 windows of one batch, more while the caches stay under 4 GiB"""
    Kp = (No + Nv + 15) // 16 * 16
    sz = [No * No * Kp * 8, No * Kp * 8, No * No * 8]
    mult = 3
    while mult < 8 and sum((mult + 1) * need[k] * sz[k] for k in range(3)) <= 4 * 2 ** 30:
        mult += 1
    return [max(mult * need[0], 3), max(mult * need[1], 6), max(mult * need[2], 3)], mult


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no", type=int, required=True)
    ap.add_argument("--nv", type=int, required=True)
    ap.add_argument("--ranks", type=int, required=True)
    ap.add_argument("--batch", type=int, default=0, help="tuples per device batch (default: the engine's choice)")
    ap.add_argument("--calls", type=int, default=10, help="run calls the list is cut into")
    a = ap.parse_args()
    No, Nv, n = a.no, a.nv, a.ranks
    batch = a.batch or default_batch(No)
    Kp = (No + Nv + 15) // 16 * 16
    sz = [No * No * Kp * 8, No * Kp * 8, No * No * 8]
    t_tuple = 12.0 * No ** 3 * (No + Nv) / 33e12
    print(f"# No {No} Nv {Nv} ranks {n} batch {batch} calls {a.calls}: slice bytes A {sz[0]} B {sz[1]} V {sz[2]}")
    print("rank     tuples   fakes  cache slots (A,B,V)        fetched A        B        V    cache hits A        B        V"
          "   ranges  fetched GB  GB/s@33TF  slices/tuple")
    for r in range(n):
        t0 = time.time()
        tl = capi.host_tuples(capi.GROUP_AND_SORT, Nv, r, n, pad=True)
        fakes = int((tl.sum(axis=1) == 0).sum())
        need = capi.cache_need(Nv, r, n, tl, batch)
        caps, mult = engine_caps(No, Nv, need)
        st = capi.check_schedule(Nv, r, n, tl, batch, caps, calls=a.calls)
        f, h = st["fetched"], st["hits"]
        gb = sum(f[k] * sz[k] for k in range(3)) / 1e9
        real = len(tl) - fakes
        print(f"{r:4d} {len(tl):10d} {fakes:7d}  {str(tuple(caps)):24s} {f[0]:10d} {f[1]:10d} {f[2]:8d}   {h[0]:10d} {h[1]:10d} {h[2]:8d}"
              f" {st['ranges']:8d} {gb:11.1f} {gb / (real * t_tuple):10.1f} {sum(f) / real:13.3f}   ({time.time() - t0:.0f} s)",
              flush=True)


if __name__ == "__main__":
    main()
