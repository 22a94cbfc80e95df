#!/usr/bin/env python
"""bench.py -- (T) FP64 throughput of the B200 engine on BASELINE.json's configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cX]
  torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, NCCL)

Workload by N when --config is absent (BASELINE.json `configs`): N = 8 -> c4 (No=100 Nv=1000, the
configuration the metric is quoted on), N = 2, 4 -> c3 (No=64 Nv=640), N = 1 -> c3 when its 169 GB of
stores fit the GPU, else c2 (No=40 Nv=400).

A "step" is one pass of the hot path (contraction + singles/energy reduction) over
`tuples_per_step` consecutive tuples of this rank's group-and-sort list, on synthetic tensors
(counter-based generator) that live in HBM before the timed region.  Every step walks
new tuples, so each step reads GBs of slices that were not touched by the previous one (inputs
larger than L2).

  value       whole-job FP64 TFLOP/s = 12 No^3 (No+Nv) x tuples of all ranks / device time,
              device time = CUDA events on the engine's stream, max over ranks
  e2e         the same metric through the C-ABI with HOST buffers, per step: ingest the step's inputs
              from pinned host memory + run the step's tuples + read the energy back.  N = 1 with host
              tensors that fit the box: the full CTF-layout tensors through atrip_b200_load_* (what
              Atrip::run does with one rank); otherwise every rank uploads, through
              atrip_b200_upload_slices, the slices it owns that the step's tuples read, from pinned
              host buffers in the reference's slice layout (what Atrip::run does with several ranks:
              SliceUnion sources, SliceUnion.cxx:305-332)
  parity      before timing: the golden whole-run case (No=10 Nv=40, reference energy from
              tests/golden) on the same N GPUs with the same store sharding and transport, and three
              tuples of the bench configuration against the reference's own functions on the host
  roofline    contraction kernel: algorithmic FLOP / launch duration (events around every launch)
              against the FP64 tensor (DMMA) ceiling measured live on this GPU
  cpu_baseline / --impl reference
              the reference's own doubles/singles/energy functions (oracle/_ref, compiled from
              the reference sources) on all host cores, on a bounded sample of the same tuples
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {  # BASELINE.json configs; scale keeps |E| = O(1e-2..1) for the parity checks
    "c1": dict(No=10, Nv=40, scale=0.01, tuples_per_step=11440, desc="No=10 Nv=40 (CPU-runnable case)"),
    "c2": dict(No=40, Nv=400, scale=0.001, tuples_per_step=196608, desc="No=40 Nv=400 FP64 random tensors"),
    "c5s": dict(No=32, Nv=480, scale=0.001, tuples_per_step=98304, desc="No=32 high Nv/No (c5 scaled to 1 GPU)"),
    "c3": dict(No=64, Nv=640, scale=0.0005, tuples_per_step=16900,
               desc="No=64 Nv=640 FP64 random tensors (170 GB of stores)"),
    "c4": dict(No=100, Nv=1000, scale=0.0002, tuples_per_step=3900, min_gpus=8,
               desc="No=100 Nv=1000 FP64 random tensors (1.0 TB of stores over 8 GPUs)"),
    "c5": dict(No=32, Nv=1200, scale=0.0005, tuples_per_step=59200, min_gpus=4,
               desc="No=32 Nv=1200 FP64 random tensors, high Nv/No (0.5 TB of stores)"),
}
SEED = 12345
GOLDEN_RUN = dict(No=10, Nv=40, seed=12345, scale=0.01)  # tests/golden/reference_vectors.json "runs"
VENDOR_FP64_TENSOR_TFLOPS = 40.0  # NVIDIA B200 datasheet, FP64 tensor core, dense (HGX B200: 37 per GPU)
E_ABS, E_REL = 1e-10, 1e-12       # north_star tolerances


def store_bytes(cfg, field="real"):
    """bytes of one full copy of the engine's stores (DESIGN.md "Data layout")"""
    No, Nv = cfg["No"], cfg["Nv"]
    z = 2 if field == "complex" else 1
    Kp = (z * (No + Nv) + 15) // 16 * 16
    return 8 * (z * Nv * No * No * Kp + (Nv * Nv + Nv) * No * Kp + z * (Nv * (Nv + 1) // 2) * No * No)


def pick_config(n_gpus):
    """the workload BASELINE.json quotes for N GPUs"""
    if n_gpus >= 8:
        return "c4"
    if n_gpus >= 2:
        return "c3"
    try:
        import torch
        free, _total = torch.cuda.mem_get_info(int(os.environ.get("LOCAL_RANK", "0")))
    except Exception:
        return "c2"
    # stores + 2 x 1 GiB of class cubes + ingest staging + CUDA context, against the memory that is free NOW
    return "c3" if free >= store_bytes(CONFIGS["c3"]) + 5 * 2 ** 30 else "c2"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        # median over the samples taken under load (upper half: idle samples between steps drop out)
        load = sm[len(sm) // 2:] if sm else []
        power = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": load[len(load) // 2] if load else None,
                "sm_max_mhz": int(float(self.rows[0][1])) if self.rows else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": reasons}


# ----------------------------------------------------------------------------- CPU reference (checker / baseline)
def _ref_tuple_energy(o, r, No, Nv, scale, abc, epsi, epsa, tai, scratch):
    """energy of one tuple through the reference's own L1 functions (oracle/_ref) or the C port"""
    S = o.synth_tuple_slices(No, Nv, abc, seed=SEED, scale=scale)  # input generation: not timed
    t0 = time.perf_counter()
    f = r if r is not None else o
    T = f.doubles(No, Nv, S, scratch) if r is not None else f.doubles(No, Nv, S)
    Z = f.singles(No, Nv, abc, tai, S, T)
    eps = float(epsa[abc[0]] + epsa[abc[1]] + epsa[abc[2]])
    same = (abc[0] == abc[1]) != (abc[1] == abc[2])
    e = (f.energy_same if same else f.energy_distinct)(eps, No, epsi, T, Z)
    return e, time.perf_counter() - t0


def _cpu_worker(args):
    """time the reference's L1 functions on `n` tuples in this process (1 BLAS thread)"""
    No, Nv, scale, tuples, use_ref = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import numpy as np
    from oracle.oracle import EPS_A, EPS_I, TAI, Oracle, Reference
    o = Oracle()
    r = Reference() if use_ref else None
    epsi, epsa = o.fill(SEED, EPS_I, scale, No), o.fill(SEED, EPS_A, scale, Nv)
    tai = o.fill(SEED, TAI, scale, No * Nv)
    scratch = (np.empty(No ** 3), np.empty(No ** 3))
    busy, esum, each = 0.0, 0.0, []
    for abc in tuples:
        e, dt = _ref_tuple_energy(o, r, No, Nv, scale, abc, epsi, epsa, tai, scratch)
        busy += dt
        esum += e
        each.append(e)
    return busy, esum, each


def cpu_reference_step(cfg, tuples, cores, pool):
    """one bounded step of the reference CPU path on `cores` worker processes; returns
    (seconds = slowest worker's compute time, tuples done, energy sum)"""
    from oracle.oracle import Reference
    use_ref = Reference.available()
    chunks = [tuples[i::cores] for i in range(cores)]
    res = pool.map(_cpu_worker, [(cfg["No"], cfg["Nv"], cfg["scale"], [tuple(int(x) for x in t) for t in ch], use_ref)
                                 for ch in chunks])
    return max(r[0] for r in res), len(tuples), sum(r[1] for r in res), ("reference" if use_ref else "port")


def cpu_sample_tuples(cfg, n):
    """n tuples drawn at random (fixed seed) from the a <= b <= c list, without the product's library"""
    import numpy as np
    Nv = cfg["Nv"]
    out = []
    rng = np.random.RandomState(SEED)
    while len(out) < n:  # uniform over a <= b <= c, not all equal
        abc = np.sort(rng.randint(0, Nv, size=(n, 3)), axis=1)
        abc = abc[~((abc[:, 0] == abc[:, 1]) & (abc[:, 1] == abc[:, 2]))]
        out.extend(abc.tolist())
    return np.array(out[:n], dtype=np.int64)


def host_cores():
    """host threads this process may use (what `nproc` prints: the affinity mask, not the machine's total)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def est_cpu_tuples(cfg, cores, seconds):
    flops = 12.0 * cfg["No"] ** 3 * (cfg["No"] + cfg["Nv"])
    per_tuple = flops / 6.0e9  # ~6 GF/s/core measured for the reference dgemm path
    return max(cores, int(seconds / per_tuple) * cores)


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    flops_per_tuple = 12.0 * cfg["No"] ** 3 * (cfg["No"] + cfg["Nv"])
    per_step = est_cpu_tuples(cfg, cores, 6.0)
    tuples = cpu_sample_tuples(cfg, per_step * (args.steps + args.warmup))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        times, kind = [], "port"
        for s in range(args.steps + args.warmup):
            sec, n, _, kind = cpu_reference_step(cfg, tuples[s * per_step:(s + 1) * per_step], cores, pool)
            if s >= args.warmup:
                times.append(sec)
    total_s = sum(times)
    value = flops_per_tuple * per_step * args.steps / total_s / 1e12
    line = {"impl": "reference", "metric": "(T) FP64 TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"] + ": " + cfg["desc"], "No": cfg["No"], "Nv": cfg["Nv"],
                       "tuples_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": cores, "kind": kind,
                             "sample": f"{per_step} tuples/step drawn uniformly from the a<=b<=c list, reference "
                                       "doubles_contribution+singles_contribution+get_energy_* (dgemm path, wheel "
                                       "OpenBLAS, 1 thread per process), one process per core"},
            "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extrapolated_full_wall_s": flops_per_tuple * (cfg["Nv"] * (cfg["Nv"] + 1) * (cfg["Nv"] + 2) // 6 - cfg["Nv"])
            / (value * 1e12)}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- parity (checker: oracle/, tests/golden)
def parity_tuples(cfg):
    """three tuples of the bench configuration: far-apart distinct, a == b ("same" kernel), adjacent"""
    Nv = cfg["Nv"]
    return [(3, Nv // 2, Nv - 3), (17 % Nv, 17 % Nv, (16 * Nv) // 25), (Nv // 4, Nv // 4 + 1, Nv // 4 + 2)]


def _parity_reference_worker(No, Nv, scale, tuples, q):
    """host side of the parity check, in its own process while the GPUs fill their stores"""
    try:
        from oracle.oracle import Reference
        _, _, each = _cpu_worker((No, Nv, scale, tuples, Reference.available()))
        q.put(("reference" if Reference.available() else "port", each))
    except Exception as e:  # the checker is missing: parity is reported as not run, the bench goes on
        q.put(("error: " + repr(e), None))


def golden_energy():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        g = json.load(f)
    for r in g["runs"]:
        if all(r[k] == GOLDEN_RUN[k] for k in GOLDEN_RUN) and not r["with_J"]:
            return float.fromhex(r["energy"])
    raise RuntimeError("golden run missing from tests/golden/reference_vectors.json")


def golden_run_check(atrip_b200, capi, dist, rank, world, local, sharded, transport):
    """whole (T) run of the golden case on the same GPUs, sharding and transport as the bench"""
    g = GOLDEN_RUN
    eng = atrip_b200.Engine(g["No"], g["Nv"], device=local, rank=rank, nranks=world, resident=not sharded,
                            transport=transport, batch_tuples=97)
    if world > 1:
        box = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0])
    eng.fill_synthetic(g["seed"], g["scale"])
    n = eng.build_tuples(capi.GROUP_AND_SORT)
    e = 0.0
    for lo in range(0, n, 331):  # several calls: exercises the start-up and the prefetch of each call
        e += eng.run(lo, min(331, n - lo))[0]
    tot = -float(eng.allreduce([e])[0])
    eng.close()
    want = golden_energy()
    return {"No": g["No"], "Nv": g["Nv"], "energy": tot, "reference": want, "abs": abs(tot - want),
            "rel": abs(tot - want) / abs(want)}


# ----------------------------------------------------------------------------- e2e helpers
def host_tensors(cfg, device):
    """pinned host tensors in CTF layout holding the synthetic inputs (N = 1 e2e leg)"""
    import torch
    from atrip_b200 import capi
    No, Nv = cfg["No"], cfg["Nv"]
    sizes = {0: No, 1: Nv, 2: Nv * No, 3: Nv * Nv * No * No, 4: Nv * Nv * No * No, 5: No ** 3 * Nv, 6: Nv ** 3 * No}
    out = {}
    for tid, n in sizes.items():
        t = torch.empty(n, dtype=torch.float64, pin_memory=True)
        capi.synth_to_host(device, SEED, tid, cfg["scale"], 0, n, t.data_ptr())
        out[tid] = t
    return out


def step_input_slices(capi, tuples, Nv, rank, world):
    """(kind -> [n,2] int64 (x,y)) the slices THIS rank holds that the tuples read: what SliceUnion::init
    would have sliced for them (ownership: schedule.hpp / RankMap.cxx:35-85)"""
    import numpy as np
    t = np.asarray(tuples, dtype=np.int64)
    t = t[t.any(axis=1)]
    a, b, c = t[:, 0], t[:, 1], t[:, 2]
    mine = lambda x: (x % world) == rank
    xs = np.unique(np.concatenate([a, b, c]))
    xs = xs[mine(xs)]
    single = np.stack([xs, np.zeros_like(xs)], axis=1)
    # ordered pairs (first index owns): (b,c) (a,c) (c,b) (a,b) (c,a) (b,a)
    py = np.concatenate([b, a, c, a, c, b])
    pz = np.concatenate([c, c, b, b, a, a])
    keep = mine(py)
    pairs = np.unique(py[keep] * Nv + pz[keep])
    ordered = np.stack([pairs // Nv, pairs % Nv], axis=1)
    # their hole parts come from TABIJ(min, max)
    lo, hi = np.minimum(ordered[:, 0], ordered[:, 1]), np.maximum(ordered[:, 0], ordered[:, 1])
    tp = np.unique(lo * Nv + hi)
    tab = np.stack([tp // Nv, tp % Nv], axis=1)
    # VABIJ (b,c) (a,c) (a,b): held by the owner of either index
    vy, vz = np.concatenate([b, a, a]), np.concatenate([c, c, b])
    keep = mine(vy) | mine(vz)
    vp = np.unique(vy[keep] * Nv + vz[keep])
    vab = np.stack([vp // Nv, vp % Nv], axis=1)
    return {capi.TA: single, capi.VIJKA: single, capi.VABCI: ordered, capi.TABIJ: tab, capi.VABIJ: vab}


# ----------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS),
                    help="default: by --gpus (8: c4, 2 and 4: c3, 1: c3 if it fits the GPU, else c2)")
    ap.add_argument("--tuples-per-step", type=int, default=0)
    ap.add_argument("--replicate", action="store_true",
                    help="N>1: every GPU holds a full replica of the stores (no slice exchange)")
    ap.add_argument("--transport", type=int, default=0, choices=[0, 1, 2],
                    help="N>1 slice exchange: 1 NCCL send/recv, 2 P2P copy-engine pulls, 0 engine default")
    ap.add_argument("--field", default="real", choices=["real", "complex"],
                    help="complex: Atrip::run<Complex> instantiation (4x the FLOPs per tuple, Atrip.cxx:578-580); "
                         "device-resident synthetic stores only (no e2e / CPU legs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-mode", default="auto", choices=["auto", "tensors", "slices"],
                    help="e2e ingest: the full CTF-layout host tensors (one rank, tensors that fit the box) or the "
                         "per-owner slices of the step's inputs; auto picks tensors when it can")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference" or os.environ.get("ATRIP_BENCH_ALLOW_SHORT"), "warmup >= 3"
    name = args.config or pick_config(args.gpus)
    cfg = dict(CONFIGS[name], name=name)
    if args.tuples_per_step:
        cfg["tuples_per_step"] = args.tuples_per_step
    if args.field == "complex":
        args.no_e2e = args.no_cpu = args.no_parity = True
        assert args.impl == "ours", "the reference arm times the real (double) instantiation"
    assert args.impl == "reference" or args.gpus >= cfg.get("min_gpus", 1), \
        f"{name} needs at least {cfg.get('min_gpus')} GPUs (stores are sharded over the ranks)"
    if args.impl == "reference":
        return run_reference_arm(args, cfg)

    import numpy as np
    import torch
    import torch.distributed as dist
    import atrip_b200
    from atrip_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the engine has no CPU fallback"
    assert world == args.gpus or world == 1, f"WORLD_SIZE {world} != --gpus {args.gpus}"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    No, Nv, tps = cfg["No"], cfg["Nv"], cfg["tuples_per_step"]
    K, W = args.steps, args.warmup
    sharded = world > 1 and not args.replicate
    t_setup = time.perf_counter()

    # N > 1: every GPU stores the slices it owns (RankMap round robin) and fetches the rest of each
    # batch from its peers on side streams, one batch ahead of the compute
    def make_engine(c):
        return atrip_b200.Engine(c["No"], c["Nv"], device=local, rank=rank, nranks=world, resident=not sharded,
                                 transport=args.transport,
                                 field=capi.FIELD_COMPLEX if args.field == "complex" else capi.FIELD_REAL)
    try:
        eng = make_engine(cfg)
    except capi.EngineError as e:
        # c3 was picked from the GPU's total memory; if somebody else holds part of it, measure c2 instead
        if args.config is not None or name != "c3" or world != 1:
            raise
        print(f"bench.py: c3 does not fit this GPU right now ({e}); falling back to c2", file=sys.stderr)
        name = "c2"
        cfg = dict(CONFIGS[name], name=name)
        No, Nv, tps = cfg["No"], cfg["Nv"], cfg["tuples_per_step"]
        eng = make_engine(cfg)

    # ------------------------------------------------ parity, part 1: host side starts now (rank 0)
    ptuples = parity_tuples(cfg)
    pq = pproc = None
    if rank == 0 and not args.no_parity:
        ctx = mp.get_context("spawn")
        pq = ctx.Queue()
        pproc = ctx.Process(target=_parity_reference_worker, args=(No, Nv, cfg["scale"], ptuples, pq), daemon=True)
        pproc.start()
    parity = None
    if not args.no_parity:
        parity = {"n_gpus": world, "golden_run": golden_run_check(atrip_b200, capi, dist, rank, world, local, sharded,
                                                                  args.transport)}

    if world > 1:  # the engine's own NCCL communicator; its 128-byte id travels over torch.distributed
        box = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0])

    def sum_over_ranks(vals):
        # final energy reduction (replaces MPI_Reduce(SUM), Atrip.cxx:1094-1107): ncclAllReduce
        return [float(x) for x in eng.allreduce(vals)]

    eng.fill_synthetic(SEED, cfg["scale"])
    n_list = eng.build_tuples(capi.GROUP_AND_SORT)
    tps = min(tps, n_list // (K + W))
    flops_per_tuple = eng.flops_per_tuple
    peak = capi.measure_dmma_peak(local) if rank == 0 else 0.0

    # ------------------------------------------------ parity, part 2: tuples of the bench configuration
    if parity is not None:
        got = [eng.tuple_debug(*abc, cubes=False)[0] for abc in ptuples]  # collective: remote slices are fetched
        if rank == 0:
            kind, want = pq.get(timeout=1200)
            pproc.join(timeout=30)
            gr = parity["golden_run"]
            if want is None:  # the host-side checker could not run: only the golden run was compared
                parity.update({"tuples": None, "checker": kind, "max_rel": gr["rel"],
                               "ok": bool(gr["abs"] <= E_ABS and gr["rel"] <= E_REL)})
            else:
                rels = [abs(g - w) / abs(w) for g, w in zip(got, want)]
                parity.update({"tuples": [{"abc": list(t), "energy": g, "reference": w, "rel": r}
                                          for t, g, w, r in zip(ptuples, got, want, rels)],
                               "checker": kind, "max_rel": max(rels + [parity["golden_run"]["rel"]]),
                               "tolerance": {"energy_abs": E_ABS, "energy_rel": E_REL}})
                parity["ok"] = bool(max(rels) <= E_REL and gr["abs"] <= E_ABS and gr["rel"] <= E_REL)
    setup_s = time.perf_counter() - t_setup

    # ------------------------------------------------ value: inputs resident in HBM
    def step(i):
        e, ct = eng.run(i * tps, tps)
        tm = eng.last_timing()
        tot = sum_over_ranks([e, ct, float(tm["tuples"])])
        return tm, tot

    for i in range(W):
        step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, tuples_done, ck_ms, rk_ms, nbatches, energy = 0.0, 0, 0.0, 0.0, 0.0, 0, 0.0
    xbytes = xmsgs = gap_ms = plan_ms = start_ms = hits = fetched = 0.0
    step_energy = {}
    for i in range(W, W + K):
        tm, tot = step(i)
        ex, ph = eng.last_exchange(), eng.last_phases()
        xbytes += ex["bytes"]
        xmsgs += ex["messages"]
        gap_ms += ph["gap_ms"]
        start_ms += ph["startup_gap_ms"]
        plan_ms += ph["plan_ms"]
        hits += ph["cache_hits"]
        fetched += ph["fetched"]
        dev_ms += tm["total_ms"]
        launches += tm["contract_launches"] + tm["reduce_launches"]
        tuples_done += tot[2]
        ck_ms += tm["contract_ms"] * ph["batches"]
        rk_ms += tm["reduce_ms"] * ph["batches"]
        nbatches += ph["batches"]
        energy += tot[0]
        step_energy[i] = tot[0]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    gap_ms_max = max_over_ranks(gap_ms + start_ms)
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    value = flops_per_tuple * tuples_done / (dev_ms * 1e-3) / 1e12

    # contraction kernel roofline: mean launch duration over EVERY launch of the timed steps (events
    # around each launch on the engine's stream)
    batch_tuples = min(tps, eng.batch_tuples)
    contract_ms = ck_ms / max(nbatches, 1)
    # tuples per launch averaged over the step (its last batch may be shorter)
    mean_batch = tps / max(1, -(-tps // eng.batch_tuples))
    achieved = flops_per_tuple * mean_batch / (contract_ms * 1e-3) / 1e12 if contract_ms > 0 else None

    # ------------------------------------------------ e2e: host buffers through the C-ABI
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, cfg, eng, capi, torch, np, rank, world, local, tps, W, K, flops_per_tuple, barrier,
                      max_over_ranks, sum_over_ranks, step_energy)

    # ------------------------------------------------ CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        n = est_cpu_tuples(cfg, cores, args.cpu_seconds)
        tl = cpu_sample_tuples(cfg, n)
        with mp.get_context("spawn").Pool(cores) as pool:
            sec, n, _, kind = cpu_reference_step(cfg, tl, cores, pool)
        cpu = {"value": flops_per_tuple * n / sec / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": kind,
               "sample": f"{n} tuples drawn uniformly from the a<=b<=c list, reference doubles+singles+energy "
                         f"functions, {cores} single-threaded processes, {sec:.1f} s"}

    if rank == 0:
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(cfg["name"], {}).get("contract_dram_bytes_per_launch")
            except Exception:
                traffic = None
        total_tuples = Nv * (Nv + 1) * (Nv + 2) // 6 - Nv
        stores = ("sharded: owned slices + fetch cache (persistent over batches and calls) filled one batch ahead "
                  "on side streams, " + {0: "engine default transport (P2P copy-engine pulls over NVLink)",
                                         1: "ncclSend/ncclRecv", 2: "P2P copy-engine pulls over NVLink"}[args.transport]
                  if sharded else "replica on every GPU" if world > 1 else "single GPU")
        line = {
            "metric": "(T) FP64 TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128" if args.field == "complex" else "f64", "data": "synthetic",
            "config": {"workload": cfg["name"] + ": " + cfg["desc"] + (" [complex field]" if args.field == "complex" else ""),
                       "No": No, "Nv": Nv, "tuples_per_step": tps,
                       "tuples_per_step_all_ranks": tps * world, "distribution": "group_and_sort (GPU == node)",
                       "stores": stores,
                       "l2_policy": "inputs larger than L2: every step walks new tuples (GBs of new slices)",
                       "seed": SEED, "scale": cfg["scale"],
                       "exchange": {"rank0_recv_bytes_per_step": xbytes / K, "rank0_messages_per_step": xmsgs / K,
                                    "rank0_recv_GBps": xbytes / (dev_ms * 1e-3) / 1e9 if dev_ms else None,
                                    "rank0_cache_hit_slices_per_step": hits / K,
                                    "rank0_fetched_slices_per_step": fetched / K,
                                    "rank0_host_plan_ms_per_step": plan_ms / K,
                                    "rank0_idle_before_contraction_ms_per_step": gap_ms / K,  # fetch not landed OR cube buffer still being reduced
                                    "rank0_startup_gap_ms_per_step": start_ms / K,
                                    "max_rank_gap_fraction_of_step": gap_ms_max / dev_ms if dev_ms else None}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * world, "parity": parity,
            "roofline": {"bound": "tensor", "kernel": "contract_kernel (FP64 DMMA)", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if achieved and peak else None,
                         "traffic": traffic, "launch_ms": contract_ms, "tuples_per_launch": mean_batch,
                         "launches_timed": nbatches,
                         "peak_source": "FP64 DMMA ceiling measured live (atrip_b200_measure_dmma_peak); "
                                        "MEASURED_PEAKS.json has no FP64 figure",
                         "vendor_peak": VENDOR_FP64_TENSOR_TFLOPS,
                         "frac_of_vendor_peak": achieved / VENDOR_FP64_TENSOR_TFLOPS if achieved else None,
                         "step_frac": value / (peak * world) if peak else None,
                         "step_frac_of_vendor_peak": value / (VENDOR_FP64_TENSOR_TFLOPS * world),
                         "reduce_ms_per_launch": rk_ms / max(nbatches, 1)},
            "cpu_baseline": cpu,
            "wall_ms_per_step": wall_ms / K, "setup_s": setup_s,
            "frac_of_fp64_tensor_peak": value / (peak * world) if peak else None,
            "extrapolated_full_wall_s": flops_per_tuple * total_tuples / (value * 1e12),
            "energy_partial": -energy,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity.get("ok", False):
        sys.exit("parity check failed: " + json.dumps(parity))


def run_e2e(args, cfg, eng, capi, torch, np, rank, world, local, tps, W, K, flops_per_tuple, barrier,
            max_over_ranks, sum_over_ranks, step_energy):
    """the metric through the C-ABI with host buffers (see the module docstring)"""
    No, Nv = cfg["No"], cfg["Nv"]
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except ImportError:
        avail = None
    full_bytes = 8 * (2 * Nv * Nv * No * No + No ** 3 * Nv + Nv ** 3 * No)
    use_tensors = world == 1 and full_bytes <= 64e9 and (avail is None or avail >= 1.5 * full_bytes)
    if args.e2e_mode != "auto":
        assert args.e2e_mode == "slices" or world == 1, "the CTF-layout tensor ingest is the one-rank path"
        use_tensors = args.e2e_mode == "tensors"
    if use_tensors:
        # ---- one rank: the full CTF-layout tensors, as Atrip::run ingests them
        host = host_tensors(cfg, local)
        h2d = sum(t.numel() * 8 for t in host.values())

        def e2e_step(i):
            eng.load_all(*[host[k].data_ptr() for k in range(7)])   # H2D + device re-tiling
            e, ct = eng.run(i * tps, tps)                           # compute
            return sum_over_ranks([e, ct, float(eng.last_timing()["tuples"])])  # D2H of the result

        for i in range(min(W, 2)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        done, ok = 0.0, True
        for i in range(W, W + K):
            tot = e2e_step(i)
            done += tot[2]
            ok = ok and tot[0] == step_energy.get(i, tot[0])
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        return {"value": flops_per_tuple * done / e2e_s / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 16, "ms_per_step": e2e_s / K * 1e3, "energy_matches_resident_run": ok,
                "what": "per step: ingest all pinned host tensors (CTF layout, atrip_b200_load_*) + run the "
                        "step's tuples + read the energy"}
    # ---- several ranks (or host tensors larger than the box's memory): per-owner slices
    first = W * tps  # every e2e step re-runs the first timed step: one step's slices are kept on the host
    tl = eng.get_tuples()[first:first + tps]
    need = step_input_slices(capi, tl, Nv, rank, world)
    nbytes = sum(len(xy) * eng.slice_elems(k) * 8 for k, xy in need.items())
    worst = max_over_ranks(float(nbytes))
    if avail is not None and worst * world * 1.3 > avail:
        return {"skipped": f"host memory: {worst * world / 1e9:.0f} GB of pinned slices needed for {world} ranks"}
    pool = {}
    for k, xy in need.items():  # host shards in the reference's slice layout, read back from the filled stores
        buf = torch.empty(max(1, len(xy) * eng.slice_elems(k)), dtype=torch.float64, pin_memory=True)
        eng.read_slices(k, xy, out=buf.data_ptr())
        pool[k] = buf

    def e2e_step():
        for k, xy in need.items():                                  # H2D + device re-tiling, owned slices only
            eng.upload_slices(k, xy, pool[k].data_ptr())
        e, ct = eng.run(first, tps)                                 # compute (remote slices from the peers)
        return sum_over_ranks([e, ct, float(eng.last_timing()["tuples"])])  # D2H of the result

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    done, ok = 0.0, True
    for _ in range(K):
        tot = e2e_step()
        done += tot[2]
        ok = ok and tot[0] == step_energy.get(W, tot[0])
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    total_h2d = sum_over_ranks([float(nbytes)])[0]
    return {"value": flops_per_tuple * done / e2e_s / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": int(total_h2d),
            "d2h_bytes_per_step": 16 * world, "ms_per_step": e2e_s / K * 1e3, "energy_matches_resident_run": ok,
            "host_pinned_bytes_per_rank_max": int(worst),
            "what": "per step, every rank: upload the slices it owns that the step's tuples read (pinned host "
                    "buffers in the reference's slice layout, atrip_b200_upload_slices; all of its TA/VIJKA "
                    "slices and the VABCI/TABIJ/VABIJ pairs of the step) + run the step's tuples, remote slices "
                    "from the peers + read the energy; every e2e step repeats the first timed step"}


if __name__ == "__main__":
    main()
