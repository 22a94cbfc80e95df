#!/usr/bin/env python
"""bench.py -- (T) FP64 throughput of the B200 engine on BASELINE.json's configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2]
  torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, NCCL)

A "step" is one pass of the hot path (contraction + singles/energy reduction) over
`tuples_per_step` consecutive tuples of this rank's group-and-sort list, on synthetic tensors
(counter-based generator) that live in HBM before the timed region.  Every step walks
new tuples, so each step reads GBs of slices that were not touched by the previous one (inputs
larger than L2: the ABPH store alone is 23 GB at c2).

  value       whole-job FP64 TFLOP/s = 12 No^3 (No+Nv) x tuples of all ranks / device time,
              device time = CUDA events on the engine's stream, max over ranks
  e2e         the same metric through the C-ABI with HOST tensors: every step ingests the
              pinned host tensors (H2D + re-tiling on the device), runs the step's tuples and
              reads the energy back -- what one Atrip::run(max_iterations = tuples_per_step) does
  roofline    contraction kernel: algorithmic FLOP / launch duration (events around the launch)
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
    # multi-GPU configs (sharded stores); host tensors of these sizes do not exist anywhere, so no e2e leg
    "c3": dict(No=64, Nv=640, scale=0.0005, tuples_per_step=16900, no_e2e=True,
               desc="No=64 Nv=640 FP64 random tensors (170 GB of stores)"),
    "c4": dict(No=100, Nv=1000, scale=0.0002, tuples_per_step=3900, no_e2e=True, min_gpus=8,
               desc="No=100 Nv=1000 FP64 random tensors (1.0 TB of stores over 8 GPUs)"),
    "c5": dict(No=32, Nv=1200, scale=0.0005, tuples_per_step=59200, no_e2e=True, min_gpus=4,
               desc="No=32 Nv=1200 FP64 random tensors, high Nv/No (0.5 TB of stores)"),
}
SEED = 12345


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


# ----------------------------------------------------------------------------- CPU reference arm
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
    busy, esum = 0.0, 0.0
    for abc in tuples:
        S = o.synth_tuple_slices(No, Nv, abc, seed=SEED, scale=scale)  # input generation: not timed
        t0 = time.perf_counter()
        if use_ref:  # the reference's own code: Equations.cxx doubles/singles/energy
            T = r.doubles(No, Nv, S, scratch)
            Z = r.singles(No, Nv, abc, tai, S, T)
            eps = float(epsa[abc[0]] + epsa[abc[1]] + epsa[abc[2]])
            same = (abc[0] == abc[1]) != (abc[1] == abc[2])
            e = (r.energy_same if same else r.energy_distinct)(eps, No, epsi, T, Z)
        else:
            T = o.doubles(No, Nv, S)
            Z = o.singles(No, Nv, abc, tai, S, T)
            eps = float(epsa[abc[0]] + epsa[abc[1]] + epsa[abc[2]])
            same = (abc[0] == abc[1]) != (abc[1] == abc[2])
            e = (o.energy_same if same else o.energy_distinct)(eps, No, epsi, T, Z)
        busy += time.perf_counter() - t0
        esum += e
    return busy, esum


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
    """n tuples spread over the whole list (same list the GPU walks)"""
    import numpy as np
    from atrip_b200 import capi
    allt = capi.host_tuples(capi.GROUP_AND_SORT, cfg["Nv"], 0, 1, pad=False)
    idx = np.linspace(0, len(allt) - 1, n).astype(np.int64)
    return allt[idx]


def est_cpu_tuples(cfg, cores, seconds):
    flops = 12.0 * cfg["No"] ** 3 * (cfg["No"] + cfg["Nv"])
    per_tuple = flops / 6.0e9  # ~6 GF/s/core measured for the reference dgemm path
    return max(cores, int(seconds / per_tuple) * cores)


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
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
                             "sample": f"{per_step} tuples/step spread over the group-and-sort list, reference "
                                       "doubles_contribution+singles_contribution+get_energy_* (dgemm path, wheel "
                                       "OpenBLAS, 1 thread per process), one process per core"},
            "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extrapolated_full_wall_s": flops_per_tuple * (cfg["Nv"] * (cfg["Nv"] + 1) * (cfg["Nv"] + 2) // 6 - cfg["Nv"])
            / (value * 1e12)}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def host_tensors(cfg, device):
    """pinned host tensors in CTF layout holding the synthetic inputs (for the e2e leg)"""
    import torch
    from atrip_b200 import capi
    No, Nv = cfg["No"], cfg["Nv"]
    sizes = {0: No, 1: Nv, 2: Nv * No, 3: Nv * Nv * No * No, 4: Nv * Nv * No * No, 5: No ** 3 * Nv, 6: Nv ** 3 * No}
    out = {}
    for tid, n in sizes.items():
        t = torch.empty(n, dtype=torch.float64).pin_memory()
        capi.synth_to_host(device, SEED, tid, cfg["scale"], 0, n, t.data_ptr())
        out[tid] = t
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--tuples-per-step", type=int, default=0)
    ap.add_argument("--replicate", action="store_true",
                    help="N>1: every GPU holds a full replica of the stores (no slice exchange)")
    ap.add_argument("--transport", type=int, default=0, choices=[0, 1, 2],
                    help="N>1 slice exchange: 1 NCCL send/recv, 2 P2P copy-engine pulls, 0 engine default")
    ap.add_argument("--field", default="real", choices=["real", "complex"],
                    help="complex: Atrip::run<Complex> instantiation (4x the FLOPs per tuple, Atrip.cxx:578-580); "
                         "device-resident synthetic stores only (no e2e / CPU legs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference" or os.environ.get("ATRIP_BENCH_ALLOW_SHORT"), "warmup >= 3"
    cfg = dict(CONFIGS[args.config], name=args.config)
    if args.tuples_per_step:
        cfg["tuples_per_step"] = args.tuples_per_step
    if cfg.get("no_e2e") or args.field == "complex":
        args.no_e2e = True
    if args.field == "complex":
        args.no_cpu = True
        assert args.impl == "ours", "the reference arm times the real (double) instantiation"
    assert args.impl == "reference" or args.gpus >= cfg.get("min_gpus", 1), \
        f"{args.config} needs at least {cfg.get('min_gpus')} GPUs (stores are sharded over the ranks)"
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
    # N > 1: every GPU stores the slices it owns (RankMap round robin) and fetches the rest of each
    # batch from its peers with ncclSend/ncclRecv on a side stream, one batch ahead of the compute
    sharded = world > 1 and not args.replicate
    eng = atrip_b200.Engine(No, Nv, device=local, rank=rank, nranks=world, resident=not sharded,
                            transport=args.transport,
                            field=capi.FIELD_COMPLEX if args.field == "complex" else capi.FIELD_REAL)
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
    dev_ms, launches, tuples_done, ck_ms, rk_ms, ck_n, energy = 0.0, 0, 0.0, 0.0, 0.0, 0, 0.0
    xbytes = xmsgs = 0.0
    for i in range(W, W + K):
        tm, tot = step(i)
        ex = eng.last_exchange()
        xbytes += ex["bytes"]
        xmsgs += ex["messages"]
        dev_ms += tm["total_ms"]
        launches += tm["contract_launches"] + tm["reduce_launches"]
        tuples_done += tot[2]
        ck_ms += tm["contract_ms"]
        rk_ms += tm["reduce_ms"]
        ck_n += 1
        energy += tot[0]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    value = flops_per_tuple * tuples_done / (dev_ms * 1e-3) / 1e12

    # contraction kernel roofline: mean launch duration (events around the launch, engine stream)
    batch_tuples = min(tps, eng.batch_tuples)
    contract_ms = ck_ms / ck_n
    achieved = flops_per_tuple * batch_tuples / (contract_ms * 1e-3) / 1e12 if contract_ms > 0 else None

    # ------------------------------------------------ e2e: host tensors through the C-ABI
    e2e = None
    if not args.no_e2e:
        # every rank pins its own copy of the host tensors: make sure the box has the memory
        need = 8 * (2 * Nv * Nv * No * No + No ** 3 * Nv + Nv ** 3 * No) * world
        try:
            import psutil
            if psutil.virtual_memory().available < 1.25 * need:
                args.no_e2e = True
                e2e = {"skipped": f"host memory: {need / 1e9:.0f} GB of pinned tensors needed for {world} ranks"}
        except ImportError:
            pass
    if not args.no_e2e:
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
        done = 0.0
        for i in range(W, W + K):
            done += e2e_step(i)[2]
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": flops_per_tuple * done / e2e_s / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": 16 * world, "ms_per_step": e2e_s / K * 1e3,
               "what": "per step: ingest all pinned host tensors (CTF layout) + run the step's tuples + read energy"}
        del host

    # ------------------------------------------------ CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        n = est_cpu_tuples(cfg, cores, args.cpu_seconds)
        tl = cpu_sample_tuples(cfg, n)
        with mp.get_context("spawn").Pool(cores) as pool:
            sec, n, _, kind = cpu_reference_step(cfg, tl, cores, pool)
        cpu = {"value": flops_per_tuple * n / sec / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": kind,
               "sample": f"{n} tuples spread over the list, reference doubles+singles+energy functions, "
                         f"{cores} single-threaded processes, {sec:.1f} s"}

    if rank == 0:
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(cfg["name"], {}).get("contract_dram_bytes_per_launch")
            except Exception:
                traffic = None
        total_tuples = Nv * (Nv + 1) * (Nv + 2) // 6 - Nv
        line = {
            "metric": "(T) FP64 TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128" if args.field == "complex" else "f64", "data": "synthetic",
            "config": {"workload": cfg["name"] + ": " + cfg["desc"] + (" [complex field]" if args.field == "complex" else ""),
                       "No": No, "Nv": Nv, "tuples_per_step": tps,
                       "tuples_per_step_all_ranks": tps * world, "distribution": "group_and_sort (GPU == node)",
                       "stores": ("sharded: owned slices + fetch cache prefetched one batch ahead on a side stream, "
                                  + {0: "engine default transport", 1: "ncclSend/ncclRecv",
                                     2: "P2P copy-engine pulls over NVLink"}[args.transport]
                                  if sharded else "replica on every GPU" if world > 1 else "single GPU"),
                       "l2_policy": "inputs larger than L2: every step walks new tuples (GBs of new slices)",
                       "seed": SEED, "scale": cfg["scale"]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * world,
            "roofline": {"bound": "tensor", "kernel": "contract_kernel (FP64 DMMA)", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if achieved and peak else None,
                         "traffic": traffic, "launch_ms": contract_ms, "tuples_per_launch": batch_tuples,
                         "peak_source": "FP64 DMMA ceiling measured live (atrip_b200_measure_dmma_peak); "
                                        "MEASURED_PEAKS.json has no FP64 figure"},
            "cpu_baseline": cpu,
            "wall_ms_per_step": wall_ms / K, "reduce_ms_per_launch": rk_ms / ck_n,
            "frac_of_fp64_tensor_peak": value / (peak * world) if peak else None,
            "extrapolated_full_wall_s": flops_per_tuple * total_tuples / (value * 1e12),
            "energy_partial": -energy,
            "exchange": {"rank0_recv_bytes_per_step": xbytes / K, "rank0_messages_per_step": xmsgs / K,
                         "rank0_recv_GBps": xbytes / (dev_ms * 1e-3) / 1e9 if dev_ms else None},
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
