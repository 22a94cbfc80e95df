"""GPU, >= 2 devices: the sharded engine (owned slices + fetch cache, both transports: NCCL
send/recv and P2P copy-engine pulls) against the reference vectors and the oracle.  One process per GPU, as in production; skipped on a single-GPU box
(the same host logic runs over gloo in tests/test_multirank_gloo.py).  Run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import sys

import numpy as np
import pytest

from conftest import fh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E_ABS, E_REL, CUBE_REL = 1e-10, 1e-12, 1e-13


def _worker(rank, world, q_id, q_out, case):
    sys.path.insert(0, ROOT)
    import atrip_b200
    from atrip_b200 import capi
    No, Nv, seed, scale, with_J, batch, host_tensors, debug, transport = case[:9]
    field = case[9] if len(case) > 9 else 0
    eng = atrip_b200.Engine(No, Nv, device=rank, rank=rank, nranks=world, with_J=with_J, batch_tuples=batch,
                            resident=False, transport=transport, field=field)
    if rank == 0:
        uid = capi.comm_unique_id()
        for _ in range(world - 1):
            q_id.put(uid)
    else:
        uid = q_id.get(timeout=120)
    eng.comm_init(uid)
    if host_tensors is None:
        eng.fill_synthetic(seed, scale)
    else:
        eng.load_all(*host_tensors)
    n = eng.build_tuples(capi.GROUP_AND_SORT)
    # walk the list in three uneven collective calls (exercises the bootstrap of each call)
    cuts = [0, n // 3, n // 3 + 1, n]
    e = ct = 0.0
    xb = 0.0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        a, b = eng.run(lo, hi - lo)
        e, ct = e + a, ct + b
        xb += eng.last_exchange()["bytes"]
    tot = eng.allreduce([e, ct])
    dbg = []
    for abc in debug:  # collective: every rank evaluates the same tuple, fetching what it lacks
        ge, T, Z = eng.tuple_debug(*abc)
        dbg.append((ge, T, Z))
    q_out.put((rank, float(tot[0]), float(tot[1]), xb, dbg if rank == world - 1 else None))
    eng.close()


def run_case(world, case):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, q_id, q_out, case)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted((q_out.get(timeout=600) for _ in range(world)), key=lambda x: x[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    return res


@pytest.mark.parametrize("transport", [1, 2], ids=["nccl", "p2p"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_runs_match_reference_vectors(oracle, golden, world, transport):
    """whole-run energies of the reference (np=1) reproduced by `world` GPUs with sharded stores,
    device fill, small batches so that many exchange steps happen"""
    for r in [golden["runs"][i] for i in (1, 4, 5)]:
        res = run_case(world, (r["No"], r["Nv"], r["seed"], r["scale"], r["with_J"], 37, None, [], transport))
        for rank, e, ct, xb, _ in res:
            assert abs(-e - fh(r["energy"])) <= E_ABS and abs(-e - fh(r["energy"])) <= E_REL * abs(e), (r, rank, -e)
            ref_ct = fh(r["ct_energy"])
            assert abs(-ct - ref_ct) <= E_ABS and abs(-ct - ref_ct) <= 1e-11 * max(abs(ref_ct), abs(e))
            assert xb > 0, "no slices travelled: the sharded path was not exercised"


@pytest.mark.parametrize("transport", [1, 2], ids=["nccl", "p2p"])
def test_sharded_complex_runs_match_reference_vectors(golden, transport):
    """F = Complex with sharded stores: a complex AX slice carries both K-doubled variants and travels
    as one slice; whole-run energies of the reference's run<Complex> incl. (cT) on 2 GPUs"""
    for r in [golden["complex_runs"][i] for i in (1, 2)]:
        res = run_case(2, (r["No"], r["Nv"], r["seed"], r["scale"], r["with_J"], 37, None, [], transport, 1))
        for rank, e, ct, xb, _ in res:
            assert abs(-e - fh(r["energy"])) <= E_ABS and abs(-e - fh(r["energy"])) <= E_REL * abs(e), (r, rank, -e)
            ref_ct = fh(r["ct_energy"])
            assert abs(-ct - ref_ct) <= E_ABS and abs(-ct - ref_ct) <= 1e-11 * max(abs(ref_ct), abs(e))
            assert xb > 0


@pytest.mark.parametrize("transport", [1, 2], ids=["nccl", "p2p"])
def test_sharded_ingest_and_tuple_debug_match_oracle(oracle, transport):
    """host tensors ingested shard by shard; collective tuple_debug (remote slices of all three
    kinds) against the oracle's Tijk / Zijk / energy; default batch size"""
    from oracle.oracle import EPS_A, EPS_I, TABIJ, TAI, VABCI, VABIJ, VIJKA
    No, Nv, seed, scale = 9, 21, 777, 0.05
    t = oracle.inputs(No, Nv, seed=seed, scale=scale)
    host = (t[EPS_I], t[EPS_A], t[TAI], t[TABIJ], t[VABIJ], t[VIJKA], t[VABCI])
    debug = [(0, 1, 2), (1, 3, 5), (2, 2, 7), (4, 9, 9), (0, 2, 4), (17, 19, 20)]
    want, _ = oracle.run(No, Nv, t)
    res = run_case(2, (No, Nv, seed, scale, False, 0, host, debug, transport))
    for rank, e, ct, xb, dbg in res:
        assert abs(-e - want) <= E_ABS and abs(-e - want) <= E_REL * abs(want)
        if dbg is None:
            continue
        for abc, (ge, T, Z) in zip(debug, dbg):
            oe, _, oT, oZ = oracle.tuple_energy(No, Nv, t, abc, want_cubes=True)
            assert abs(ge - oe) <= E_REL * abs(oe), abc
            assert np.abs(T - oT).max() <= CUBE_REL * np.abs(oT).max(), abc
            assert np.abs(Z - oZ).max() <= CUBE_REL * np.abs(oZ).max(), abc


def _big_worker(rank, world, q_id, q_out, No, Nv, seed, scale, tuples):
    sys.path.insert(0, ROOT)
    import atrip_b200
    from atrip_b200 import capi
    eng = atrip_b200.Engine(No, Nv, device=rank, rank=rank, nranks=world, resident=False)
    if rank == 0:
        uid = capi.comm_unique_id()
        for _ in range(world - 1):
            q_id.put(uid)
    else:
        uid = q_id.get(timeout=300)
    eng.comm_init(uid)
    eng.fill_synthetic(seed, scale)
    out = [eng.tuple_debug(*abc, cubes=False)[0] for abc in tuples]  # remote slices pulled from the owners
    q_out.put((rank, out))
    eng.close()  # collective: waits for the peers that may still read this rank's stores


def test_c4_tuples_on_8_gpus_match_reference_functions(oracle):
    """BASELINE config 4 (No=100, Nv=1000, 1 TB of stores over 8 GPUs): tuple energies from the
    sharded engine against the reference's doubles/singles/energy functions evaluated on the
    same synthetic slices (generated slice by slice on the host, no full tensors)"""
    import torch
    import torch.multiprocessing as mp
    from oracle.oracle import EPS_A, EPS_I, TAI, Reference
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs")
    No, Nv, seed, scale, world = 100, 1000, 12345, 0.0002, 8
    tuples = [(3, 500, 997), (17, 17, 640), (250, 251, 252)]
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_big_worker, args=(r, world, q_id, q_out, No, Nv, seed, scale, tuples)) for r in range(world)]
    for p in procs:
        p.start()
    # meanwhile: the same tuples on the host
    ref = Reference() if Reference.available() else None
    epsi, epsa = oracle.fill(seed, EPS_I, scale, No), oracle.fill(seed, EPS_A, scale, Nv)
    tai = oracle.fill(seed, TAI, scale, No * Nv)
    want = []
    for abc in tuples:
        S = oracle.synth_tuple_slices(No, Nv, abc, seed=seed, scale=scale)
        f = ref if ref is not None else oracle
        T = f.doubles(No, Nv, S, (np.empty(No ** 3), np.empty(No ** 3))) if ref is not None else f.doubles(No, Nv, S)
        Z = f.singles(No, Nv, abc, tai, S, T)
        eps = float(epsa[abc[0]] + epsa[abc[1]] + epsa[abc[2]])
        same = (abc[0] == abc[1]) != (abc[1] == abc[2])
        want.append((f.energy_same if same else f.energy_distinct)(eps, No, epsi, T, Z))
    try:
        res = [q_out.get(timeout=900) for _ in range(world)]
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    for rank, got in res:
        for abc, g, w in zip(tuples, got, want):
            assert abs(g - w) <= E_REL * abs(w), (rank, abc, g, w)
