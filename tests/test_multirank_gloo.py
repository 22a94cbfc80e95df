"""CPU, world_size 2 over gloo: the engine's N>1 host logic end to end without a GPU.

Each process plays one rank of the sharded engine: it holds only the slices it owns (numpy
stand-ins, one number per slice), walks its group-and-sort list in batches and runs the SAME
protocol as engine.cu: exchange_step / run_list -- request lists in the engine's wire format
travel one batch ahead of the data, data lands in one of two cache regions, the (mock) tuple
energies are summed and all-reduced.  Checks every slot against the slice it must hold and the
all-reduced sum against a single-rank evaluation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KA, KB, KV = 0, 1, 2


def slice_value(kind, gid):
    return float(gid * 4 + kind) + 0.5


def build_store(capi, Nv, r, n):
    from test_schedule import stores_of
    ids = stores_of(Nv, n)[r]
    return [np.array([slice_value(k, g) for g in ids[k]]) for k in range(3)]


def mock_energy(Nv, t):
    from test_schedule import wanted_ids
    return sum(slice_value(k, g) for k, ids in enumerate(wanted_ids(t, Nv)) for g in ids)


def worker(rank, world, port, Nv, batch, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from atrip_b200 import capi
    from test_schedule import wanted_ids
    store = build_store(capi, Nv, rank, world)
    owned = capi.shard_sizes(Nv, rank, world)
    tl = capi.host_tuples(capi.GROUP_AND_SORT, Nv, rank, world, pad=True)
    need = capi.cache_need(Nv, rank, world, tl, batch)
    cache = [np.full(2 * max(need[k], 1), np.nan) for k in range(3)]
    nb = (len(tl) + batch - 1) // batch
    cap = 1 + 36 * batch  # request_capacity_ints

    def plan(k):
        base = [owned[q] + (k & 1) * need[q] for q in range(3)]
        return capi.plan_batch(Nv, rank, world, tl[k * batch:(k + 1) * batch], base) + (base,)

    def encode(ranges, peer):
        buf = torch.zeros(cap, dtype=torch.int32)
        mine = [r for r in ranges.tolist() if r[0] == peer]
        buf[0] = len(mine)
        for i, (_, kind, src, cnt, _) in enumerate(mine):
            buf[1 + 3 * i], buf[2 + 3 * i], buf[3 + 3 * i] = kind, src, cnt
        return buf

    def exchange(k, mine, nxt, peer_req):
        """one exchange step; returns the peers' request lists for batch k+1"""
        ops, recvs, got = [], [], {}
        for p in range(world):
            if p == rank:
                continue
            if k >= 0:
                rq = peer_req[p]
                for i in range(int(rq[0])):
                    kind, src, cnt = (int(v) for v in rq[1 + 3 * i:4 + 3 * i])
                    assert src + cnt <= owned[kind]
                    ops.append(dist.isend(torch.from_numpy(store[kind][src:src + cnt].copy()), p))
                for peer, kind, src, cnt, dst in mine[1].tolist():
                    if peer == p:
                        buf = torch.empty(cnt, dtype=torch.float64)
                        ops.append(dist.irecv(buf, p))
                        recvs.append((kind, (k & 1) * need[kind] + dst, cnt, buf))
            if nxt is not None:
                ops.append(dist.isend(encode(nxt[1], p), p))
                got[p] = torch.zeros(cap, dtype=torch.int32)
                ops.append(dist.irecv(got[p], p))
        for o in ops:
            o.wait()
        for kind, dst, cnt, buf in recvs:
            cache[kind][dst:dst + cnt] = buf.numpy()
        return got

    plans = {0: plan(0)}
    peer_req = exchange(-1, None, plans[0], None)
    if nb > 1:
        plans[1] = plan(1)
    nxt_req = exchange(0, plans[0], plans.get(1), peer_req)
    esum, checked = 0.0, 0
    for k in range(nb):
        recs, ranges, base = plans[k]
        for t, rec in zip(tl[k * batch:(k + 1) * batch], recs):
            if rec[3]:
                continue
            got = (rec[4:7], rec[7:13], rec[13:16])
            for kind, want in enumerate(wanted_ids(t, Nv)):
                for slot, gid in zip(got[kind], want):
                    v = store[kind][slot] if slot < owned[kind] else cache[kind][slot - owned[kind]]
                    assert v == slice_value(kind, gid), (rank, k, t, kind, slot, gid, v)
                    esum += v
                    checked += 1
        if k + 1 < nb:
            if k + 2 < nb:
                plans[k + 2] = plan(k + 2)
            nxt_req = exchange(k + 1, plans[k + 1], plans.get(k + 2), nxt_req)
        plans.pop(k)
    tot = torch.tensor([esum, float(checked)], dtype=torch.float64)
    dist.all_reduce(tot)
    if rank == 0:
        result.put((float(tot[0]), int(tot[1])))
    dist.destroy_process_group()


@pytest.mark.parametrize("Nv,batch", [(14, 9), (11, 4)])
def test_two_ranks_exchange_and_allreduce(lib, oracle, Nv, batch):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, Nv, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    total, checked = q.get(timeout=10)
    allt = oracle.all_tuples(Nv)
    assert checked == 12 * len(allt)
    assert total == sum(mock_energy(Nv, t) for t in allt)
