"""GPU: per-slice ingestion (atrip_b200_upload_slices / atrip_b200_read_slices), the path a rank of a
multi-rank Atrip::run takes (reference SliceUnion<F>::init, SliceUnion.cxx:305-332: every rank slices
only the sources it owns).  Slices in the reference's slice layout are cut out of the oracle's host
tensors; the stores they build must be bit-identical to the bulk ingest of the full tensors."""
import sys

import numpy as np
import pytest

from conftest import ROOT, fh
from oracle.oracle import EPS_A, EPS_I, JABCI, JIJKA, TABIJ, TAI, VABCI, VABIJ, VIJKA

pytestmark = pytest.mark.gpu
E_ABS, E_REL = 1e-10, 1e-12


def host_slices(capi, t, No, Nv, kind, xy, cplx=False):
    """the slices CTF::slice would yield (Unions.hpp:96-112, 134-151, 177-196, 219-236, 258-277)"""
    dt = np.complex128 if cplx else np.float64
    T4 = lambda a, shape: np.asarray(a, dtype=dt).reshape(shape, order="F")
    out = []
    for x, y in xy:
        if kind == capi.TA:
            s = T4(t[TABIJ], (Nv, Nv, No, No))[x]
        elif kind in (capi.VIJKA, capi.JIJKA):
            s = T4(t[VIJKA if kind == capi.VIJKA else JIJKA], (No, No, No, Nv))[..., x]
        elif kind in (capi.VABCI, capi.JABCI):
            s = T4(t[VABCI if kind == capi.VABCI else JABCI], (Nv, Nv, Nv, No))[x, y]
        elif kind == capi.TABIJ:
            s = T4(t[TABIJ], (Nv, Nv, No, No))[x, y]
        else:
            s = T4(t[VABIJ], (Nv, Nv, No, No))[x, y]
        out.append(np.ravel(s, order="F"))
    return np.concatenate(out) if out else np.zeros(0, dtype=dt)


def upload_everything(capi, eng, t, No, Nv, rank, world, with_J=False, cplx=False):
    kinds = [capi.TA, capi.VIJKA, capi.VABCI, capi.TABIJ, capi.VABIJ] + ([capi.JIJKA, capi.JABCI] if with_J else [])
    for kind in kinds:
        xy = capi.owned_slices(kind, Nv, rank, world)
        eng.upload_slices(kind, xy, host_slices(capi, t, No, Nv, kind, xy.tolist(), cplx))


@pytest.mark.parametrize("No,Nv,seed,scale,with_J", [(5, 11, 12345, 0.05, False), (7, 13, 99, 0.05, True),
                                                       (10, 24, 3, 0.02, False)])
def test_slice_upload_equals_bulk_ingest(oracle, golden, No, Nv, seed, scale, with_J):
    import atrip_b200
    from atrip_b200 import capi
    t = oracle.inputs(No, Nv, seed=seed, scale=scale, with_J=with_J)
    a = atrip_b200.Engine(No, Nv, with_J=with_J)
    a.load_all(t[EPS_I], t[EPS_A], t[TAI], t[TABIJ], t[VABIJ], t[VIJKA], t[VABCI], t.get(JIJKA), t.get(JABCI))
    b = atrip_b200.Engine(No, Nv, with_J=with_J)
    b.set_epsilon(t[EPS_I], t[EPS_A])
    b.set_Tai(t[TAI])
    upload_everything(capi, b, t, No, Nv, 0, 1, with_J)
    for eng in (a, b):
        eng.build_tuples(capi.GROUP_AND_SORT)
    assert a.run() == b.run()  # bit-identical stores -> bit-identical energies
    # read-back is the exact inverse of the upload, for every kind, batched and one by one
    for kind in (capi.TA, capi.VIJKA, capi.VABCI, capi.TABIJ, capi.VABIJ) + ((capi.JIJKA, capi.JABCI) if with_J else ()):
        xy = capi.owned_slices(kind, Nv, 0, 1)[::3]
        want = host_slices(capi, t, No, Nv, kind, xy.tolist())
        assert np.array_equal(b.read_slices(kind, xy), want), kind
        assert np.array_equal(a.read_slices(kind, xy), want), kind
    one = host_slices(capi, t, No, Nv, capi.VABCI, [(2, 1)])
    b.upload_slice(capi.VABCI, 2, 1, one * 2.0)  # single-slice entry point: overwrite one source
    assert np.array_equal(b.read_slices(capi.VABCI, [(2, 1)]), one * 2.0)
    with pytest.raises(capi.EngineError):
        b.upload_slices(capi.TABIJ, [(3, 1)], np.zeros(No * No))  # x <= y only
    a.close()
    b.close()


def test_slice_upload_complex_field(oracle, golden):
    import atrip_b200
    from atrip_b200 import capi
    r = golden["complex_runs"][3]
    No, Nv = r["No"], r["Nv"]
    t = oracle.inputs_z(No, Nv, seed=r["seed"], scale=r["scale"], with_J=r["with_J"])
    eng = atrip_b200.Engine(No, Nv, with_J=r["with_J"], field=capi.FIELD_COMPLEX)
    eng.set_epsilon(t[EPS_I], t[EPS_A])
    eng.set_Tai(t[TAI])
    upload_everything(capi, eng, t, No, Nv, 0, 1, r["with_J"], cplx=True)
    eng.build_tuples(capi.GROUP_AND_SORT)
    e, ct = eng.run()
    assert abs(-e - fh(r["energy"])) <= E_ABS and abs(-e - fh(r["energy"])) <= E_REL * abs(e)
    assert abs(-ct - fh(r["ct_energy"])) <= 1e-11 * max(abs(ct), abs(e))
    for kind in (capi.TA, capi.VIJKA, capi.VABCI, capi.TABIJ, capi.VABIJ):
        xy = capi.owned_slices(kind, Nv, 0, 1)[::5]
        assert np.array_equal(eng.read_slices(kind, xy), host_slices(capi, t, No, Nv, kind, xy.tolist(), True)), kind
    eng.close()


def _worker(rank, world, q_id, q_out, case):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, ROOT + "/tests")
    import atrip_b200
    from atrip_b200 import capi
    from oracle.oracle import Oracle
    No, Nv, seed, scale, with_J, transport = case
    t = Oracle().inputs(No, Nv, seed=seed, scale=scale, with_J=with_J)
    eng = atrip_b200.Engine(No, Nv, device=rank, rank=rank, nranks=world, with_J=with_J, batch_tuples=41,
                            resident=False, transport=transport)
    if rank == 0:
        uid = capi.comm_unique_id()
        for _ in range(world - 1):
            q_id.put(uid)
    else:
        uid = q_id.get(timeout=120)
    eng.comm_init(uid)
    eng.set_epsilon(t[EPS_I], t[EPS_A])
    eng.set_Tai(t[TAI])
    upload_everything(capi, eng, t, No, Nv, rank, world, with_J)  # only what this rank owns
    n = eng.build_tuples(capi.GROUP_AND_SORT)
    e1 = eng.run(0, n)
    upload_everything(capi, eng, t, No, Nv, rank, world, with_J)  # a collective re-upload after a run
    e2 = eng.run(0, n)
    tot = eng.allreduce([e1[0], e1[1], e2[0], e2[1]])
    # TABIJ slices held only through the second index come back from the (y,x)' hole rows
    xy = capi.owned_slices(capi.TABIJ, Nv, rank, world)
    back = eng.read_slices(capi.TABIJ, xy)
    q_out.put((rank, [float(v) for v in tot], bool(np.array_equal(back, host_slices(capi, t, No, Nv, capi.TABIJ, xy.tolist())))))
    eng.close()


@pytest.mark.parametrize("transport", [1, 2], ids=["nccl", "p2p"])
@pytest.mark.parametrize("world", [2, 4])
def test_per_owner_upload_on_sharded_stores(golden, world, transport):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    for r in [golden["runs"][i] for i in (4, 5)]:
        ctx = mp.get_context("spawn")
        q_id, q_out = ctx.Queue(), ctx.Queue()
        case = (r["No"], r["Nv"], r["seed"], r["scale"], r["with_J"], transport)
        procs = [ctx.Process(target=_worker, args=(k, world, q_id, q_out, case)) for k in range(world)]
        for p in procs:
            p.start()
        try:
            res = [q_out.get(timeout=600) for _ in range(world)]
            for p in procs:
                p.join(timeout=60)
                assert p.exitcode == 0
        finally:
            for p in procs:
                if p.is_alive():
                    p.kill()
        for rank, tot, tab_ok in res:
            assert tab_ok, rank
            for e, ct in ((tot[0], tot[1]), (tot[2], tot[3])):
                assert abs(-e - fh(r["energy"])) <= E_ABS and abs(-e - fh(r["energy"])) <= E_REL * abs(e), (r, rank, -e)
                assert abs(-ct - fh(r["ct_energy"])) <= 1e-11 * max(abs(ct), abs(e))
