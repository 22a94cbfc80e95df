"""CPU: host-side logic of bench.py -- the workload picked per GPU count (BASELINE.json configs), the
reference arm's tuple sample, the parity tuples, and the per-owner slice lists of the e2e leg against the
engine's own ownership map (replaces RankMap::find, RankMap.cxx:35-85)."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT

spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_workload_by_gpu_count_follows_baseline_configs():
    assert bench.pick_config(8) == "c4" and (bench.CONFIGS["c4"]["No"], bench.CONFIGS["c4"]["Nv"]) == (100, 1000)
    assert bench.pick_config(2) == "c3" and bench.pick_config(4) == "c3"
    assert (bench.CONFIGS["c3"]["No"], bench.CONFIGS["c3"]["Nv"]) == (64, 640)
    import torch
    if not torch.cuda.is_available():  # no GPU here: the one-GPU choice cannot see free memory
        assert bench.pick_config(1) == "c2"
    # BASELINE.json: c4 needs 8 GPUs (1.0 TB of stores), >= 2000 steady-state tuples per GPU and step (survey 8d)
    assert bench.CONFIGS["c4"]["min_gpus"] == 8 and bench.CONFIGS["c4"]["tuples_per_step"] >= 2000
    assert 1.0e12 < bench.store_bytes(bench.CONFIGS["c4"]) < 1.2e12
    assert 1.6e11 < bench.store_bytes(bench.CONFIGS["c3"]) < 1.8e11


@pytest.mark.parametrize("name", ["c2", "c3", "c4", "c5"])
def test_sample_and_parity_tuples_are_valid(name):
    cfg = bench.CONFIGS[name]
    t = bench.cpu_sample_tuples(cfg, 500)
    assert t.shape == (500, 3) and t.min() >= 0 and t.max() < cfg["Nv"]
    assert np.all(t[:, 0] <= t[:, 1]) and np.all(t[:, 1] <= t[:, 2]) and not np.any((t[:, 0] == t[:, 1]) & (t[:, 1] == t[:, 2]))
    pt = bench.parity_tuples(cfg)
    kinds = set()
    for a, b, c in pt:
        assert 0 <= a <= b <= c < cfg["Nv"] and not (a == b == c)
        kinds.add((a == b) != (b == c))
    assert kinds == {True, False}  # both energy kernels (get_energy_same / get_energy_distinct) are exercised


@pytest.mark.parametrize("world", [2, 4, 8])
def test_e2e_slice_lists_follow_the_ownership_map(world):
    """every slice a rank uploads in the e2e leg is one it holds, and together the ranks upload every slice the
    step's tuples read"""
    from atrip_b200 import capi
    Nv = 96
    tuples = capi.host_tuples(capi.GROUP_AND_SORT, Nv, rank=1, nranks=world)[100:400]
    tuples = tuples[tuples.any(axis=1)].astype(np.int64)
    got = {}
    for rank in range(world):
        need = bench.step_input_slices(capi, tuples, Nv, rank, world)
        for kind, xy in need.items():
            for x, y in xy:
                yy = int(y) if kind >= 200 else 0
                if kind == capi.TABIJ:  # feeds the hole rows of the ordered pairs (x,y) and (y,x): either one held
                    held = max(capi.local_slot(capi.VABCI, int(x), yy, Nv, rank, world),
                               capi.local_slot(capi.VABCI, yy, int(x), Nv, rank, world)) >= 0
                else:
                    held = capi.local_slot(kind, int(x), yy, Nv, rank, world) >= 0
                assert held, (kind, x, y, rank)
                got.setdefault(kind, set()).add((int(x), yy))
    want = {capi.TA: set(), capi.VIJKA: set(), capi.VABCI: set(), capi.TABIJ: set(), capi.VABIJ: set()}
    for a, b, c in tuples:
        for x in (a, b, c):
            want[capi.TA].add((int(x), 0))
            want[capi.VIJKA].add((int(x), 0))
        for y, z in ((b, c), (a, c), (c, b), (a, b), (c, a), (b, a)):
            want[capi.VABCI].add((int(y), int(z)))
            want[capi.TABIJ].add((int(min(y, z)), int(max(y, z))))
        for y, z in ((b, c), (a, c), (a, b)):
            want[capi.VABIJ].add((int(y), int(z)))
    for kind in want:
        assert want[kind] <= got.get(kind, set()), kind
