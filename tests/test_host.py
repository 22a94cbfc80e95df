"""CPU: host-side logic of the product and the shape of its C-ABI (no compute calls)."""
import hashlib
import os
import re

import numpy as np
import pytest

from atrip_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "atrip_b200.h")).read()
    declared = set(re.findall(r"\b(atrip_b200_[a-zA-Z0-9_]+)\s*\(", hdr))
    declared -= {"atrip_b200_ctx", "atrip_b200_config"}
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.atrip_b200_version()


def test_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        return
    try:
        capi.Engine(4, 8)
    except capi.EngineError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("engine creation must fail without a CUDA device")


def test_group_and_sort_matches_reference_vectors(lib, golden):
    for rec in golden["distributions"]:
        counts = []
        for me, g in enumerate(rec["nodes"]):
            tl = capi.host_tuples(capi.GROUP_AND_SORT, rec["Nv"], me, rec["n_nodes"], pad=False)
            assert len(tl) == g["count"]
            assert hashlib.sha256(np.ascontiguousarray(tl).tobytes()).hexdigest() == g["sha256"]
            counts.append(len(tl))
        # padded lists: same length on every rank, fakes only at the end (Tuples.cxx:346-377)
        for me in range(rec["n_nodes"]):
            tl = capi.host_tuples(capi.GROUP_AND_SORT, rec["Nv"], me, rec["n_nodes"], pad=True)
            assert len(tl) == max(counts)
            assert np.all(tl[counts[me]:] == 0) and not np.any(np.all(tl[:counts[me]] == 0, axis=1))


def test_group_and_sort_partitions_all_tuples(lib, oracle):
    for Nv, n in [(9, 2), (14, 3), (25, 4), (32, 8), (7, 8)]:
        parts = [capi.host_tuples(capi.GROUP_AND_SORT, Nv, r, n, pad=False) for r in range(n)]
        allt = np.concatenate(parts)
        assert len(allt) == oracle.n_tuples(Nv)
        assert len({tuple(t) for t in allt.tolist()}) == len(allt)
        assert np.all(allt[:, 0] <= allt[:, 1]) and np.all(allt[:, 1] <= allt[:, 2])
        # every tuple has at least one home index on its rank (atrip.org:1938-1942)
        for r, p in enumerate(parts):
            assert np.all(np.any(p % n == r, axis=1))
        # single rank degenerates to the lexicographic list
    assert np.array_equal(capi.host_tuples(capi.GROUP_AND_SORT, 11, 0, 1), oracle.all_tuples(11))


def test_naive_distribution(lib, oracle):
    Nv, n = 10, 3
    allt = oracle.all_tuples(Nv)
    per = -(-len(allt) // n)
    for r in range(n):
        tl = capi.host_tuples(capi.NAIVE, Nv, r, n, pad=True)
        assert len(tl) == per
        want = allt[per * r: per * (r + 1)]
        assert np.array_equal(tl[:len(want)], want) and np.all(tl[len(want):] == 0)


def test_slice_owner_matches_rankmap(lib, oracle):
    # single-index slices: RankMap round robin (RankMap.cxx:43-82) for any Nv; pair slices: the
    # reference's (x + y Nv) % n whenever Nv % n == 0 (every BASELINE config), else the pair
    # stays with its first index (deliberate, SURVEY.md 8e: keeps the co-location for any Nv)
    for Nv, n in [(40, 8), (36, 4), (37, 8), (10, 3)]:
        for x in range(0, Nv, 5):
            assert capi.slice_owner(capi.TA, x, 0, Nv, n) == oracle.L.oracle_owner_single(x, n)
            assert capi.slice_owner(capi.VIJKA, x, 0, Nv, n) == oracle.L.oracle_owner_single(x, n)
            for y in range(0, Nv, 7):
                want = oracle.L.oracle_owner_pair(x, y, Nv, n) if Nv % n == 0 else x % n
                assert capi.slice_owner(capi.VABCI, x, y, Nv, n) == want
                assert capi.slice_owner(capi.TABIJ, x, y, Nv, n) == want


def test_ring_release_is_ordered_after_all_dmmas_in_sass(lib):
    """every contract_kernel instantiation hands a TMA ring stage back only after all DMMAs that
    consume its fragment loads (tools/check_sass_order.py; profiles/r01_ring_release_race.txt)"""
    import shutil
    import sys
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import check_sass_order
    from atrip_b200 import capi
    seen, bad = check_sass_order.check(capi.lib_path())
    assert seen >= 30 and not bad, bad


def test_contraction_plan_is_valid_for_every_No(lib):
    """the tile planner (engine.cu: plan_contraction, exposed as atrip_b200_host_plan) must return
    a launchable plan for every supported No: the row tile fits the rows a stage reserves and the
    TMA box limits (<= 256 per dimension), the ring fits the 227 KB of shared memory with at least
    three stages, and the tiles cover the whole No^2 x No class matrix"""
    variants = set()
    for No in range(1, 257):
        p = capi.host_plan(No)
        MI, NI, nw, tu, tv = p["MI"], p["NI"], p["warps"], p["tu"], p["tv"]
        assert p["arows"] == nw * MI * 8 and tu * tv <= p["arows"], (No, p)
        assert 1 <= tu <= min(No, 256) and 1 <= tv <= min(No, 256), (No, p)
        utiles, vtiles = -(-No // tu), -(-No // tv)
        assert p["row_tiles"] == utiles * vtiles and p["col_tiles"] * NI * 8 >= No, (No, p)
        assert 3 <= p["stages"] <= 12 and p["smem"] <= 232448, (No, p)
        assert p["smem"] == p["stages"] * (p["arows"] + NI * 8) * 128 + 1024, (No, p)
        assert 4 <= nw and (nw + 1) * 32 <= 512 and 0.0 < p["useful"] <= 1.0 + 1e-9, (No, p)
        variants.add((MI, NI))
    assert len(variants) >= 10  # the planner really uses the compiled family
    # the bench configs keep the plans the measurements in profiles/ were taken with
    assert [(capi.host_plan(n)["MI"], capi.host_plan(n)["NI"]) for n in (40, 64, 100, 32)] == \
        [(5, 5), (4, 8), (2, 13), (4, 4)]
    with pytest.raises(capi.EngineError):
        capi.host_plan(257)


def test_sass_order_checker_flags_a_hoisted_release(tmp_path, monkeypatch):
    """the checker itself: a synthetic listing with the arrive between the last LDS and the DMMAs that consume it
    (the round-1 race) is reported, the correct order is accepted"""
    import importlib.util
    import subprocess as sp
    spec = importlib.util.spec_from_file_location("check_sass_order", os.path.join(ROOT, "tools", "check_sass_order.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    def listing(body):
        lines = ["\t\tFunction : _ZN2ab15contract_kernelILi2ELi2ELi512ELi0EEEvNS_12ContractMapsENS_14ContractParamsE"]
        for n, ins in enumerate(body):
            lines.append(f"        /*{16 * n:04x}*/                   {ins} ;")
        return "\n".join(lines) + "\n"

    good = ["LDS.64 R2, [R0]", "DMMA.8x8x4 R4, R2, R2, R4", "MEMBAR.ALL.CTA", "FENCE.VIEW.ASYNC.S", "WARPSYNC.ALL",
            "@!P0 SYNCS.ARRIVE.TRANS64.A1T0 RZ, [UR4], RZ"]
    bad = ["LDS.64 R2, [R0]", "WARPSYNC.ALL", "@!P0 SYNCS.ARRIVE.TRANS64.A1T0 RZ, [UR4], RZ", "DMMA.8x8x4 R4, R2, R2, R4"]
    unfenced = ["LDS.64 R2, [R0]", "DMMA.8x8x4 R4, R2, R2, R4", "WARPSYNC.ALL", "@!P0 SYNCS.ARRIVE.TRANS64.A1T0 RZ, [UR4], RZ"]
    for body, nbad in ((good, 0), (bad, 1), (unfenced, 1)):
        class R:
            stdout = listing(body)
        monkeypatch.setattr(sp, "run", lambda *a, **k: R)
        seen, flagged = mod.check("unused.so")
        assert seen == 1 and len(flagged) == nbad, (body, flagged)


def test_c_abi_header_is_plain_c_and_links(lib, tmp_path):
    """the boundary is a C ABI: include/atrip_b200.h compiles as C99 (no C++ constructs, no torch types) and a C
    program linked against the library reaches the host-only entry points (no GPU: create must fail loudly)"""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "atrip_b200.h"
int main(void) {
  atrip_b200_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.No = 4; cfg.Nv = 8; cfg.nranks = 1; cfg.resident = 1;
  atrip_b200_ctx *ctx = 0;
  int rc = atrip_b200_create(&ctx, &cfg);
  int64_t plan[11];
  int prc = atrip_b200_host_plan(40, 0, plan);
  int64_t n = atrip_b200_host_tuples(1, 8, 0, 1, 1, 0, 0);
  printf("version %s create %d (%s) plan %d %lld %lld tuples %lld owner %d\n", atrip_b200_version(), rc,
         rc ? atrip_b200_last_error() : "ok", prc, (long long)plan[0], (long long)plan[1], (long long)n,
         (int)atrip_b200_host_slice_owner(200, 5, 2, 8, 4));
  if (ctx) atrip_b200_destroy(ctx);
  return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "atrip_b200", "csrc")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           "-o", str(exe), str(src), "-L" + libdir, "-latrip_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([str(exe)], text=True)
    assert "version atrip_b200" in out and " plan 0 5 5 " in out and " tuples 112 " in out and out.strip().endswith("owner 1"), out
    import torch
    if not torch.cuda.is_available():
        assert " create 1 (" in out and "no CPU fallback" in out, out


@pytest.mark.parametrize("Nv,n", [(1, 1), (2, 1), (2, 4), (3, 8), (5, 8), (9, 2)])
def test_tiny_and_ragged_tuple_lists(lib, Nv, n):
    """edge cases of the distribution: fewer virtual orbitals than ranks, a single orbital (no tuple at all), lists
    padded with FAKE_TUPLE to the longest rank's length; every tuple a<=b<=c (not all equal) appears exactly once"""
    from atrip_b200 import capi
    seen, lengths = set(), set()
    for r in range(n):
        tl = capi.host_tuples(capi.GROUP_AND_SORT, Nv, r, n, pad=True)
        lengths.add(len(tl))
        for a, b, c in tl[tl.any(axis=1)].tolist():
            assert a <= b <= c < Nv and not (a == b == c) and (a, b, c) not in seen
            seen.add((a, b, c))
    assert len(lengths) == 1 and len(seen) == Nv * (Nv + 1) * (Nv + 2) // 6 - Nv


def test_plan_rejects_unsupported_No(lib):
    from atrip_b200 import capi
    for bad in (0, 257, -3):
        with pytest.raises(Exception):
            capi.host_plan(bad)
    assert capi.host_plan(256)["useful"] > 0.99 and capi.host_plan(1)["stages"] >= 3
