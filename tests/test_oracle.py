"""CPU: the oracle (oracle/atrip_oracle.c) against vectors produced by the reference itself
(tests/golden/reference_vectors.json, generator tests/golden/make_golden.py) and, when the
reference build oracle/_ref exists (this container), against the reference directly."""
import hashlib

import numpy as np
import pytest

from conftest import fh
from oracle.oracle import EPS_A, EPS_I, TAI, Reference


def test_synth_is_counter_based(oracle):
    a = oracle.fill(12345, 3, 0.1, 100)
    b = oracle.fill(12345, 3, 0.1, 10, first=50)
    assert np.array_equal(a[50:60], b)
    assert np.all(np.abs(a) <= 0.05)
    ei, ea = oracle.fill(1, 0, 0.1, 1000), oracle.fill(1, 1, 0.1, 1000)
    assert ei.min() >= -2.0 and ei.max() < -0.5 and ea.min() >= 0.5 and ea.max() < 4.0


def test_runs_match_reference_vectors(oracle, golden):
    for r in golden["runs"]:
        if r["No"] * r["Nv"] > 200:  # keep the CPU suite short; the big ones are GPU parity cases
            continue
        t = oracle.inputs(r["No"], r["Nv"], seed=r["seed"], scale=r["scale"], with_J=r["with_J"])
        e, ct = oracle.run(r["No"], r["Nv"], t)
        assert abs(e - fh(r["energy"])) <= 1e-12 * abs(e) + 1e-15, r
        assert abs(ct - fh(r["ct_energy"])) <= 1e-11 * max(abs(e), abs(ct)) + 1e-15, r


def test_tuples_match_reference_vectors(oracle, golden):
    for rec in golden["tuples"]:
        No, Nv = rec["No"], rec["Nv"]
        t = oracle.inputs(No, Nv, seed=rec["seed"], scale=rec["scale"])
        idx = [0, 1, No, No * No, No ** 3 // 2, No ** 3 - 1]
        for g in rec["tuples"]:
            e, _, T, Z = oracle.tuple_energy(No, Nv, t, tuple(g["abc"]), want_cubes=True)
            tmax = fh(g["Tabsmax"])
            assert abs(e - fh(g["energy"])) <= 1e-12 * abs(e)
            assert np.allclose(T[idx], [fh(x) for x in g["Tsample"]], rtol=0, atol=1e-13 * tmax)
            assert np.allclose(Z[idx], [fh(x) for x in g["Zsample"]], rtol=0, atol=1e-13 * tmax)
            assert abs(T.sum() - fh(g["Tsum"])) <= 1e-11 * tmax * No ** 1.5


def test_group_and_sort_matches_reference_vectors(oracle, golden):
    for rec in golden["distributions"]:
        for me, g in enumerate(rec["nodes"]):
            tl = oracle.group_and_sort(rec["n_nodes"], me, rec["Nv"])
            assert len(tl) == g["count"]
            assert hashlib.sha256(np.ascontiguousarray(tl.astype(np.uint64)).tobytes()).hexdigest() == g["sha256"]


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_against_reference_build(oracle):
    ref = Reference()
    No, Nv = 6, 14
    t = oracle.inputs(No, Nv, seed=31, scale=0.05, with_J=True)
    e, ct = oracle.run(No, Nv, t)
    er, ctr = ref.run(No, Nv, t)
    assert abs(e - er) <= 1e-13 * abs(er) and abs(ct - ctr) <= 1e-12 * abs(er)
    for abc in [(0, 1, 2), (3, 3, 5), (3, 5, 5), (11, 12, 13)]:
        S = oracle.tuple_slices(No, Nv, t, abc)
        T, Tr = oracle.doubles(No, Nv, S), ref.doubles(No, Nv, S)
        assert np.abs(T - Tr).max() <= 1e-14 * np.abs(Tr).max()
        Z, Zr = oracle.singles(No, Nv, abc, t[TAI], S, Tr), ref.singles(No, Nv, abc, t[TAI], S, Tr)
        assert np.array_equal(Z, Zr)
        eps = float(t[EPS_A][list(abc)].sum())
        assert oracle.energy_distinct(eps, No, t[EPS_I], Tr, Zr) == ref.energy_distinct(eps, No, t[EPS_I], Tr, Zr)
        assert oracle.energy_same(eps, No, t[EPS_I], Tr, Zr) == ref.energy_same(eps, No, t[EPS_I], Tr, Zr)
    for n in (1, 2, 3, 5):
        for me in range(n):
            assert np.array_equal(oracle.group_and_sort(n, me, 17), ref.group_and_sort(n, me, 17))
