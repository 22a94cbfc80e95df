"""CPU: the F = Complex oracle (oracle/atrip_oracle_z.c) against vectors produced by the reference's
own run<Complex> / L1 instantiations (tests/golden/reference_vectors.json: complex_runs,
complex_tuples) and, when oracle/_ref exists (this container), against the reference directly."""
import numpy as np
import pytest

from conftest import fh
from oracle.oracle import EPS_A, EPS_I, TAI, Reference


def zh(p):
    return complex(fh(p[0]), fh(p[1]))


def test_complex_synth_layout(oracle):
    z = oracle.fill_z(12345, 3, 0.1, 64)
    d = oracle.fill(12345, 3, 0.1, 128)
    assert np.array_equal(z.real, d[0::2]) and np.array_equal(z.imag, d[1::2])
    assert np.array_equal(oracle.fill_z(12345, 3, 0.1, 10, first=20), z[20:30])
    e = oracle.fill_z(12345, 0, 0.1, 16)  # eps: the real case's values, zero imaginary part
    assert np.array_equal(e.real, oracle.fill(12345, 0, 0.1, 16)) and not e.imag.any()


def test_complex_runs_match_reference_vectors(oracle, golden):
    for r in golden["complex_runs"]:
        if r["No"] * r["Nv"] > 100:  # CPU suite stays short; the larger ones are GPU parity cases
            continue
        t = oracle.inputs_z(r["No"], r["Nv"], seed=r["seed"], scale=r["scale"], with_J=r["with_J"])
        e, ct = oracle.run_z(r["No"], r["Nv"], t)
        assert abs(e - fh(r["energy"])) <= 1e-12 * abs(e) + 1e-15, r
        assert abs(ct - fh(r["ct_energy"])) <= 1e-11 * max(abs(e), abs(ct)) + 1e-15, r


def test_complex_tuples_match_reference_vectors(oracle, golden):
    for rec in golden["complex_tuples"]:
        No, Nv = rec["No"], rec["Nv"]
        if No > 16:
            continue
        t = oracle.inputs_z(No, Nv, seed=rec["seed"], scale=rec["scale"])
        idx = [0, 1, No, No * No, No ** 3 // 2, No ** 3 - 1]
        for g in rec["tuples"]:
            e, _, T, Z = oracle.tuple_energy_z(No, Nv, t, tuple(g["abc"]), want_cubes=True)
            tmax = fh(g["Tabsmax"])
            assert abs(e - fh(g["energy"])) <= 1e-12 * abs(e)
            assert np.abs(T[idx] - np.array([zh(x) for x in g["Tsample"]])).max() <= 1e-13 * tmax
            assert np.abs(Z[idx] - np.array([zh(x) for x in g["Zsample"]])).max() <= 1e-13 * tmax
            assert abs(T.sum() - zh(g["Tsum"])) <= 1e-11 * tmax * No ** 1.5


def test_complex_reduces_to_real_for_real_inputs(oracle):
    """zero imaginary parts: the complex path must give the real path's numbers"""
    No, Nv = 5, 9
    tr = oracle.inputs(No, Nv, seed=3, scale=0.1, with_J=True)
    tz = {k: v.astype(np.complex128) for k, v in tr.items()}
    er, ctr = oracle.run(No, Nv, tr)
    ez, ctz = oracle.run_z(No, Nv, tz)
    assert abs(er - ez) <= 1e-13 * abs(er) and abs(ctr - ctz) <= 1e-12 * max(abs(er), abs(ctr))


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_complex_oracle_against_reference_build(oracle):
    ref = Reference()
    if not ref.has_complex:
        pytest.skip("prebuilt oracle/_ref predates the complex entry points")
    No, Nv = 6, 12
    t = oracle.inputs_z(No, Nv, seed=31, scale=0.05, with_J=True)
    e, ct = oracle.run_z(No, Nv, t)
    er, ctr = ref.run_z(No, Nv, t)
    assert abs(e - er) <= 1e-12 * abs(er) and abs(ct - ctr) <= 1e-11 * max(abs(er), abs(ctr))
    for abc in [(0, 1, 2), (2, 2, 9), (3, 11, 11), (9, 10, 11)]:
        S = oracle.tuple_slices_z(No, Nv, t, abc)
        T1, T2 = oracle.doubles_z(No, Nv, S), ref.doubles_z(No, Nv, S)
        assert np.abs(T1 - T2).max() <= 1e-13 * np.abs(T2).max()
        Z1, Z2 = oracle.singles_z(No, Nv, abc, t[TAI], S, T1), ref.singles_z(No, Nv, abc, t[TAI], S, T2)
        assert np.abs(Z1 - Z2).max() <= 1e-13 * np.abs(Z2).max()
        eps = float((t[EPS_A][abc[0]] + t[EPS_A][abc[1]] + t[EPS_A][abc[2]]).real)
        same = (abc[0] == abc[1]) != (abc[1] == abc[2])
        e1 = (oracle.energy_same_z if same else oracle.energy_distinct_z)(eps, No, t[EPS_I], T2, Z2)
        e2 = (ref.energy_same_z if same else ref.energy_distinct_z)(eps, No, t[EPS_I], T2, Z2)
        assert abs(e1 - e2) <= 1e-13 * abs(e2)
