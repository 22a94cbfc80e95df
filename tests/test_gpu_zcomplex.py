"""GPU: the F = std::complex<double> instantiation (reference Atrip.cxx:1136) through the C-ABI
(atrip_b200_config.field = 1) and through atrip::Atrip::run<Complex>, against vectors produced by
the reference's own run<Complex> / doubles/singles/energy<Complex> (tests/golden:
complex_runs, complex_tuples) and against the complex oracle on the same seeded inputs.
Tolerances as in test_gpu_parity.py."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, fh
from oracle.oracle import EPS_A, EPS_I, JABCI, JIJKA, TABIJ, TAI, VABCI, VABIJ, VIJKA

pytestmark = pytest.mark.gpu

E_ABS, E_REL, CUBE_REL = 1e-10, 1e-12, 1e-13


@pytest.fixture(scope="module")
def ab():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import atrip_b200
    from atrip_b200 import capi
    assert capi.load_library() is not None
    return atrip_b200


def zh(p):
    return complex(fh(p[0]), fh(p[1]))


def engine_from_host(ab, o, No, Nv, seed, scale, with_J=False, **kw):
    from atrip_b200 import capi
    t = o.inputs_z(No, Nv, seed=seed, scale=scale, with_J=with_J)
    eng = ab.Engine(No, Nv, with_J=with_J, field=capi.FIELD_COMPLEX, **kw)
    eng.load_all(t[EPS_I], t[EPS_A], t[TAI], t[TABIJ], t[VABIJ], t[VIJKA], t[VABCI], t.get(JIJKA), t.get(JABCI))
    return eng, t


def close_energy(e, ref):
    return abs(e - ref) <= E_ABS and abs(e - ref) <= E_REL * abs(ref)


def test_complex_slices_fill_and_ingest(ab, oracle):
    """the K-doubled stores hold exactly the reference's complex slices, from either source"""
    from atrip_b200 import capi
    No, Nv, seed, scale = 5, 11, 5, 0.1
    engI, t = engine_from_host(ab, oracle, No, Nv, seed, scale)
    engF = ab.Engine(No, Nv, field=capi.FIELD_COMPLEX)
    engF.fill_synthetic(seed, scale)
    for eng in (engF, engI):
        for abc in [(0, 3, 7), (2, 2, 9), (4, 10, 10)]:
            S = oracle.tuple_slices_z(No, Nv, t, abc)
            a, b, c = abc
            assert np.array_equal(eng.read_slice(capi.TA, a), S["TA"])
            assert np.array_equal(eng.read_slice(capi.VIJKA, b), S["HB"])
            assert np.array_equal(eng.read_slice(capi.VABCI, a, c), S["VAC"])
            assert np.array_equal(eng.read_slice(capi.VABCI, c, b), S["VCB"])
            assert np.array_equal(eng.read_slice(capi.TABIJ, b, c), S["TBC"])
            assert np.array_equal(eng.read_slice(capi.VABIJ, a, b), S["VABij"])
    engI.close()
    engF.close()


@pytest.mark.parametrize("No,Nv", [(4, 8), (8, 24), (13, 29), (16, 33), (33, 40), (40, 56), (64, 72)])
def test_complex_cubes_and_energy_vs_oracle(ab, oracle, No, Nv):
    """element-wise complex Tijk / Zijk and tuple energies against the oracle, several kernel plans"""
    from atrip_b200 import capi
    seed, scale = 2000 + No, 0.1
    t = oracle.inputs_z(No, Nv, seed=seed, scale=scale)
    eng = ab.Engine(No, Nv, field=capi.FIELD_COMPLEX)
    eng.fill_synthetic(seed, scale)
    for abc in [(0, 1, 2), (0, 0, 1), (0, 1, 1), (Nv - 3, Nv - 2, Nv - 1), (2, 2, Nv - 1), (1, Nv // 2, Nv - 2)]:
        e, _, T, Z = oracle.tuple_energy_z(No, Nv, t, abc, want_cubes=True)
        ge, gT, gZ = eng.tuple_debug(*abc)
        tmax = np.abs(T).max()
        assert np.abs(gT - T).max() <= CUBE_REL * tmax, abc
        assert np.abs(gZ - Z).max() <= CUBE_REL * max(tmax, np.abs(Z).max()), abc
        assert abs(ge - e) <= E_REL * abs(e), abc
    eng.close()


def test_complex_tuples_match_reference_vectors(ab, golden):
    """per-tuple values of the reference's own L1 functions instantiated for Complex"""
    from atrip_b200 import capi
    for rec in golden["complex_tuples"]:
        No, Nv = rec["No"], rec["Nv"]
        eng = ab.Engine(No, Nv, field=capi.FIELD_COMPLEX)
        eng.fill_synthetic(rec["seed"], rec["scale"])
        idx = [0, 1, No, No * No, No ** 3 // 2, No ** 3 - 1]
        for g in rec["tuples"]:
            e, T, Z = eng.tuple_debug(*g["abc"])
            tmax = fh(g["Tabsmax"])
            assert abs(e - fh(g["energy"])) <= E_REL * abs(e), (No, g["abc"])
            assert np.abs(T[idx] - np.array([zh(x) for x in g["Tsample"]])).max() <= CUBE_REL * tmax
            assert np.abs(Z[idx] - np.array([zh(x) for x in g["Zsample"]])).max() <= CUBE_REL * tmax
            assert abs(T.sum() - zh(g["Tsum"])) <= 1e-11 * tmax * No ** 1.5
        eng.close()


def test_complex_runs_match_reference_vectors(ab, oracle, golden):
    """whole Atrip::run<Complex> energies of the reference incl. (cT); ingest and device fill agree"""
    from atrip_b200 import capi
    for r in golden["complex_runs"]:
        eng, _ = engine_from_host(ab, oracle, r["No"], r["Nv"], r["seed"], r["scale"], with_J=r["with_J"])
        eng.build_tuples(capi.GROUP_AND_SORT)
        e, ct = eng.run()
        eng.close()
        assert close_energy(-e, fh(r["energy"])), (r, -e)
        ref_ct = fh(r["ct_energy"])
        assert abs(-ct - ref_ct) <= E_ABS and abs(-ct - ref_ct) <= 1e-11 * max(abs(ref_ct), abs(e)), (r, -ct)
        eng = ab.Engine(r["No"], r["Nv"], with_J=r["with_J"], field=capi.FIELD_COMPLEX, batch_tuples=23)
        eng.fill_synthetic(r["seed"], r["scale"])
        eng.build_tuples(capi.GROUP_AND_SORT)
        e2, ct2 = eng.run()
        eng.close()
        assert abs(e2 - e) <= 1e-13 * abs(e) and abs(ct2 - ct) <= 1e-12 * max(abs(e), abs(ct))


def test_complex_with_zero_imaginary_parts_equals_real_engine(ab, oracle):
    """field = 1 on real data must reproduce the field = 0 engine"""
    from atrip_b200 import capi
    No, Nv, seed, scale = 9, 21, 8, 0.05
    tr = oracle.inputs(No, Nv, seed=seed, scale=scale)
    er = ab.Engine(No, Nv)
    er.load_all(tr[EPS_I], tr[EPS_A], tr[TAI], tr[TABIJ], tr[VABIJ], tr[VIJKA], tr[VABCI])
    er.build_tuples(capi.GROUP_AND_SORT)
    e_real, _ = er.run()
    er.close()
    tz = {k: v.astype(np.complex128) for k, v in tr.items()}
    ez = ab.Engine(No, Nv, field=capi.FIELD_COMPLEX)
    ez.load_all(tz[EPS_I], tz[EPS_A], tz[TAI], tz[TABIJ], tz[VABIJ], tz[VIJKA], tz[VABCI])
    ez.build_tuples(capi.GROUP_AND_SORT)
    e_cplx, _ = ez.run()
    assert ez.flops_per_tuple == 4 * 12 * No ** 3 * (No + Nv)  # Atrip.cxx:578-580
    ez.close()
    assert abs(e_real - e_cplx) <= 1e-12 * abs(e_real)


def test_atrip_run_complex_host_api(golden):
    """atrip::Atrip::run<Complex> on CTF::Tensor<Complex> inputs (C++ host API)"""
    host = os.path.join(ROOT, "atrip_b200", "host")
    exe = os.path.join(host, "synth_driver")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", host, "libatrip.so", "synth_driver"])
    for r in golden["complex_runs"][:4]:
        cmd = [exe, str(r["No"]), str(r["Nv"]), str(r["seed"]), repr(r["scale"]), "0", "group",
               "cT" if r["with_J"] else "T", "complex"]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        out = p.stdout + p.stderr
        assert p.returncode == 0, out
        m = re.search(r"RESULT energy (\S+) \S+ ct_energy (\S+)", out)
        assert m, out
        e, ct = fh(m.group(1)), fh(m.group(2))
        assert close_energy(e, fh(r["energy"])), (r, e)
        ref_ct = fh(r["ct_energy"])
        assert abs(ct - ref_ct) <= E_ABS and abs(ct - ref_ct) <= 1e-11 * max(abs(ref_ct), abs(e)), (r, ct)
