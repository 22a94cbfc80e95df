"""GPU: the CUDA path through the C-ABI against (1) vectors produced by the reference itself
(tests/golden/), (2) the oracle on the same seeded inputs, (3) size-independent properties at
the bench sizes.  Tolerances follow BASELINE.json north_star: energy |dE| <= 1e-10 Eh and
rel <= 1e-12 on inputs scaled to |E| = O(0.01..1); cubes 1e-13 relative to max|Tijk|."""
import numpy as np
import pytest

from conftest import fh
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


def engine_from_host(ab, o, No, Nv, seed, scale, with_J=False, **kw):
    t = o.inputs(No, Nv, seed=seed, scale=scale, with_J=with_J)
    eng = ab.Engine(No, Nv, with_J=with_J, **kw)
    eng.load_all(t[EPS_I], t[EPS_A], t[TAI], t[TABIJ], t[VABIJ], t[VIJKA], t[VABCI], t.get(JIJKA), t.get(JABCI))
    return eng, t


def close_energy(e, ref):
    return abs(e - ref) <= E_ABS and abs(e - ref) <= E_REL * abs(ref)


def test_runs_match_reference_vectors(ab, oracle, golden):
    """whole Atrip::run energies of the reference (np=1, GROUP_AND_SORT), ingest path"""
    from atrip_b200 import capi
    for r in golden["runs"]:
        eng, _ = engine_from_host(ab, oracle, r["No"], r["Nv"], r["seed"], r["scale"], with_J=r["with_J"])
        eng.build_tuples(capi.GROUP_AND_SORT)
        e, ct = eng.run()
        assert close_energy(-e, fh(r["energy"])), (r, -e)
        ref_ct = fh(r["ct_energy"])
        assert abs(-ct - ref_ct) <= E_ABS and abs(-ct - ref_ct) <= 1e-11 * max(abs(ref_ct), abs(e)), (r, -ct)
        eng.close()


def test_runs_device_fill_equals_ingest(ab, oracle, golden):
    """synthetic stores generated on the device give the same energies bit for bit"""
    from atrip_b200 import capi
    for r in golden["runs"][:4]:
        eng, _ = engine_from_host(ab, oracle, r["No"], r["Nv"], r["seed"], r["scale"], with_J=r["with_J"])
        eng.build_tuples(capi.GROUP_AND_SORT)
        e1 = eng.run()
        eng.close()
        eng = ab.Engine(r["No"], r["Nv"], with_J=r["with_J"])
        eng.fill_synthetic(r["seed"], r["scale"])
        eng.build_tuples(capi.GROUP_AND_SORT)
        assert eng.run() == e1
        eng.close()


def test_tuples_match_reference_vectors(ab, oracle, golden):
    """per-tuple Tijk / Zijk / energy of the reference's doubles/singles/energy functions"""
    for rec in golden["tuples"]:
        No, Nv = rec["No"], rec["Nv"]
        eng = ab.Engine(No, Nv)
        eng.fill_synthetic(rec["seed"], rec["scale"])
        idx = [0, 1, No, No * No, No ** 3 // 2, No ** 3 - 1]
        for g in rec["tuples"]:
            e, T, Z = eng.tuple_debug(*g["abc"])
            tmax = fh(g["Tabsmax"])
            assert abs(e - fh(g["energy"])) <= E_REL * abs(e), (rec["No"], g["abc"])
            assert np.abs(T[idx] - [fh(x) for x in g["Tsample"]]).max() <= CUBE_REL * tmax
            assert np.abs(Z[idx] - [fh(x) for x in g["Zsample"]]).max() <= CUBE_REL * tmax
            assert abs(T.sum() - fh(g["Tsum"])) <= 1e-11 * tmax * No ** 1.5
        eng.close()


@pytest.mark.parametrize("No,Nv", [(4, 8), (8, 24), (10, 40), (13, 29), (16, 33), (33, 40), (40, 56), (64, 72),
                                   (100, 104), (7, 120)])
def test_cubes_and_energy_vs_oracle(ab, oracle, No, Nv):
    """odd and even sizes, all kernel variants: element-wise cubes against the oracle"""
    seed, scale = 1000 + No, 0.1
    t = oracle.inputs(No, Nv, seed=seed, scale=scale)
    eng = ab.Engine(No, Nv)
    eng.fill_synthetic(seed, scale)
    tuples = [(0, 1, 2), (0, 0, 1), (0, 1, 1), (Nv - 3, Nv - 2, Nv - 1), (0, Nv - 1, Nv - 1), (2, 2, Nv - 1),
              (1, Nv // 2, Nv - 2)]
    for abc in tuples:
        e, _, T, Z = oracle.tuple_energy(No, Nv, t, abc, want_cubes=True)
        ge, gT, gZ = eng.tuple_debug(*abc)
        tmax = np.abs(T).max()
        assert np.abs(gT - T).max() <= CUBE_REL * tmax, abc
        assert np.abs(gZ - Z).max() <= CUBE_REL * max(tmax, np.abs(Z).max()), abc
        assert abs(ge - e) <= E_REL * abs(e), abc
    eng.close()


def test_slices_bit_exact_fill_and_ingest(ab, oracle):
    """stores hold exactly the reference's slices (Unions.hpp layouts), from either source"""
    from atrip_b200 import capi
    No, Nv, seed, scale = 6, 19, 5, 0.1
    engI, t = engine_from_host(ab, oracle, No, Nv, seed, scale)
    engF = ab.Engine(No, Nv)
    engF.fill_synthetic(seed, scale)
    for eng in (engI, engF):
        for x in range(0, Nv, 3):
            assert np.array_equal(eng.read_slice(capi.TA, x), oracle.slice_TA(No, Nv, t[TABIJ], x))
            assert np.array_equal(eng.read_slice(capi.VIJKA, x), oracle.slice_HHHA(No, Nv, t[VIJKA], x))
            for y in range(0, Nv, 4):
                assert np.array_equal(eng.read_slice(capi.VABCI, x, y), oracle.slice_ABPH(No, Nv, t[VABCI], x, y))
                assert np.array_equal(eng.read_slice(capi.TABIJ, x, y), oracle.slice_ABHH(No, Nv, t[TABIJ], x, y))
                if x <= y:
                    assert np.array_equal(eng.read_slice(capi.VABIJ, x, y), oracle.slice_ABHH(No, Nv, t[VABIJ], x, y))
    engI.close()
    engF.close()


def test_explicit_tuple_lists_fakes_and_batches(ab, oracle):
    """set_tuples with fake tuples interleaved, several batch sizes, sub-ranges: same sums"""
    No, Nv, seed, scale = 9, 21, 8, 0.05
    t = oracle.inputs(No, Nv, seed=seed, scale=scale)
    allt = oracle.all_tuples(Nv)
    sub = allt[::7]
    with_fakes = np.concatenate([sub[:50], np.zeros((5, 3), np.uint64), sub[50:], np.zeros((3, 3), np.uint64)])
    want, _ = oracle.run(No, Nv, t, tuples=sub)
    got = []
    for batch in (0, 1, 37, 4096):
        eng = ab.Engine(No, Nv, batch_tuples=batch)
        eng.fill_synthetic(seed, scale)
        eng.set_tuples(with_fakes)
        e, _ = eng.run()
        assert close_energy(-e, want)
        assert eng.last_timing()["tuples"] == len(sub)  # fakes are skipped, not counted
        n = eng.num_tuples()
        e1, _ = eng.run(0, n // 3)
        e2, _ = eng.run(n // 3, n - n // 3)
        assert abs((e1 + e2) - e) <= 1e-13 * abs(e)
        got.append(e)
        eng.close()
    assert max(got) - min(got) <= 1e-13 * abs(got[0])
    # empty range and empty list
    eng = ab.Engine(No, Nv)
    eng.fill_synthetic(seed, scale)
    eng.set_tuples(np.zeros((0, 3), np.uint64))
    assert eng.run() == (0.0, 0.0)
    eng.close()


def test_group_and_sort_shards_sum_to_total(ab, oracle):
    """multi-GPU sharding logic on one device: the per-rank lists of a 4-GPU job, run one after
    the other, add up to the single-rank energy (the final allreduce is a plain sum)"""
    from atrip_b200 import capi
    No, Nv, seed, scale = 8, 22, 77, 0.05
    t = oracle.inputs(No, Nv, seed=seed, scale=scale)
    want, _ = oracle.run(No, Nv, t)
    total = 0.0
    for r in range(4):
        eng = ab.Engine(No, Nv, rank=r, nranks=4)
        eng.fill_synthetic(seed, scale)
        eng.build_tuples(capi.GROUP_AND_SORT)
        e, _ = eng.run()
        total += e
        eng.close()
    assert close_energy(-total, want)


def test_determinism_and_bench_size_properties(ab):
    """bench-size (No=40, Nv=400) slice of tuples: run-to-run bit reproducibility, additivity over
    sub-ranges and invariance to the batch size -- properties that need no CPU reference"""
    from atrip_b200 import capi
    No, Nv = 40, 400
    eng = ab.Engine(No, Nv)
    eng.fill_synthetic(12345, 0.01)
    n = eng.build_tuples(capi.GROUP_AND_SORT)
    assert n == Nv * (Nv + 1) * (Nv + 2) // 6 - Nv
    first, count = n // 2, 3000
    e1 = eng.run(first, count)
    e2 = eng.run(first, count)
    assert e1 == e2 and np.isfinite(e1[0]) and e1[0] != 0.0
    ea = eng.run(first, 1000)[0] + eng.run(first + 1000, 2000)[0]
    assert abs(ea - e1[0]) <= 1e-12 * abs(e1[0])
    tup = eng.get_tuples()[first:first + count]
    eng.close()
    eng = ab.Engine(No, Nv, batch_tuples=250)
    eng.fill_synthetic(12345, 0.01)
    eng.set_tuples(tup)
    assert abs(eng.run()[0] - e1[0]) <= 1e-12 * abs(e1[0])
    eng.close()


@pytest.mark.parametrize("No,Nv,field", [(16, 48, 0), (16, 24, 1), (8, 24, 0), (24, 40, 1), (64, 64, 0)])
def test_repeated_runs_are_bit_identical_without_k_padding(ab, No, Nv, field):
    """regression for the ring-release race (profiles/r01_ring_release_race.txt): shapes whose
    contraction length has NO zero padding ((No+Nv) % 16 == 0; real and complex field) used a
    straight-line loop body in which ptxas handed ring stages back before their last fragment loads
    had landed -- run-to-run different cubes in ~1 of 10 large launches.  Same batch, many times:
    energies and the integer checksum of the class cubes must be identical every time."""
    from atrip_b200 import capi
    eng = ab.Engine(No, Nv, field=field)
    assert eng.kp == (2 if field else 1) * (No + Nv)
    eng.fill_synthetic(5, 0.01)
    n = eng.build_tuples(capi.GROUP_AND_SORT)
    cnt = min(n, eng.batch_tuples)
    seen = set()
    for _ in range(25 if No < 64 else 8):
        e, _ = eng.run(0, cnt)
        seen.add((e, eng.cubes_checksum()))
    eng.close()
    assert len(seen) == 1, seen


@pytest.mark.parametrize("No,Nv,tuples", [
    (40, 400, [(3, 57, 399), (11, 11, 200)]),      # c2, the N=1 bench workload (padded K, 5x5 plan)
    (64, 192, [(5, 64, 191), (100, 100, 101)]),    # the c3 plan (4x8 fragments) with UNPADDED K = 256
])
def test_bench_size_tuples_vs_oracle(ab, oracle, No, Nv, tuples):
    """a few tuples at the bench size against the oracle's per-tuple path fed slice by slice
    (the full tensors would be 25 GB on the host: the oracle is given synthetic slices)"""
    seed, scale = 12345, 0.1
    eng = ab.Engine(No, Nv)
    eng.fill_synthetic(seed, scale)
    epsi, epsa = oracle.fill(seed, EPS_I, scale, No), oracle.fill(seed, EPS_A, scale, Nv)
    tai = oracle.fill(seed, TAI, scale, No * Nv)

    for abc in tuples:
        a, b, c = abc
        S = oracle.synth_tuple_slices(No, Nv, abc, seed=seed, scale=scale)
        T = oracle.doubles(No, Nv, S)
        Z = oracle.singles(No, Nv, abc, tai, S, T)
        eps = float(epsa[a] + epsa[b] + epsa[c])
        same = (a == b) != (b == c)
        e = (oracle.energy_same if same else oracle.energy_distinct)(eps, No, epsi, T, Z)
        ge, gT, gZ = eng.tuple_debug(*abc)
        assert np.abs(gT - T).max() <= CUBE_REL * np.abs(T).max()
        assert np.abs(gZ - Z).max() <= CUBE_REL * np.abs(Z).max()
        assert abs(ge - e) <= E_REL * abs(e)
    eng.close()


def test_two_live_engines_of_different_size(ab, golden):
    """regression: the dynamic shared-memory cap is an attribute of the kernel FUNCTION, not of a context -- an
    engine with a small No created while a larger one is alive used to lower the cap under it and the larger
    engine's next reduction launch failed with "invalid argument" (bench.py keeps its engine while the golden
    case runs).  Every combination of (T)/(cT) and both fields, interleaved."""
    from atrip_b200 import capi
    runs = {(r["No"], r["Nv"]): r for r in golden["runs"] if not r["with_J"]}
    big_r, small_r = runs[(16, 24)], runs[(4, 8)]
    big = ab.Engine(big_r["No"], big_r["Nv"], with_J=True)
    big.fill_synthetic(big_r["seed"], big_r["scale"])
    big.build_tuples(capi.GROUP_AND_SORT)
    e0 = big.run()
    small = ab.Engine(small_r["No"], small_r["Nv"], with_J=True)     # lowers nothing any more
    small.fill_synthetic(small_r["seed"], small_r["scale"])
    small.build_tuples(capi.GROUP_AND_SORT)
    es = small.run()
    small_z = ab.Engine(small_r["No"], small_r["Nv"], field=capi.FIELD_COMPLEX)
    small_z.fill_synthetic(small_r["seed"], small_r["scale"])
    small_z.build_tuples(capi.GROUP_AND_SORT)
    small_z.run()
    assert big.run() == e0                                           # the large engine still launches
    assert close_energy(-e0[0], fh(big_r["energy"])) and close_energy(-es[0], fh(small_r["energy"]))
    for e in (small_z, small, big):
        e.close()


@pytest.mark.parametrize("No,Nv", [(1, 2), (1, 3), (2, 2), (3, 5), (2, 9), (9, 2), (17, 3)])
def test_smallest_and_lopsided_problems(ab, oracle, No, Nv):
    """edge cases: a single occupied orbital (every energy term vanishes at i=j=k), two virtual orbitals (only
    "same" tuples), more occupied than virtual orbitals; (T) and (cT), real and complex field, against the oracle
    (green on a B200: profiles/r02o_edge_cases_gpu.txt)"""
    from atrip_b200 import capi
    t = oracle.inputs(No, Nv, seed=5, scale=0.3, with_J=True)
    want, want_ct = oracle.run(No, Nv, t)
    eng = ab.Engine(No, Nv, with_J=True)
    eng.fill_synthetic(5, 0.3)
    eng.build_tuples(capi.GROUP_AND_SORT)
    e, ct = eng.run()
    eng.close()
    assert abs(-e - want) <= E_REL * abs(want) + 1e-15, (-e, want)
    assert abs(-ct - want_ct) <= 1e-11 * max(abs(want), abs(want_ct)) + 1e-15, (-ct, want_ct)
    tz = oracle.inputs_z(No, Nv, seed=5, scale=0.3)
    wz, _ = oracle.run_z(No, Nv, tz)
    eng = ab.Engine(No, Nv, field=capi.FIELD_COMPLEX)
    eng.fill_synthetic(5, 0.3)
    eng.build_tuples(capi.GROUP_AND_SORT)
    ez, _ = eng.run()
    eng.close()
    assert abs(-ez - wz) <= E_REL * abs(wz) + 1e-15, (-ez, wz)
