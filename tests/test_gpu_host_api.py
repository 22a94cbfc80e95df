"""GPU: the C++ drop-in boundary.  atrip::Atrip::run<double> (include/atrip/Atrip.hpp,
atrip_b200/host/Atrip.cxx) through the same calls the reference's bench makes, and the
reference's own bench/main.cxx compiled unchanged against that API (built where
/root/reference exists; the binary travels to the GPU box)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT, fh

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "atrip_b200", "host")


def run(cmd, **kw):
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, **kw)
    return p.returncode, p.stdout + p.stderr


def result(out):
    m = re.search(r"RESULT energy (\S+) \S+ ct_energy (\S+)", out)
    assert m, out
    return fh(m.group(1)), fh(m.group(2))


@pytest.fixture(scope="module")
def driver():
    exe = os.path.join(HOST, "synth_driver")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", HOST, "libatrip.so", "synth_driver"])
    return exe


def test_atrip_run_matches_reference_vectors(driver, golden):
    """Atrip::Input -> Atrip::run -> Output on host CTF tensors; energies of the reference"""
    for r in golden["runs"]:
        cmd = [driver, str(r["No"]), str(r["Nv"]), str(r["seed"]), repr(r["scale"]), "0", "group"]
        if r["with_J"]:
            cmd.append("cT")
        rc, out = run(cmd)
        assert rc == 0, out
        e, ct = result(out)
        ref, ref_ct = fh(r["energy"]), fh(r["ct_energy"])
        assert abs(e - ref) <= 1e-10 and abs(e - ref) <= 1e-12 * abs(ref), (r, e)
        assert abs(ct - ref_ct) <= 1e-10 and abs(ct - ref_ct) <= 1e-11 * max(abs(ref), abs(ref_ct)), (r, ct)
        assert "Atrip: Energy:" in out and "atrip:flops(doubles)" in out


def test_atrip_run_naive_equals_group(driver):
    """both distributions enumerate the same tuples at np=1 (Tuples.cxx:89-141, 310-407)"""
    e1 = result(run([driver, "6", "15", "3", "0.05", "0", "group"])[1])
    e2 = result(run([driver, "6", "15", "3", "0.05", "0", "naive"])[1])
    assert abs(e1[0] - e2[0]) <= 1e-13 * abs(e1[0])


def test_max_iterations_follows_reference(driver, oracle):
    """the reference leaves its loop after iteration index max_iterations, i.e. processes
    max_iterations + 1 tuples (Atrip.cxx:1052-1056)"""
    No, Nv, seed, scale, mx = 6, 15, 3, 0.05, 40
    e, _ = result(run([driver, str(No), str(Nv), str(seed), repr(scale), str(mx), "group"])[1])
    t = oracle.inputs(No, Nv, seed=seed, scale=scale)
    want, _ = oracle.run(No, Nv, t, tuples=oracle.all_tuples(Nv)[:mx + 1])
    assert abs(e - want) <= 1e-12 * abs(want)


def test_reference_bench_driver_runs_against_this_api():
    exe = os.path.join(HOST, "atrip_bench")
    if not os.path.exists(exe):
        pytest.skip("atrip_bench is built only where /root/reference is present")
    rc, out = run([exe, "--no", "6", "--nv", "20", "--dist", "group", "--nocheckpoint", "-%", "50"], cwd="/tmp")
    assert rc == 0, out
    assert "Atrip throwed" not in out, out
    assert re.search(r"^Energy: ", out, re.M) and re.search(r"^Energy \(cT\): ", out, re.M), out
    assert "Progress(%)" in out  # the driver's register_iteration_descriptor callback fired


def test_ijkabc_mode_matches_reference_vectors(driver):
    """Input::ijkabc (Atrip.cxx:183-187: Tai negated; :1108-1111: no final sign flip), F = double and
    F = Complex, against whole runs of the reference with the same switch (tests/golden/ijkabc_vectors.json)"""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "ijkabc_vectors.json")) as f:
        g = json.load(f)
    for field, runs in (("real", g["runs"]), ("complex", g["complex_runs"])):
        for r in runs:
            cmd = [driver, str(r["No"]), str(r["Nv"]), str(r["seed"]), repr(r["scale"]), "0", "group",
                   "cT" if r["with_J"] else "T", field, "ijkabc"]
            rc, out = run(cmd)
            assert rc == 0, out
            e, ct = result(out)
            ref, ref_ct = fh(r["energy"]), fh(r["ct_energy"])
            assert ref > 0  # the unflipped sign
            assert abs(e - ref) <= 1e-10 and abs(e - ref) <= 1e-12 * abs(ref), (field, r, e)
            assert abs(ct - ref_ct) <= 1e-10 and abs(ct - ref_ct) <= 1e-11 * max(abs(ref), abs(ref_ct)), (field, r, ct)


@pytest.mark.parametrize("world", [2, 4])
def test_atrip_run_on_several_ranks(driver, golden, world):
    """one rank per GPU through the C++ API (reference Atrip.cxx:54-63, 84-91, 119): Atrip::init sees np
    ranks (MPI stand-in in multi-process mode), every rank slices and uploads only the sources it owns
    (SliceUnion.cxx:305-332), slices travel between the GPUs, the energy is summed by ncclAllReduce"""
    import sys
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from launch_ranks import launch
    for r in [golden["runs"][i] for i in (1, 3, 5)]:
        cmd = [driver, str(r["No"]), str(r["Nv"]), str(r["seed"]), repr(r["scale"]), "0", "group",
               "cT" if r["with_J"] else "T"]
        rc, out, outs = launch(world, cmd)
        assert rc == 0, outs
        e, ct = result(out)
        ref, ref_ct = fh(r["energy"]), fh(r["ct_energy"])
        assert abs(e - ref) <= 1e-10 and abs(e - ref) <= 1e-12 * abs(ref), (r, e)
        assert abs(ct - ref_ct) <= 1e-10 and abs(ct - ref_ct) <= 1e-11 * max(abs(ref), abs(ref_ct)), (r, ct)
        assert f"np: {world}" in out
    # complex field on two ranks
    r = golden["complex_runs"][1]
    rc, out, outs = launch(2, [driver, str(r["No"]), str(r["Nv"]), str(r["seed"]), repr(r["scale"]), "0", "group",
                               "cT" if r["with_J"] else "T", "complex"])
    assert rc == 0, outs
    e, ct = result(out)
    assert abs(e - fh(r["energy"])) <= 1e-12 * abs(e)


def test_reference_bench_driver_on_two_ranks():
    """the reference's own bench/main.cxx, unchanged, one process per GPU"""
    import sys
    import torch
    exe = os.path.join(HOST, "atrip_bench")
    if not os.path.exists(exe):
        pytest.skip("atrip_bench is built only where /root/reference is present")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from launch_ranks import launch
    rc, out, outs = launch(2, [exe, "--no", "6", "--nv", "20", "--dist", "group", "--nocheckpoint", "-%", "50"], cwd="/tmp")
    assert rc == 0, outs
    assert "Atrip throwed" not in out and re.search(r"^Energy: ", out, re.M), out


def _write_tensor_files(oracle, tmp_path, r):
    """the reference's on-disk tensor format (bench/main.cxx:54-75: CTF read_dense_from_file = raw native-endian
    FP64 in global column-major order): the golden run's inputs, one file per tensor"""
    from oracle.oracle import EPS_A, EPS_I, JABCI, JIJKA, TABIJ, TAI, VABCI, VABIJ, VIJKA
    t = oracle.inputs(r["No"], r["Nv"], seed=r["seed"], scale=r["scale"], with_J=r["with_J"])
    flags = {"--ei": EPS_I, "--ea": EPS_A, "--Tph": TAI, "--Tpphh": TABIJ, "--Vpphh": VABIJ, "--Vhhhp": VIJKA,
             "--Vppph": VABCI}
    if r["with_J"]:
        flags.update({"--Jhhhp": JIJKA, "--Jppph": JABCI})
    args = []
    for flag, tid in flags.items():
        path = os.path.join(str(tmp_path), flag.strip("-") + ".bin")
        t[tid].astype("<f8").tofile(path)
        args += [flag, path]
    return args


@pytest.mark.parametrize("idx", [5, 4])
def test_reference_bench_driver_with_tensor_files(oracle, golden, tmp_path, idx):
    """file inputs (survey row f3): the reference's unchanged bench/main.cxx reads every tensor with
    read_dense_from_file (--ei --ea --Tph --Tpphh --Vpphh --Vhhhp --Vppph, --cT --Jhhhp --Jppph) and prints the
    reference's own energies for those inputs (golden whole runs: No=10 Nv=40, and No=7 Nv=13 with J)"""
    exe = os.path.join(HOST, "atrip_bench")
    if not os.path.exists(exe):
        pytest.skip("atrip_bench is built only where /root/reference is present")
    r = golden["runs"][idx]
    cmd = [exe, "--no", str(r["No"]), "--nv", str(r["Nv"]), "--dist", "group", "--nocheckpoint", "-%", "50"]
    cmd += _write_tensor_files(oracle, tmp_path, r) + (["--cT"] if r["with_J"] else [])
    rc, out = run(cmd, cwd=str(tmp_path))
    assert rc == 0 and "Atrip throwed" not in out and "Random initialization" not in out, out
    e = float(re.search(r"^Energy: (\S+)", out, re.M).group(1))
    ct = float(re.search(r"^Energy \(cT\): (\S+)", out, re.M).group(1))
    ref, ref_ct = fh(r["energy"]), fh(r["ct_energy"])
    # the driver prints 15 significant digits (Atrip.cxx:1113-1116)
    assert abs(e - ref) <= 2e-14 * abs(ref) + 1e-15, (e, ref)
    assert abs(ct - ref_ct) <= 2e-14 * max(abs(ref), abs(ref_ct)) + 1e-15, (ct, ref_ct)
