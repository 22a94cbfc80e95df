"""GPU: checkpoint / resume through the C++ API (reference Atrip.cxx:586-621 read side; the
reference's write side is disabled at HEAD, `&& false` at :716, ours writes the same file format,
tests/test_checkpoint.py).  A run interrupted by max_iterations leaves a checkpoint; a second run
resumes from it and must end at the energy of an uninterrupted run."""
import os
import re
import subprocess
import tempfile

import pytest

from conftest import ROOT, fh

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "atrip_b200", "host")


def drive(args, env=None):
    exe = os.path.join(HOST, "synth_driver")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", HOST, "libatrip.so", "synth_driver"])
    p = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    out = p.stdout + p.stderr
    assert p.returncode == 0, out
    m = re.search(r"RESULT energy (\S+) \S+ ct_energy (\S+)", out)
    assert m, out
    return fh(m.group(1)), out


@pytest.mark.parametrize("field", ["real", "complex"])
def test_resume_from_checkpoint_reaches_the_uninterrupted_energy(field):
    base = ["6", "15", "3", "0.05"]
    tail = ["group", "T"] + (["complex"] if field == "complex" else [])
    full, _ = drive(base + ["0"] + tail)
    with tempfile.TemporaryDirectory() as tmp:
        ck = os.path.join(tmp, "atrip-checkpoint.yaml")
        env = {"SYNTH_CHECKPOINT": ck, "SYNTH_CHECKPOINT_EVERY": "10"}
        part, _ = drive(base + ["40"] + tail, env)            # leaves after iteration index 40 (41 tuples)
        assert os.path.exists(ck) and abs(part - full) > 1e-6 * abs(full)
        txt = open(ck).read()
        assert "Iteration: 40" in txt and "No: 6" in txt and "Nv: 15" in txt, txt
        resumed, out = drive(base + ["0"] + tail, env)
        assert "Reading checkpoint" in out and "iteration from checkpoint 40" in out
        assert abs(resumed - full) <= 1e-12 * abs(full), (resumed, full)
        # a run that finished the list removes its checkpoint: the next run starts from scratch
        assert not os.path.exists(ck)
        again, out = drive(base + ["0"] + tail, env)
        assert "Reading checkpoint" not in out and abs(again - full) <= 1e-12 * abs(full)


def test_checkpoint_of_another_calculation_is_refused():
    with tempfile.TemporaryDirectory() as tmp:
        ck = os.path.join(tmp, "atrip-checkpoint.yaml")
        open(ck, "w").write("No: 6\nNv: 15\nNranks: 1\nNnodes: 8\nEnergy: -0.5\nIteration: 40\nRankRoundRobin: false\n")
        exe = os.path.join(HOST, "synth_driver")
        p = subprocess.run([exe, "6", "15", "3", "0.05", "0", "group", "T"], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, SYNTH_CHECKPOINT=ck))
        assert p.returncode != 0 and "does not belong to this calculation" in p.stdout + p.stderr
