"""GPU: bench.py prints ONE JSON line carrying every key of the driver's contract (a short c2 run,
value leg only plus a shortened CPU-baseline leg), and the numbers are self-consistent; both e2e
ingest paths on the CPU-runnable configuration."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_bench_line_has_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "c2", "--steps", "2", "--warmup", "3",
                        "--no-e2e", "--cpu-seconds", "2", "--tuples-per-step", "6000"], capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline",
              "parity"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "vendor_peak", "step_frac"):
        assert k in r, k
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0.5 < r["frac"] <= 1.0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 30.0 < r["peak"] < 45.0
    assert 0.0 < d["value"] <= r["achieved"] * 1.001          # whole step cannot beat its dominant kernel
    assert d["gpu_launches"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and 0.0 < c["value"] < d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    # in-bench parity: the golden whole run and three tuples of the bench configuration
    par = d["parity"]
    assert par["ok"] is True and par["n_gpus"] == 1 and par["max_rel"] <= 1e-12
    assert par["golden_run"]["abs"] <= 1e-10 and len(par["tuples"]) == 3


@pytest.mark.parametrize("mode", ["tensors", "slices"])
def test_bench_e2e_paths_reproduce_the_resident_energy(mode):
    """e2e leg through atrip_b200_load_* (CTF-layout host tensors) and through atrip_b200_upload_slices
    (per-owner slices of the step's inputs): the step energy equals the device-resident run's"""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "c1", "--steps", "2", "--warmup", "3",
                        "--no-cpu", "--e2e-mode", mode], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][0])
    e = d["e2e"]
    assert e["energy_matches_resident_run"] is True and e["h2d_bytes_per_step"] > 0 and e["value"] > 0
    assert d["parity"]["ok"] is True
