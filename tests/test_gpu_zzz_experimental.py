"""GPU: the two reduction kernels against each other.  The default (T)-pass kernel of the real field is the
bulk-copy kernel (reduction_async.cuh); ATRIP_B200_REDUCE=sync selects the register-staged kernel
(reduction.cuh, still used for the (cT) pass), ATRIP_B200_REDUCE=async-rev the reversed tuple walk.  Every
engine runs in a child process with a time limit, so that a kernel that hangs is killed with its CUDA
context instead of stalling the suite."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

CHILD = r"""
import json, sys
sys.path.insert(0, %r)
import atrip_b200
from atrip_b200 import capi
No, Nv = int(sys.argv[1]), int(sys.argv[2])
eng = atrip_b200.Engine(No, Nv)
eng.fill_synthetic(7, 0.05)
n = eng.build_tuples(capi.GROUP_AND_SORT)
e = [eng.run(0, min(n, 3000))[0], eng.run(0, 1)[0], eng.run(5, 40)[0]]
eng.close()
print("RESULT " + json.dumps([x.hex() for x in e]))
""" % ROOT


def totals(No, Nv, reduce_mode):
    env = {k: v for k, v in os.environ.items() if k != "ATRIP_B200_REDUCE"}
    if reduce_mode:
        env["ATRIP_B200_REDUCE"] = reduce_mode
    p = subprocess.run([sys.executable, "-c", CHILD, str(No), str(Nv)], env=env, capture_output=True, text=True,
                       timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return np.array([float.fromhex(x) for x in json.loads(line[7:])])


@pytest.mark.parametrize("mode", ["sync", "async-rev"])
@pytest.mark.parametrize("No,Nv", [(8, 16), (16, 24), (33, 40), (40, 56)])
def test_reduction_kernels_agree(No, Nv, mode):
    want = totals(No, Nv, None)
    try:
        got = totals(No, Nv, mode)
    except subprocess.TimeoutExpired:
        pytest.fail("the reduction did not finish within 120 s (killed)")
    assert np.all(np.abs(got - want) <= 1e-13 * np.abs(want)), (got, want)
