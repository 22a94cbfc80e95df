"""GPU probe: smallest problems against the oracle (No = 1..3, Nv = 2..5, with and without J, both fields)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import atrip_b200
from atrip_b200 import capi
from oracle.oracle import Oracle
o = Oracle()
bad = 0
for No, Nv in ((1, 2), (1, 3), (2, 2), (3, 5), (2, 9), (9, 2), (17, 3)):
    for with_J in (False, True):
        t = o.inputs(No, Nv, seed=5, scale=0.3, with_J=with_J)
        want, want_ct = o.run(No, Nv, t)
        eng = atrip_b200.Engine(No, Nv, with_J=with_J)
        eng.fill_synthetic(5, 0.3)
        eng.build_tuples(capi.GROUP_AND_SORT)
        e, ct = eng.run()
        eng.close()
        ok = abs(-e - want) <= 1e-12 * max(abs(want), 1e-300) + 1e-15 and abs(-ct - want_ct) <= 1e-11 * max(abs(want), abs(want_ct)) + 1e-15
        bad += not ok
        print(f"No {No} Nv {Nv} J {int(with_J)}: gpu {-e!r} {-ct!r} oracle {want!r} {want_ct!r} {'ok' if ok else 'MISMATCH'}", flush=True)
    tz = o.inputs_z(No, Nv, seed=5, scale=0.3)
    wz, _ = o.run_z(No, Nv, tz)
    eng = atrip_b200.Engine(No, Nv, field=capi.FIELD_COMPLEX)
    eng.fill_synthetic(5, 0.3)
    eng.build_tuples(capi.GROUP_AND_SORT)
    e, _ = eng.run()
    eng.close()
    ok = abs(-e - wz) <= 1e-12 * max(abs(wz), 1e-300) + 1e-15
    bad += not ok
    print(f"No {No} Nv {Nv} complex: gpu {-e!r} oracle {wz!r} {'ok' if ok else 'MISMATCH'}", flush=True)
print("EDGE CASES:", "all ok" if not bad else f"{bad} mismatches")
