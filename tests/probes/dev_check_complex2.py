"""developer probe (GPU box): localise a whole-run mismatch of the complex engine.
usage: dev_check_complex2.py No Nv seed scale"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import atrip_b200
from atrip_b200 import capi
from oracle.oracle import EPS_A, EPS_I, TABIJ, TAI, VABCI, VABIJ, VIJKA, Oracle

No, Nv, seed, scale = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
o = Oracle()
t = o.inputs_z(No, Nv, seed=seed, scale=scale)
t0 = time.time()


def mk(src, **kw):
    eng = atrip_b200.Engine(No, Nv, field=1, **kw)
    if src == "fill":
        eng.fill_synthetic(seed, scale)
    else:
        eng.load_all(t[EPS_I], t[EPS_A], t[TAI], t[TABIJ], t[VABIJ], t[VIJKA], t[VABCI])
    eng.build_tuples(capi.GROUP_AND_SORT)
    return eng


engF, engI = mk("fill"), mk("ingest")
tl = engF.get_tuples()
n = len(tl)
want = np.array([o.tuple_energy_z(No, Nv, t, tuple(int(x) for x in abc))[0] for abc in tl])
print("oracle per-tuple energies: n", n, "sum", -want.sum(), "t", time.time() - t0, flush=True)
for name, eng in (("fill", engF), ("ingest", engI)):
    e, _ = eng.run()
    print("full run", name, -e, "diff", e - want.sum(), flush=True)

# slices of the ingest engine vs the fill engine (bitwise)
bad = 0
for x in range(Nv):
    for kind in (capi.TA, capi.VIJKA):
        bad += not np.array_equal(engF.read_slice(kind, x), engI.read_slice(kind, x))
    for y in range(Nv):
        bad += not np.array_equal(engF.read_slice(capi.VABCI, x, y), engI.read_slice(capi.VABCI, x, y))
        bad += not np.array_equal(engF.read_slice(capi.TABIJ, x, y), engI.read_slice(capi.TABIJ, x, y))
        if x <= y:
            bad += not np.array_equal(engF.read_slice(capi.VABIJ, x, y), engI.read_slice(capi.VABIJ, x, y))
print("slices differing between fill and ingest:", bad, flush=True)


def singles(eng, label):
    got = np.array([eng.run(i, 1)[0] for i in range(n)])
    rel = np.abs(got - want) / np.abs(want)
    idx = np.where(rel > 1e-11)[0]
    print(label, "tuples off:", len(idx), "max rel", float(rel.max()), "sum diff", float(got.sum() - want.sum()))
    for i in idx[:12]:
        print("   ", i, tl[i].tolist(), got[i], want[i], flush=True)
    return idx


singles(engF, "one tuple per run, auto nsplit")
for ns in ("1", "2", "3"):
    os.environ["ATRIP_B200_NSPLIT"] = ns
    singles(engF, "one tuple per run, nsplit " + ns)
os.environ.pop("ATRIP_B200_NSPLIT")

# batches of growing size from the start of the list
for cnt in (2, 3, 8, 37, 148, 149, 600, n):
    for ns in (None, "4"):
        if ns:
            os.environ["ATRIP_B200_NSPLIT"] = ns
        e, _ = engF.run(0, cnt)
        os.environ.pop("ATRIP_B200_NSPLIT", None)
        print("batch", cnt, "nsplit", ns or "auto", "diff", e - want[:cnt].sum(), flush=True)
# small device batches (many launches per run)
for b in (1, 7, 148):
    eng = mk("fill", batch_tuples=b)
    e, _ = eng.run()
    print("device batch", b, "full-run diff", e - want.sum(), flush=True)
    eng.close()
print("t", time.time() - t0)
