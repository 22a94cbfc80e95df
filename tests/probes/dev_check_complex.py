"""developer probe (GPU box): complex-field engine against the oracle, printing numbers instead of
asserting, so that one run localises a defect (stores / contraction / reduction / run loop)."""
import os
import sys
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import atrip_b200
from atrip_b200 import capi
from oracle.oracle import EPS_A, EPS_I, JABCI, JIJKA, TABIJ, TAI, VABCI, VABIJ, VIJKA, Oracle

o = Oracle()


def step(name, f):
    try:
        f()
    except Exception:
        print("FAILED", name)
        traceback.print_exc()


def slices():
    No, Nv, seed, scale = 5, 11, 5, 0.1
    t = o.inputs_z(No, Nv, seed=seed, scale=scale)
    for src in ("fill", "ingest"):
        eng = atrip_b200.Engine(No, Nv, field=1)
        if src == "fill":
            eng.fill_synthetic(seed, scale)
        else:
            eng.load_all(t[EPS_I], t[EPS_A], t[TAI], t[TABIJ], t[VABIJ], t[VIJKA], t[VABCI])
        for abc in [(0, 3, 7), (2, 2, 9), (4, 10, 10)]:
            S = o.tuple_slices_z(No, Nv, t, abc)
            a, b, c = abc
            d = {"TA": np.abs(eng.read_slice(capi.TA, a) - S["TA"]).max(),
                 "HB": np.abs(eng.read_slice(capi.VIJKA, b) - S["HB"]).max(),
                 "VAC": np.abs(eng.read_slice(capi.VABCI, a, c) - S["VAC"]).max(),
                 "VCB": np.abs(eng.read_slice(capi.VABCI, c, b) - S["VCB"]).max(),
                 "TBC": np.abs(eng.read_slice(capi.TABIJ, b, c) - S["TBC"]).max(),
                 "VABij": np.abs(eng.read_slice(capi.VABIJ, a, b) - S["VABij"]).max()}
            print("slices", src, abc, {k: float(v) for k, v in d.items()}, flush=True)
        eng.close()


def cubes():
    for No, Nv in [(4, 8), (13, 29), (40, 56)]:
        seed, scale = 2000 + No, 0.1
        t = o.inputs_z(No, Nv, seed=seed, scale=scale)
        eng = atrip_b200.Engine(No, Nv, field=1)
        eng.fill_synthetic(seed, scale)
        for abc in [(0, 1, 2), (0, 0, 1), (2, 2, Nv - 1), (1, Nv // 2, Nv - 2)]:
            e, _, T, Z = o.tuple_energy_z(No, Nv, t, abc, want_cubes=True)
            ge, gT, gZ = eng.tuple_debug(*abc)
            tm = np.abs(T).max()
            print("cubes", No, Nv, abc, "T.re", float(np.abs(gT.real - T.real).max() / tm), "T.im",
                  float(np.abs(gT.imag - T.imag).max() / tm), "Z", float(np.abs(gZ - Z).max() / tm), "e", ge, e,
                  flush=True)
        eng.close()


def runs():
    for No, Nv, J in [(4, 8, False), (5, 11, True)]:
        t = o.inputs_z(No, Nv, seed=7, scale=0.1, with_J=J)
        want = o.run_z(No, Nv, t)
        for src in ("fill", "ingest"):
            eng = atrip_b200.Engine(No, Nv, with_J=J, field=1)
            if src == "fill":
                eng.fill_synthetic(7, 0.1)
            else:
                eng.load_all(t[EPS_I], t[EPS_A], t[TAI], t[TABIJ], t[VABIJ], t[VIJKA], t[VABCI], t.get(JIJKA),
                             t.get(JABCI))
            eng.build_tuples(capi.GROUP_AND_SORT)
            e, ct = eng.run()
            print("run", No, Nv, J, src, -e, -ct, "want", want, flush=True)
            eng.close()


step("slices", slices)
step("cubes", cubes)
step("runs", runs)
