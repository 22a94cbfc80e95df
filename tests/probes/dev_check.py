"""developer smoke: engine vs oracle on a few sizes (GPU box)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle.oracle import Oracle, EPS_I, EPS_A, TAI, TABIJ, VABIJ, VIJKA, VABCI
import atrip_b200
from atrip_b200 import capi

o = Oracle()
sizes = [(4, 8), (8, 24), (10, 40), (16, 33), (13, 29)]
if len(sys.argv) > 1:
    sizes = [tuple(int(x) for x in s.split(",")) for s in sys.argv[1:]]
for No, Nv in sizes:
    t = o.inputs(No, Nv, seed=12345, scale=0.1)
    for mode in ("fill", "ingest"):
        eng = atrip_b200.Engine(No, Nv)
        if mode == "fill":
            eng.fill_synthetic(12345, 0.1)
        else:
            eng.load_all(t[EPS_I], t[EPS_A], t[TAI], t[TABIJ], t[VABIJ], t[VIJKA], t[VABCI])
        # slices
        bad = 0
        for x in (0, Nv - 1, Nv // 2):
            bad += not np.array_equal(eng.read_slice(capi.TA, x), o.slice_TA(No, Nv, t[TABIJ], x))
            bad += not np.array_equal(eng.read_slice(capi.VIJKA, x), o.slice_HHHA(No, Nv, t[VIJKA], x))
            for y in (0, Nv - 1, x):
                bad += not np.array_equal(eng.read_slice(capi.VABCI, x, y), o.slice_ABPH(No, Nv, t[VABCI], x, y))
                bad += not np.array_equal(eng.read_slice(capi.TABIJ, x, y), o.slice_ABHH(No, Nv, t[TABIJ], x, y))
                if x <= y:
                    bad += not np.array_equal(eng.read_slice(capi.VABIJ, x, y), o.slice_ABHH(No, Nv, t[VABIJ], x, y))
        worst = 0
        for abc in [(0, 1, 2), (0, 0, 1), (1, 1, 1 + 1), (0, Nv - 1, Nv - 1), (Nv - 3, Nv - 2, Nv - 1), (2, 5, 5), (3, 3, 7)]:
            e, ct, T, Z = o.tuple_energy(No, Nv, t, abc, want_cubes=True)
            ge, gT, gZ = eng.tuple_debug(*abc)
            dT = np.abs(gT - T).max() / np.abs(T).max()
            dZ = np.abs(gZ - Z).max() / np.abs(Z).max()
            de = abs(ge - e) / abs(e)
            worst = max(worst, dT, dZ, de)
            if max(dT, dZ, de) > 1e-12:
                print("   tuple", abc, "dT", dT, "dZ", dZ, "de", de, ge, e)
        n = eng.build_tuples(capi.GROUP_AND_SORT)
        t0 = time.time()
        E, ct = eng.run()
        dt = time.time() - t0
        Eo, _ = o.run(No, Nv, t)
        print(f"No {No} Nv {Nv} {mode}: slice mismatches {bad}, worst tuple rel {worst:.2e}, run E {-E!r} oracle {Eo!r} "
              f"rel {abs(-E - Eo) / abs(Eo):.2e} ({n} tuples, {dt:.3f}s) {eng.last_timing()}", flush=True)
        eng.close()
