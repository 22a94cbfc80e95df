"""Generates tests/golden/reference_vectors.json from THE REFERENCE ITSELF.

Run in a container that has /root/reference (after `make -C oracle`): the values below are
outputs of the reference's own unmodified sources (oracle/_ref/libatrip_ref.so built by
oracle/Makefile from /root/reference/src/atrip/*.cxx), fed with the counter-based synthetic
inputs of oracle_fill (specification in DESIGN.md).  The GPU box has no /root/reference, so the
vectors are committed; floats are stored as C99 hex strings (bit exact).

  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import (EPS_A, EPS_I, TAI, Oracle, Reference)  # noqa: E402

RUNS = [  # (No, Nv, seed, scale, with_J)
    (4, 8, 12345, 0.1, False), (4, 8, 777, 0.1, True), (5, 11, 12345, 0.05, False), (8, 16, 1, 0.04, False),
    (7, 13, 99, 0.05, True), (10, 40, 12345, 0.01, False), (16, 24, 5, 0.01, False),
]
TUPLES = [  # (No, Nv, seed, scale, [(a,b,c)...])  per-tuple values through the reference L1 functions
    (10, 40, 12345, 0.1, [(0, 1, 2), (0, 0, 1), (3, 3, 7), (2, 5, 5), (0, 39, 39), (37, 38, 39), (11, 17, 29)]),
    (13, 29, 4242, 0.1, [(0, 1, 2), (5, 5, 6), (5, 6, 6), (26, 27, 28), (1, 14, 28)]),
    (33, 40, 7, 0.05, [(0, 1, 2), (4, 4, 9), (4, 9, 9), (10, 20, 30)]),
    (40, 48, 3, 0.05, [(0, 1, 2), (1, 1, 47), (5, 17, 33)]),
]
RUNS_Z = [  # F = Complex: (No, Nv, seed, scale, with_J) through Atrip::run<Complex>
    (4, 8, 12345, 0.1, False), (4, 8, 777, 0.1, True), (5, 11, 12345, 0.05, False), (7, 13, 99, 0.05, True),
    (10, 24, 12345, 0.02, False), (16, 24, 5, 0.01, False),
]
TUPLES_Z = [  # F = Complex: per-tuple values through the reference L1 functions
    (10, 24, 12345, 0.1, [(0, 1, 2), (0, 0, 1), (3, 3, 7), (2, 5, 5), (0, 23, 23), (21, 22, 23), (5, 11, 17)]),
    (13, 29, 4242, 0.1, [(0, 1, 2), (5, 5, 6), (5, 6, 6), (26, 27, 28)]),
    (33, 36, 7, 0.05, [(0, 1, 2), (4, 4, 9), (10, 20, 30)]),
]
DISTS = [(8, 1), (13, 2), (13, 3), (21, 4), (40, 8), (33, 5)]  # (Nv, n_nodes)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    o, r = Oracle(), Reference()
    out = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libatrip_ref.so (reference, unmodified)",
           "runs": [], "tuples": [], "distributions": []}
    for No, Nv, seed, scale, J in RUNS:
        t = o.inputs(No, Nv, seed=seed, scale=scale, with_J=J)
        e, ct = r.run(No, Nv, t)
        out["runs"].append(dict(No=No, Nv=Nv, seed=seed, scale=scale, with_J=J, energy=e.hex(), ct_energy=ct.hex()))
        print("run", No, Nv, seed, J, e, ct)
    for No, Nv, seed, scale, tl in TUPLES:
        t = o.inputs(No, Nv, seed=seed, scale=scale)
        rec = dict(No=No, Nv=Nv, seed=seed, scale=scale, tuples=[])
        for abc in tl:
            S = o.tuple_slices(No, Nv, t, abc)
            T = r.doubles(No, Nv, S)
            Z = r.singles(No, Nv, abc, t[TAI], S, T)
            epsabc = float(t[EPS_A][abc[0]] + t[EPS_A][abc[1]] + t[EPS_A][abc[2]])
            same = (abc[0] == abc[1]) != (abc[1] == abc[2])
            e = (r.energy_same if same else r.energy_distinct)(epsabc, No, t[EPS_I], T, Z)
            idx = [0, 1, No, No * No, No ** 3 // 2, No ** 3 - 1]
            rec["tuples"].append(dict(abc=list(abc), energy=e.hex(), Tsum=float(T.sum()).hex(),
                                      Zsum=float(Z.sum()).hex(), Tabsmax=float(np.abs(T).max()).hex(),
                                      Tsample=[float(T[i]).hex() for i in idx],
                                      Zsample=[float(Z[i]).hex() for i in idx]))
            print("tuple", No, Nv, abc, e)
        out["tuples"].append(rec)
    out["complex_runs"], out["complex_tuples"] = [], []
    for No, Nv, seed, scale, J in RUNS_Z:
        t = o.inputs_z(No, Nv, seed=seed, scale=scale, with_J=J)
        e, ct = r.run_z(No, Nv, t)
        out["complex_runs"].append(dict(No=No, Nv=Nv, seed=seed, scale=scale, with_J=J, energy=e.hex(),
                                        ct_energy=ct.hex()))
        print("complex run", No, Nv, seed, J, e, ct)
    for No, Nv, seed, scale, tl in TUPLES_Z:
        t = o.inputs_z(No, Nv, seed=seed, scale=scale)
        rec = dict(No=No, Nv=Nv, seed=seed, scale=scale, tuples=[])
        for abc in tl:
            S = o.tuple_slices_z(No, Nv, t, abc)
            T = r.doubles_z(No, Nv, S)
            Z = r.singles_z(No, Nv, abc, t[TAI], S, T)
            epsabc = float((t[EPS_A][abc[0]] + t[EPS_A][abc[1]] + t[EPS_A][abc[2]]).real)
            same = (abc[0] == abc[1]) != (abc[1] == abc[2])
            e = (r.energy_same_z if same else r.energy_distinct_z)(epsabc, No, t[EPS_I], T, Z)
            idx = [0, 1, No, No * No, No ** 3 // 2, No ** 3 - 1]
            hx = lambda z: [float(z.real).hex(), float(z.imag).hex()]
            rec["tuples"].append(dict(abc=list(abc), energy=e.hex(), Tsum=hx(T.sum()), Zsum=hx(Z.sum()),
                                      Tabsmax=float(np.abs(T).max()).hex(),
                                      Tsample=[hx(T[i]) for i in idx], Zsample=[hx(Z[i]) for i in idx]))
            print("complex tuple", No, Nv, abc, e)
        out["complex_tuples"].append(rec)
    for Nv, n in DISTS:
        rec = dict(Nv=Nv, n_nodes=n, nodes=[])
        for me in range(n):
            tl = r.group_and_sort(n, me, Nv)
            rec["nodes"].append(dict(count=len(tl), sha256=digest(tl.astype(np.uint64)),
                                     head=tl[:4].tolist(), tail=tl[-2:].tolist()))
        out["distributions"].append(rec)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("written")


if __name__ == "__main__":
    main()
