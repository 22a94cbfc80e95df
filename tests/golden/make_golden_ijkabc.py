"""Generates tests/golden/ijkabc_vectors.json from THE REFERENCE ITSELF (oracle/_ref, see
make_golden.py): whole runs with Input::ijkabc set (reference Atrip.cxx:183-187: Tai is negated;
:1108-1111: the final sign flip is skipped), F = double and F = Complex, with and without (cT).

  python tests/golden/make_golden_ijkabc.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, Reference  # noqa: E402

RUNS = [(4, 8, 12345, 0.1, False), (5, 11, 777, 0.05, True), (8, 16, 1, 0.04, False)]  # (No, Nv, seed, scale, with_J)
RUNS_Z = [(4, 8, 12345, 0.1, False), (5, 11, 777, 0.05, True)]


def main():
    o, r = Oracle(), Reference()
    out = {"generator": "tests/golden/make_golden_ijkabc.py",
           "source": "oracle/_ref/libatrip_ref.so (reference, unmodified), Input::ijkabc = true", "runs": [],
           "complex_runs": []}
    r.set_ijkabc(True)
    try:
        for No, Nv, seed, scale, J in RUNS:
            e, ct = r.run(No, Nv, o.inputs(No, Nv, seed=seed, scale=scale, with_J=J))
            out["runs"].append(dict(No=No, Nv=Nv, seed=seed, scale=scale, with_J=J, energy=e.hex(), ct_energy=ct.hex()))
            print("ijkabc run", No, Nv, seed, J, e, ct)
        for No, Nv, seed, scale, J in RUNS_Z:
            e, ct = r.run_z(No, Nv, o.inputs_z(No, Nv, seed=seed, scale=scale, with_J=J))
            out["complex_runs"].append(dict(No=No, Nv=Nv, seed=seed, scale=scale, with_J=J, energy=e.hex(),
                                            ct_energy=ct.hex()))
            print("ijkabc complex run", No, Nv, seed, J, e, ct)
    finally:
        r.set_ijkabc(False)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ijkabc_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("written")


if __name__ == "__main__":
    main()
