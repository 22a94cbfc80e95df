"""developer probe (GPU box): is a run-to-run difference born in the contraction (cube checksum
changes) or in the reduction (checksum constant, energy changes)?  Repeats the same batch."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import atrip_b200
    from atrip_b200 import capi
    No, Nv, field, reps = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    eng = atrip_b200.Engine(No, Nv, field=field)
    eng.fill_synthetic(5, 0.01)
    n = eng.build_tuples(capi.GROUP_AND_SORT)
    cnt = min(n, eng.batch_tuples)
    for ns in ("1", "4"):
        os.environ["ATRIP_B200_NSPLIT"] = ns
        res = {}
        for r in range(reps):
            e, _ = eng.run(0, cnt)
            key = (repr(e), hex(eng.cubes_checksum()))
            res[key] = res.get(key, 0) + 1
        print(f"  No {No} Nv {Nv} field {field} KPAD {os.environ.get('ATRIP_B200_KPAD', '0')} kp {eng.kp} "
              f"tuples {cnt} nsplit {ns}: {len(res)} distinct (energy, cube checksum) in {reps} runs", flush=True)
        for k, v in sorted(res.items(), key=lambda x: -x[1])[:6]:
            print("     ", v, "x", k, flush=True)
    eng.close()
    sys.exit(0)

CASES = [(16, 24, 1, "0"), (16, 24, 1, "1"), (16, 48, 0, "0"), (16, 24, 0, "0"), (10, 24, 1, "0"), (24, 40, 1, "0")]
if len(sys.argv) > 1 and sys.argv[1] == "short":
    CASES = [(16, 24, 1, "0"), (16, 48, 0, "0"), (24, 40, 1, "0")]
for No, Nv, field, kpad in CASES:
    env = dict(os.environ, ATRIP_B200_KPAD=kpad)
    p = subprocess.run([sys.executable, __file__, "child", str(No), str(Nv), str(field), "40"], env=env,
                       capture_output=True, text=True, timeout=60)
    print(p.stdout + p.stderr[-2000:], flush=True)
