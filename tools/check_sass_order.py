"""Checks the consumer-side release of the contraction kernel's TMA ring in the compiled SASS.

A consumer warp may hand a ring stage back to the producer (mbarrier arrive on the stage's "empty"
barrier, SASS `SYNCS.ARRIVE.TRANS64.A1T0`) only after every LDS of that stage has delivered its data.
The source orders  fragment loads + DMMAs -> fence.proxy.async (every lane) -> __syncwarp() -> arrive;
the proxy fence (SASS `MEMBAR.ALL.CTA` + `FENCE.VIEW.ASYNC.S`) is the ordering construct.  Before it
existed ptxas was free to move the arrive (no register dependency): in a straight-line loop body it
placed the arrive between the last LDS and the DMMAs that consume them, and the kernel then produced
run-to-run different cubes (profiles/r01_ring_release_race.txt).  This script is the second guard: for
every contract_kernel instantiation it asserts
    last LDS < FENCE.VIEW.ASYNC (< WARPSYNC, where ptxas kept it) < arrive     and     last DMMA < arrive
(the latter alone is what made the unfenced kernel safe: in-order issue, DMMAs wait for their LDS
operands).

  python tools/check_sass_order.py [path/to/libatrip_b200.so]      exit code 1 on a violation
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    bad, seen = [], 0
    for chunk in out.split("Function : ")[1:]:
        name = chunk.split("\n", 1)[0].strip()
        if "contract_kernel" not in name:
            continue
        seen += 1
        ins = [l for l in chunk.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
        arrive = [i for i, l in enumerate(ins) if "SYNCS.ARRIVE.TRANS64.A1T0" in l]
        dmma = [i for i, l in enumerate(ins) if "DMMA" in l]
        lds = [i for i, l in enumerate(ins) if re.search(r"\bLDS", l)]
        wsync = [i for i, l in enumerate(ins) if "WARPSYNC" in l]
        fence = [i for i, l in enumerate(ins) if "FENCE.VIEW.ASYNC" in l]
        ok = len(arrive) == 1 and dmma and lds
        if ok:
            a = arrive[0]
            fbefore = [f for f in fence if max(lds) < f < a]
            # the proxy fence sits between the last fragment load and the arrive ...
            ok = max(dmma) < a and max(lds) < a and bool(fbefore)
            # ... and in front of the warp rendezvous where ptxas kept one.  The source always has the
            # __syncwarp(); ptxas drops the WARPSYNC only where it has proven the warp converged (after
            # setmaxnreg.sync.aligned, or when the loop body holds nothing but predicated instructions
            # behind an earlier rendezvous) -- all lanes then issue every LDS / DMMA / arrive as one
            # instruction, in order.
            between = [w for w in wsync if max(lds) < w < a]
            ok = ok and all(min(fbefore) < w for w in between)
        if not ok:
            bad.append(name)
    return seen, bad


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "atrip_b200", "csrc", "libatrip_b200.so")
    seen, bad = check(lib)
    print(f"{seen} contract_kernel instantiations checked, {len(bad)} with the stage release not after all DMMAs")
    for b in bad:
        print("  VIOLATION:", b)
    sys.exit(1 if bad or not seen else 0)
