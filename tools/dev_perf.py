"""developer perf probe: engine throughput on synthetic stores (GPU box)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import atrip_b200
from atrip_b200 import capi
No, Nv, ntup = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
t0 = time.time()
eng = atrip_b200.Engine(No, Nv, batch_tuples=batch)
eng.fill_synthetic(12345, 0.01)
print("create+fill %.2fs" % (time.time() - t0), flush=True)
t0 = time.time()
n = eng.build_tuples(capi.GROUP_AND_SORT)
print("tuples %d in %.2fs" % (n, time.time() - t0), flush=True)
for rep in range(3):
    first = (n // 3) if n > 3 * ntup else 0
    E, ct = eng.run(first, min(ntup, n - first))
    tm = eng.last_timing()
    fl = eng.flops_per_tuple * tm["tuples"]
    print(f"run {tm['tuples']} tuples: {tm['total_ms']:.2f} ms -> {fl / tm['total_ms'] / 1e9:.2f} TFLOP/s; "
          f"contract {tm['contract_ms']:.3f} ms/launch reduce {tm['reduce_ms']:.3f} ms/launch "
          f"launches {tm['contract_launches']}+{tm['reduce_launches']} E={E!r}", flush=True)
