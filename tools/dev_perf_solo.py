"""developer perf probe (GPU box): the kernel shapes of a sharded configuration on ONE GPU.
Rank 0 of `nranks` fills its shard and runs tuples whose indices it all owns (a, b, c multiples of nranks),
so no slice has to travel: ATRIP_B200_SOLO_SHARD=1.  Used to take ncu captures of the c4 (No=100 Nv=1000)
contraction / reduction kernels, which do not fit a replica on one GPU.
  ATRIP_B200_SOLO_SHARD=1 python tools/dev_perf_solo.py No Nv nranks ntuples [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ATRIP_B200_SOLO_SHARD", "1")
import numpy as np
import atrip_b200
from atrip_b200 import capi
No, Nv, nranks, ntup = (int(x) for x in sys.argv[1:5])
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 0
t0 = time.time()
eng = atrip_b200.Engine(No, Nv, rank=0, nranks=nranks, resident=False, batch_tuples=batch)
eng.fill_synthetic(12345, 0.0002)
print("create+fill %.2fs" % (time.time() - t0), flush=True)
# lexicographic tuples over the indices rank 0 owns, c fastest (like the rank's own group-and-sort list:
# the home index varies fastest); start in the middle of the range
own = np.arange(0, Nv, nranks)
tl = []
ia = len(own) // 3
for ib in range(ia + 1, len(own)):
    for ic in range(ib + 1, len(own)):
        tl.append((own[ia], own[ib], own[ic]))
        if len(tl) >= 3 * ntup:
            break
    if len(tl) >= 3 * ntup:
        break
eng.set_tuples(np.array(tl, dtype=np.uint64))
n = eng.num_tuples()
for rep in range(3):
    E, ct = eng.run(rep * ntup, min(ntup, n - rep * ntup))
    tm = eng.last_timing()
    fl = eng.flops_per_tuple * tm["tuples"]
    print(f"run {tm['tuples']} tuples: {tm['total_ms']:.2f} ms -> {fl / tm['total_ms'] / 1e9:.2f} TFLOP/s; "
          f"contract {tm['contract_ms']:.3f} ms/launch reduce {tm['reduce_ms']:.3f} ms/launch "
          f"launches {tm['contract_launches']}+{tm['reduce_launches']} batch {eng.batch_tuples} E={E!r}", flush=True)
