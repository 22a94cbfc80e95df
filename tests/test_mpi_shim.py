"""CPU, world_size 3: the multi-process mode of the MPI stand-in (include/shim/mpi.h) that lets the
C++ API (atrip::Atrip::init / run, atrip_b200/host) run one rank per GPU on a box without MPI: rank and
size from the launcher's environment, MPI_Bcast / MPI_Allreduce / MPI_Reduce / MPI_Allgather /
MPI_Barrier across the processes.  Serial behaviour (np = 1) when the mode is not switched on."""
import os
import subprocess
import sys
import tempfile

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
from launch_ranks import launch  # noqa: E402

SRC = r"""
#include <mpi.h>
#include <cstdio>
#include <cstring>
int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  int r, n;
  MPI_Comm_rank(MPI_COMM_WORLD, &r);
  MPI_Comm_size(MPI_COMM_WORLD, &n);
  unsigned char id[128];
  for (int i = 0; i < 128; i++) id[i] = (unsigned char)(r == 1 ? 7 * i + 3 : 0);   // root 1 owns the payload
  MPI_Bcast(id, 128, MPI_BYTE, 1, MPI_COMM_WORLD);
  int ok = 1;
  for (int i = 0; i < 128; i++) ok &= id[i] == (unsigned char)(n > 1 ? 7 * i + 3 : 0);
  double v[2] = {0.1 * (r + 1), -1.0 * r}, s[2], m[2];
  for (int rep = 0; rep < 50; rep++) {                                              // many collectives in a row
    MPI_Allreduce(v, s, 2, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    MPI_Barrier(MPI_COMM_WORLD);
  }
  MPI_Reduce(v, m, 2, MPI_DOUBLE, MPI_MAX, 0, MPI_COMM_WORLD);
  int mine = 10 + r, all[16] = {0};
  MPI_Allgather(&mine, 1, MPI_INT, all, 1, MPI_INT, MPI_COMM_WORLD);
  int self_n = -1;
  MPI_Comm_size(MPI_COMM_SELF, &self_n);
  std::printf("rank %d of %d bcast %d sum %a %a max %a gather", r, n, ok, s[0], s[1], m[0]);
  for (int i = 0; i < n; i++) std::printf(" %d", all[i]);
  std::printf(" self %d\n", self_n);
  MPI_Finalize();
  return 0;
}
"""


def build(tmp):
    src, exe = os.path.join(tmp, "t.cxx"), os.path.join(tmp, "t")
    open(src, "w").write(SRC)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include", "shim"), "-o", exe, src])
    return exe


def test_serial_by_default():
    with tempfile.TemporaryDirectory() as tmp:
        exe = build(tmp)
        env = dict(os.environ, RANK="2", WORLD_SIZE="4")  # a launcher's variables alone do not switch it on
        env.pop("ATRIP_SHIM_MPI", None)
        out = subprocess.check_output([exe], text=True, env=env)
        assert out.startswith("rank 0 of 1 bcast 1 ") and out.strip().endswith("gather 10 self 1"), out


def test_three_ranks_collectives():
    with tempfile.TemporaryDirectory() as tmp:
        exe = build(tmp)
        rc, _, outs = launch(3, [exe], timeout=120)
        assert rc == 0, outs
        want_sum = (0.1 * 1 + 0.1 * 2) + 0.1 * 3  # rank order
        for r, o in enumerate(outs):
            f = o.split()
            assert f[:6] == ["rank", str(r), "of", "3", "bcast", "1"], o
            assert float.fromhex(f[7]) == want_sum and float.fromhex(f[8]) == -3.0, o
            assert float.fromhex(f[10]) == 0.1 * 3 and f[11:] == ["gather", "10", "11", "12", "self", "1"], o
