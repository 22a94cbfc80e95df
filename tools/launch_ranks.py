"""Starts one process per rank of a C++ driver built against the MPI shim (include/shim/mpi.h,
multi-process mode): every rank gets RANK, WORLD_SIZE, LOCAL_RANK, ATRIP_SHIM_MPI=1 and a fresh
ATRIP_SHIM_MPI_DIR; waits for all of them and relays rank 0's output.

  python tools/launch_ranks.py <nranks> <program> [args...]
"""
import os
import shutil
import subprocess
import sys
import tempfile


def launch(nranks, cmd, timeout=600, env=None, cwd=None):
    """returns (max return code, stdout+stderr of rank 0, list of all outputs)"""
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="atrip_shim_mpi_", dir=base)
    procs = []
    try:
        for r in range(nranks):
            e = dict(os.environ, **(env or {}))
            e.update(RANK=str(r), WORLD_SIZE=str(nranks), LOCAL_RANK=str(r), ATRIP_SHIM_MPI="1", ATRIP_SHIM_MPI_DIR=d)
            procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=e, cwd=cwd))
        outs, rc = [], 0
        for p in procs:
            try:
                o, _ = p.communicate(timeout=timeout)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                o, _ = p.communicate()
                rc = max(rc, 124)
            outs.append(o)
            rc = max(rc, abs(p.returncode or 0))
        return rc, outs[0], outs
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    n = int(sys.argv[1])
    rc, out0, outs = launch(n, sys.argv[2:])
    sys.stdout.write(out0)
    if rc:
        for r, o in enumerate(outs[1:], 1):
            sys.stderr.write(f"---- rank {r}\n{o}")
    sys.exit(rc)
