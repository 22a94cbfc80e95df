"""Summarise an .ncu-rep (read here, no GPU needed) into the lines kept under profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(path, note=""):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        if note:
            print(note)
        for k in head:
            if k in KEYS or "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
                v = d.get(k, "")
                if "issue_stalled" in k:
                    try:
                        if float(v) < 0.3:
                            continue
                    except ValueError:
                        continue
                print(f"{k} [{u.get(k, '')}]: {v}")
        print()


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
