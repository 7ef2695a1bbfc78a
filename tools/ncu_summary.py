#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/x/prof.ncu-rep profiles/x_prof_summary.txt [launch_index]"""
import csv
import io
import subprocess
import sys

KEEP = ("Kernel Name", "gpu__time_duration", "dram__bytes", "gpu__dram_throughput", "sm__warps_active", "launch__",
        "sm__throughput", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__issue_active", "smsp__thread_inst_executed_per_inst_executed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate", "sm__pipe_fp64", "sm__inst_executed_pipe", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "l1tex__t_sector_hit_rate", "smsp__average_warp", "smsp__warp_issue_stalled",
        "smsp__pcsamp_warps_issue_stalled", "sm__sass_thread_inst_executed_op", "smsp__sass_inst_executed_op_shared")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    d = data[idx]
    with open(out, "w") as f:
        f.write(f"# {rep}: ncu --set full --clock-control none, captured launch {idx} of {len(data)} "
                f"(the .ncu-rep itself is scratch, not committed)\n")
        for i, h in enumerate(hdr):
            if any(h.startswith(k) for k in KEEP):
                f.write(f"{h} = {d[i]} {units[i]}\n")


if __name__ == "__main__":
    main()
