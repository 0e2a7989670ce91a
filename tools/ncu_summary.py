#!/usr/bin/env python
"""Key throughput / stall metrics of every captured launch of an ncu report (dev tool).
usage: tools/ncu_summary.py REPORT.ncu-rep [KERNEL_REGEX]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"]
    if len(sys.argv) > 2:
        cmd += ["--kernel-name", f"regex:{sys.argv[2]}"]
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("=====", d.get("Kernel Name", "?")[:60])
        for k in KEYS:
            if k in d:
                print(f"  {k} = {d[k]}")
        st = {k: float(v) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
        print("  top stalls:", ", ".join(f"{k.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for k, v in top))


if __name__ == "__main__":
    main()
