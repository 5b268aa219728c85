#!/usr/bin/env python
"""Prints the headline ncu metrics of every kernel in a report (ncu --page raw --csv)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active", "lts__t_bytes.sum",
        "smsp__average_warp_latency_issue_stalled_barrier", "launch__shared_mem_per_block_dynamic"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name", "?")[:90])
        for k in KEYS:
            if k in d:
                print(f"   {k} = {d[k]} {units[hdr.index(k)]}")
        stalls = {k: float(v) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and
                  k.endswith("_per_issue_active.ratio") and v}
        for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6]:
            print(f"   stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} = {v:.2f}")


if __name__ == "__main__":
    main()
