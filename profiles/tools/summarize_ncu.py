#!/usr/bin/env python
"""Write a markdown summary of an `ncu --set full` report for profiles/ (run where ncu is installed).
usage: summarize_ncu.py <report.ncu-rep> <title> > profiles/<name>.md"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct"]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {title}\n")
    print("| metric | unit | " + " | ".join(f"launch {i + 1}" for i in range(len(data))) + " |")
    print("|---|---|" + "---|" * len(data))
    kn = hdr.index("Kernel Name")
    print("| kernel | | " + " | ".join(r[kn] for r in data) + " |")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    for w in ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
              "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor"):
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    for k, r in enumerate(data):     # warp-state breakdown: average warps per issue slot in each stall state
        st = sorted(((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h
                     and "not_issued" not in h), reverse=True)
        print(f"\nlaunch {k + 1}: warps per issue slot by state: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:10]))
    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    for k, r in enumerate(data):
        print(f"\nlaunch {k + 1}: DRAM traffic = {float(r[rd]):.1f} + {float(r[wr]):.1f} = "
              f"{float(r[rd]) + float(r[wr]):.1f} {units[rd]} (read + write)")


if __name__ == "__main__":
    main()
