#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full --import-source on) into the short text summary kept in profiles/.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor")


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        print("kernel:", d.get("Kernel Name", "?"))
        for h, u in zip(hdr, units):
            if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                print("  %-78s %-10s %s" % (h, u, d[h]))
    src = page(rep, "source")
    if len(src) > 2:
        hdr, data = src[1], src[2:]
        ix = {h: i for i, h in enumerate(hdr)}

        def f(r, k):
            try:
                return float(r[ix[k]])
            except Exception:
                return 0.0

        tot = sum(f(r, "# Samples") for r in data) or 1.0
        print("\nwarp-stall sampling: %d samples, %d instructions executed" % (tot, sum(f(r, "Instructions Executed") for r in data)))
        for k in ("stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_math", "stall_not_selected",
                  "stall_selected", "stall_branch_resolving", "stall_no_inst", "stall_mio", "stall_lg"):
            print("  %-24s %5.1f %%" % (k, 100 * sum(f(r, k) for r in data) / tot))
        print("\ntop SASS instructions by samples:")
        for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:12]:
            print("  %5.1f %%  %s" % (100 * f(r, "# Samples") / tot, r[ix["Source"]].strip()[:90]))


if __name__ == "__main__":
    main(sys.argv[1])
