#!/usr/bin/env python
"""Where a profiled kernel spends its warp time and its issue slots, by SOURCE LINE.

    python tools/ncu_lines.py report.ncu-rep [top_n]        (no GPU needed; reads a saved ncu report)

Input: a report captured with `ncu --set full --import-source on` of a kernel built with -lineinfo.  Every SASS
instruction carries the warp stall samples and "Instructions Executed" ncu recorded; they are summed per source line
(file:line of the innermost inlined frame ncu attributes the instruction to)."""
import collections
import csv
import subprocess
import sys


def main(argv):
    out = subprocess.run(["ncu", "-i", argv[0], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True, check=True).stdout
    top_n = int(argv[1]) if len(argv) > 1 else 40
    hdr, cur_file, cur = None, None, None
    src = {}
    agg = collections.defaultdict(collections.Counter)
    for r in csv.reader(out.splitlines()):
        if len(r) == 2:
            if r[0] == "File Path":
                cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r:
            continue
        if r[0] not in ("", "-"):
            cur = (cur_file, int(r[0]))
            src[cur] = r[1]
            continue
        try:
            int(r[2], 16)
        except (ValueError, IndexError):
            continue
        d = dict(zip(hdr[4:], r[4:]))
        a = agg[cur]
        a["inst"] += int(d.get("Instructions Executed", 0) or 0)
        a["smp"] += int(d.get("# Samples", 0) or 0)
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                a[k] += int(v or 0)
    ti = sum(a["inst"] for a in agg.values()) or 1
    ts = sum(a["smp"] for a in agg.values()) or 1
    print("%d executed warp instructions, %d stall samples" % (ti, ts))
    print("%-26s %7s %7s  %-34s %s" % ("file:line", "time %", "issue %", "top stalls", "source"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top_n]:
        top = sorted(((v, k[6:]) for k, v in a.items() if k.startswith("stall_")), reverse=True)[:2]
        print("%-26s %6.1f%% %6.1f%%  %-34s %s" % ("%s:%d" % key, 100.0 * a["smp"] / ts, 100.0 * a["inst"] / ti,
                                                   ", ".join("%s %.0f%%" % (k, 100.0 * v / max(a["smp"], 1)) for v, k in top),
                                                   src.get(key, "").strip()[:90]))


if __name__ == "__main__":
    main(sys.argv[1:])
