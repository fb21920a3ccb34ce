#!/usr/bin/env python
"""Where a profiled tile-owner kernel spends its instructions and its time, by phase of the kernel body.

    python tools/ncu_regions.py report.ncu-rep            (no GPU needed; reads a saved ncu report)

Input: a report captured with `ncu --set full --import-source on` (the source travels inside the report).
The SASS of the kernel is walked in address order with the per-instruction "Instructions Executed" and warp
stall samples ncu recorded; each instruction is assigned to a phase of the kernel body by the source line it
came from (phase boundaries = the section comments of pileup_tiled.cu / pileup_wide.cu); helper code inlined
from shared functions (csa, add8, shuffles ...) takes the phase of the code around it.  Stall samples are
per-warp samples, i.e. a proxy for where warps spend TIME; instructions executed is where the issue slots go."""
import collections
import csv
import subprocess
import sys

# (text that starts the phase, label) -- in source order; matched against the source embedded in the report
PHASES = [("auto prefetch_raw", "prefetch next tile's metadata"), ("for (long long tile =", "tile setup"),
          ("while (c0 < hi)", "sub-chunk bounds + bulk copy"), ("int* diff = sm.diff[dbuf]", "per-read metadata"),
          ("// metadata + difference array complete", "barrier + mbarrier wait"),
          ("coverage of this warp", "coverage scan"), ("= lower_bound_warp(sm.gs", "window search"),
          ("for (int base = a", "MAIN LOOP"), ("c0 = c1;", "final flush")]
FLUSH_HELPERS = ("quarter_sum", "octet_sum", "extract8", "flush_window")


def main(argv):
    out = subprocess.run(["ncu", "-i", argv[0], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True, check=True).stdout
    hdr, cur_file, cur = None, None, None
    src, sass, kernel = {}, [], "?"
    for r in csv.reader(out.splitlines()):
        if len(r) == 2:
            if r[0] == "File Path":
                cur_file = r[1]
            elif r[0] == "Function Name":
                kernel = r[1]
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
            addr = int(r[2], 16)
        except ValueError:
            continue
        sass.append((addr, cur, dict(zip(hdr[4:], r[4:]))))
    sass.sort(key=lambda t: t[0])
    main_file = collections.Counter(f for (f, _), in [(k,) for k in src] if f.endswith(".cu")).most_common(1)[0][0]
    # the whole file as embedded in the report (the combined view lists only lines that own SASS)
    full = subprocess.run(["ncu", "-i", argv[0], "--page", "source", "--csv", "--print-source", "cuda"],
                          capture_output=True, text=True, check=True).stdout
    text_of, in_main = {}, False
    for r in csv.reader(full.splitlines()):
        if len(r) == 2 and r[0] == "File Name":
            in_main = r[1] == main_file
        elif len(r) == 2 and in_main and r[0].isdigit():
            text_of[int(r[0])] = r[1]
    marks = []
    for l in sorted(text_of):
        for key, label in PHASES:
            if key in text_of[l] and label not in [m[1] for m in marks]:
                marks.append((l, label))
    kernel_first = max(l for l, t in text_of.items() if "__global__" in t and l < min(m[0] for m in marks))
    helper_of, current = {}, None
    for l in sorted(text_of):
        t = text_of[l]
        if "__device__" in t or "__global__" in t or t.startswith(("struct ", "template ")):
            current = next((h for h in FLUSH_HELPERS if h + "(" in t or h + "_wide(" in t), None) if "__device__" in t else current
            if t.startswith("struct ") or "__global__" in t:
                current = None
        helper_of[l] = current

    def phase(key):
        f, l = key
        if f != main_file:
            return None
        if l >= kernel_first:
            label = "kernel prologue"
            for no, lab in marks:
                if no <= l:
                    label = lab
            return label
        return "flush" if helper_of.get(l) else None

    labels, last = [], "kernel prologue"
    for _, key, _ in sass:
        p = phase(key)
        if p is None:
            p = last
        last = p
        labels.append("flush" if p == "final flush" else p)
    stall_keys = [k for k in sass[0][2] if k.startswith("stall_") and "Not Issued" not in k]
    agg = collections.defaultdict(collections.Counter)
    for lab, (_, _, d) in zip(labels, sass):
        a = agg[lab]
        a["static"] += 1
        a["inst"] += int(d["Instructions Executed"])
        a["smp"] += int(d["# Samples"])
        for k in stall_keys:
            a[k] += int(d[k])
    ti = sum(a["inst"] for a in agg.values())
    ts = sum(a["smp"] for a in agg.values())
    print("kernel:", kernel)
    print("SASS: %d instructions static, %d executed (warp level), %d warp stall samples" % (len(sass), ti, ts))
    print("%-30s %7s %8s %8s   %s" % ("phase", "static", "issue %", "time %", "top stall reasons (share of the phase's samples)"))
    for lab, a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"]):
        top = sorted(((a[k], k) for k in stall_keys), reverse=True)[:4]
        print("%-30s %7d %7.1f%% %7.1f%%   %s" % (lab, a["static"], 100.0 * a["inst"] / ti, 100.0 * a["smp"] / max(ts, 1),
                                                  ", ".join("%s %.0f%%" % (k[6:], 100.0 * c / max(a["smp"], 1)) for c, k in top)))


if __name__ == "__main__":
    main(sys.argv[1:])
