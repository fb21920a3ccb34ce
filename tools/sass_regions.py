#!/usr/bin/env python
"""Static SASS cost model of a kernel: instructions attributed to source regions.

    python tools/sass_regions.py [kernel-substring] [--lib PATH]

Extracts the sm_100a cubin from libkindel_b200.so, disassembles it with source line info (the build uses
-lineinfo) and counts, for one kernel, how many SASS instructions come from each source function of
kindel_b200/csrc/*.cu (by line range; inlined code is attributed to the function it was inlined from) and how
the code is laid out (contiguous runs = inlined instances).  No GPU needed.  Together with how often a region
runs per tile (reads per tile / window, flushes per window) this gives the instruction budget per base that an
issue-bound kernel lives on -- see DESIGN.md section 4."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def function_ranges(path):
    """[(first_line, last_line, name)] of the top-level-ish functions / lambdas of a .cu file (brace matching)."""
    out, stack = [], []
    sig = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:__global__|__device__|inline|static|auto)\b.*?\b([A-Za-z_]\w*)\s*(?:\(|=\s*\[)")
    pending = None
    depth = 0
    with open(path) as fh:
        for no, line in enumerate(fh, 1):
            code = line.split("//")[0]
            m = sig.match(code)
            if m and depth <= 2 and pending is None:
                pending = (no, m.group(1))
            elif pending is not None and pending[1] == "__launch_bounds__":  # kernel name on the next line
                m2 = re.match(r"\s*([A-Za-z_]\w*)\s*\(", code)
                if m2:
                    pending = (pending[0], m2.group(1))
            for ch in code:
                if ch == "{":
                    depth += 1
                    if pending is not None and not stack:
                        stack.append((pending[0], pending[1], depth))
                        pending = None
                elif ch == "}":
                    if stack and stack[-1][2] == depth:
                        first, name, _ = stack.pop()
                        out.append((first, no, name))
                    depth -= 1
            if pending is not None and code.strip().endswith(";"):
                pending = None
    return out


# section markers inside the tile-owner kernels (text that starts a section, label)
MARKERS = [("auto prefetch_raw", "prefetch lambda"), ("for (long long tile =", "tile setup"),
           ("---- sub-chunk [c0, c1)", "sub-chunk bounds + bulk copy"), ("metadata: all loads", "per-read metadata"),
           ("---- coverage of this warp", "coverage scan"), ("---- this warp's window against", "window search"),
           ("for (int base = a & ~7", "MAIN LOOP"), ("if (kFresh && !stored)", "flush call sites")]
_marks = {}


def kernel_section(path, line):
    if path not in _marks:
        found = []
        with open(path) as fh:
            for no, text in enumerate(fh, 1):
                for key, label in MARKERS:
                    if key in text:
                        found.append((no, label))
        _marks[path] = found
    label = "prologue"
    for no, lab in _marks[path]:
        if no <= line:
            label = lab
    return label


def main(argv):
    want = next((a for a in argv if not a.startswith("--")), "pileup_tiled_kernelILb1")
    lib = os.path.join(ROOT, "kindel_b200", "_lib", "libkindel_b200.so")
    if "--lib" in argv:
        lib = argv[argv.index("--lib") + 1]
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
        cubin = next(os.path.join(tmp, f) for f in sorted(os.listdir(tmp)) if f.startswith("api.") and f.endswith(".cubin"))
        text = subprocess.run(["nvdisasm", "--print-line-info", cubin], check=True, capture_output=True, text=True).stdout
    sections = re.split(r"^//-+ \.text\.(\S+) -+$", text, flags=re.M)
    kernels = {sections[i]: sections[i + 1] for i in range(1, len(sections) - 1, 2)}
    name = next((k for k in kernels if want in k), None)
    if name is None:
        raise SystemExit("no kernel matching %r; have: %s" % (want, ", ".join(kernels)))
    ranges = {}
    counts, runs = {}, []
    cur_file, cur_line = None, 0
    total = 0
    for line in kernels[name].splitlines():
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur_file, cur_line = m.group(1), int(m.group(2))
            continue
        if not re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
            continue
        if cur_file not in ranges:
            ranges[cur_file] = function_ranges(cur_file) if cur_file and os.path.exists(cur_file) else []
        region = "(" + os.path.basename(cur_file or "?") + ")"
        best = None
        for first, last, fn in ranges[cur_file]:
            if first <= cur_line <= last and (best is None or last - first < best[1] - best[0]):
                best, region = (first, last), fn
        if region.endswith("_kernel"):  # inside a kernel body: sub-regions by the section comments
            region += ": " + kernel_section(cur_file, cur_line)
        counts[region] = counts.get(region, 0) + 1
        total += 1
        if runs and runs[-1][0] == region:
            runs[-1][1] += 1
        else:
            runs.append([region, 1])
    print("kernel:", name)
    print("SASS instructions:", total)
    for region, n in sorted(counts.items(), key=lambda kv: -kv[1]):
        print("  %-28s %6d  %5.1f %%" % (region, n, 100.0 * n / total))
    print("layout (runs of >= 12 instructions):")
    for region, n in runs:
        if n >= 12:
            print("  %-28s %6d" % (region, n))


if __name__ == "__main__":
    main(sys.argv[1:])
