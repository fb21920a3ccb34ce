#!/usr/bin/env python
"""Opcode census of libkindel_b200.so (no GPU needed): per kernel, how many SASS instructions of the kinds that
prove what the DESIGN claims -- UBLKCP (1-D bulk copy = TMA), SYNCS (mbarrier), LDGSTS (cp.async), USETMAXREG,
BAR (named barriers), RED / ATOM (global atomics), ATOMS (shared atomics), SHFL, LOP3 (the bit-sliced adders),
SHF (funnel shifts), and that there is no tensor-core op (UTCMMA / HMMA: there is no contraction in this path).

    python tools/opcode_census.py > profiles/r02_opcode_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kindel_b200", "_lib", "libkindel_b200.so")
KINDS = ["UBLKCP", "SYNCS", "LDGSTS", "USETMAXREG", "BAR", "REDG", "ATOMG", "ATOMS", "SHFL", "LOP3", "SHF", "LDS", "STS",
         "LDG", "STG", "UTCMMA", "HMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernel = None
    counts = collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kernel = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            kernel = re.sub(r"\(.*", "", kernel)
            counts[kernel] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and kernel:
            op = m.group(1)
            counts[kernel]["total"] += 1
            for k in KINDS:
                if op == k or op.startswith(k + "."):
                    counts[kernel][k] += 1
    print("libkindel_b200.so: arch %s, %d kernels" % (",".join(arch), len(counts)))
    print("%-64s %6s " % ("kernel", "total") + " ".join("%6s" % k[:6] for k in KINDS))
    for k, c in counts.items():
        print("%-64s %6d " % (k[-64:], c["total"]) + " ".join("%6d" % c[x] for x in KINDS))
    return 0


if __name__ == "__main__":
    sys.exit(main())
