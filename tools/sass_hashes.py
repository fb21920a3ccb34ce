#!/usr/bin/env python
"""Per-kernel SASS hashes of libkindel_b200.so.

    python tools/sass_hashes.py                 print  "<sha256[:32]>  <instructions>  <mangled name>" per kernel
    python tools/sass_hashes.py --write FILE    rewrite the hash lines of FILE, keeping its '#' header

A hash covers the instruction text and encodings of `cuobjdump -sass` with column padding normalised, so it
changes exactly when the machine code of that kernel changes.  profiles/r01_kernel_sass_hashes.txt records the
build whose default kernels ran on the GPU; tests/test_abi.py compares the kernels marked there as validated."""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kindel_b200", "_lib", "libkindel_b200.so")


def kernel_hashes(lib=LIB):
    text = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : (\S+)\n", text)
    out = {}
    for i in range(1, len(parts) - 1, 2):
        body = [" ".join(re.sub(r"/\*[0-9a-f]{4}\*/", "", l).split())
                for l in parts[i + 1].splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        out[parts[i]] = (hashlib.sha256("\n".join(body).encode()).hexdigest()[:32], len(body))
    return out


def read_recorded(path):
    rec = {}
    with open(path) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            digest, count, name = line.split()
            rec[name] = (digest, int(count))
    return rec


def main(argv):
    hashes = kernel_hashes()
    lines = ["%s  %5d  %s" % (h, n, name) for name, (h, n) in sorted(hashes.items())]
    if "--write" in argv:
        path = argv[argv.index("--write") + 1]
        header = [l.rstrip("\n") for l in open(path) if l.startswith("#")] if os.path.exists(path) else []
        with open(path, "w") as fh:
            fh.write("\n".join(header + lines) + "\n")
    else:
        print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1:])
