#!/usr/bin/env python
"""Are the kernels of two builds the same machine code?   python tools/sass_same.py before.sass after.sass

Inputs are `cuobjdump -sass lib.so` dumps.  Kernels are matched by mangled name (a template that gained a
defaulted bool parameter, `...ILb1EE` -> `...ILb1ELb0EE`, is matched to its old name) and compared instruction by
instruction including the encodings, ignoring only column padding.  Used to show that an experimental
instantiation or a host-emulation guard left the GPU-validated kernels untouched."""
import re
import sys


def kernels(path):
    parts = re.split(r"\n\s*Function : (\S+)\n", open(path).read())
    out = {}
    for i in range(1, len(parts) - 1, 2):
        out[parts[i]] = [" ".join(re.sub(r"/\*[0-9a-f]{4}\*/", "", l).split())
                         for l in parts[i + 1].splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    return out


def main(argv):
    a, b = kernels(argv[0]), kernels(argv[1])
    bad = 0
    for name in sorted(a):
        new = name
        for pat, rep in ((None, None), (r"ILb([01])EE", r"ILb\1ELb0EE"), (r"ILb([01])EE", r"ILb\1ENS_6WsCfg1EE")):
            cand = name if pat is None else re.sub(pat, rep, name)
            if cand in b:
                new = cand
                break
        if new not in b:
            print("GONE      ", name)
            bad += 1
            continue
        same = a[name] == b[new]
        bad += not same
        print("%-10s %s%s" % ("same" if same else "DIFFERENT", name, "" if new == name else "  (now " + new + ")"))
    matched = set(a) | {re.sub(r"ILb([01])EE", r"ILb\1ELb0EE", n) for n in a} | {re.sub(r"ILb([01])EE", r"ILb\1ENS_6WsCfg1EE", n) for n in a}
    for name in sorted(set(b) - matched):
        print("new        %s (%d instructions)" % (name, len(b[name])))
    return 1 if bad else 0


if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1:]))
