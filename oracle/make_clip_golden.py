"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/clip_cases.json from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_clip_golden

For each of the deterministic clip-heavy alignments of tests/clip_cases.py it records what the reference
returns from `parse_bam` (sha256 of each contig's int32 [19, L+1] table, the layout of
tests/helpers.reference_alignment_to_table) and from `bam_to_consensus` under a per-seed option set
(realign mostly on; reference kindel/kindel.py:488-555): FASTA records, `changes`, report text.  The GPU box
has no reference tree; there the `-m gpu` tests compare the engine with this file, and the `-m "not gpu"`
tests compare oracle tables + host code with it.
"""
from __future__ import annotations

import hashlib
import json
import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refload  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "clip_cases.json")
N_CASES = 96


def options(seed):
    """Option set of a case: (realign, min_depth, min_overlap, clip_decay_threshold, mask_ends, trim_ends, uppercase)."""
    rng = random.Random(77_000 + seed)
    return [rng.random() < 0.85, rng.choice([1, 1, 2, 5]), rng.choice([1, 3, 7, 9]), rng.choice([0.0, 0.1, 0.1, 0.3, 0.9]),
            rng.choice([0, 1, 5, 20, 50]), rng.random() < 0.5, rng.random() < 0.3]


def table_sha256(table_i64) -> str:
    return hashlib.sha256(np.ascontiguousarray(table_i64, dtype=np.int32).tobytes()).hexdigest()


def report_body(text):
    """Report lines without the echo of the input path."""
    return [l for l in text.splitlines() if not l.startswith("- bam_path")]


def main():
    import clip_cases
    import helpers

    if not refload.available():
        raise SystemExit("the reference tree is not present; goldens can only be generated in the build container")
    k = refload.load_reference()
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        for seed in range(N_CASES):
            path = os.path.join(tmp, "clip%d.sam" % seed)
            with open(path, "wt") as fh:
                fh.write(clip_cases.clip_case(seed))
            opts = options(seed)
            alns = k.parse_bam(path)
            res = k.bam_to_consensus(path, *opts)
            cases.append({
                "seed": seed,
                "options": opts,
                "contigs": list(alns.keys()),
                "table_sha256": {name: table_sha256(helpers.reference_alignment_to_table(aln)[0])
                                 for name, aln in alns.items()},
                "fasta": [[r.name, r.sequence] for r in res.consensuses],
                "changes": {c: "".join("-" if x is None else x for x in ch) for c, ch in res.refs_changes.items()},
                "reports": {c: report_body(rep) for c, rep in res.refs_reports.items()},
            })
    with open(OUT, "wt") as fh:
        json.dump({"reference": "bede/kindel v1.2.1 @ 14d727b", "generator": "tests/clip_cases.py", "cases": cases},
                  fh, indent=0, sort_keys=True)
    patched = sum(any(l.startswith("- clip-dominant regions") and l.strip() != "- clip-dominant regions:"
                      for l in rep) for c in cases for rep in c["reports"].values())
    print("clip cases:", len(cases), "reports with merged clip-dominant regions:", patched, "bytes:", os.path.getsize(OUT))


if __name__ == "__main__":
    main()
