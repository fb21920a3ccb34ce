"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

For every BAM/SAM fixture of the reference's own test-suite and for a set of synthetic edge-case alignments
 it
  1. re-encodes the input with this repo's own writers (names / qualities stripped, so the files
     are not copies of the reference's fixtures) into tests/golden/inputs/,
  2. runs the reference's `parse_bam`, `bam_to_consensus` (plain and --realign), `weights` and
     `features` on the RE-ENCODED file through oracle/refload.py,
  3. stores what the reference returned in tests/golden/<name>.npz (+ a JSON manifest).
The GPU box has no /root/reference: there the `-m gpu` parity tests compare the engine with these
files, and `-m "not gpu"` tests compare the C oracle with them.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refload, samdecode  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
INPUTS = os.path.join(GOLDEN, "inputs")

FIXTURES = [  # (golden name, path under /root/reference/tests): all 17 BAM/SAM files of the reference's suite
    ("bwa_1_1", "data_bwa_mem/1.1.sub_test.bam"),
    ("bwa_2_1", "data_bwa_mem/2.1.sub_test.bam"),
    ("bwa_3_1", "data_bwa_mem/3.1.sub_test.bam"),
    ("bwa_4_1", "data_bwa_mem/4.1.sub_test.bam"),
    ("bwa_5_1", "data_bwa_mem/5.1.sub_test.bam"),
    ("bwa_6_1", "data_bwa_mem/6.1.sub_test.bam"),
    ("seg_1_1", "data_segemehl/1.1.sub_test.bam"),
    ("seg_2_1", "data_segemehl/2.1.sub_test.bam"),
    ("seg_3_1", "data_segemehl/3.1.sub_test.bam"),
    ("seg_4_1", "data_segemehl/4.1.sub_test.bam"),
    ("seg_5_1", "data_segemehl/5.1.sub_test.bam"),
    ("seg_6_1", "data_segemehl/6.1.sub_test.bam"),
    ("mm2_multi", "data_minimap2/1.1.multi.bam"),
    ("mm2_gp120", "data_minimap2/hxb2-gp120-mutated.bam"),
    ("ext_1_debug", "data_ext/1.issue23.debug.sam"),
    ("ext_2_bc63", "data_ext/2.issue23.bc63.sam"),
    ("ext_3_bc75", "data_ext/3.issue23.bc75.sam"),
]

# the 6.1 Mb fixture: its dense table is 463 MB, so the golden is a digest (sha256 of the int32 table the
# reference's alignment converts to, of the consensus FASTA and of `changes`) plus column sums
DIGEST_FIXTURES = [("bact_tiny", "data_minimap2_bact/bact.tiny.bam")]

_OPS = "MIDNSHP=X"


def _cigar_text(cigars):
    if len(cigars) == 1 and cigars[0][1] is None:
        return "*"
    return "".join("%d%s" % (n, op) for n, op in cigars)


def reencode(src, dst_stem):
    """Write the records of `src` again with our own writers; returns the new path."""
    from kindel_b200 import bamio

    header, records = samdecode.read_alignment_file(src)
    contigs = [(sn[3:], int(next(f for f in fields if f.startswith("LN:"))[3:]))
               for sn, fields in header["@SQ"].items()]
    index = {name: i for i, (name, _) in enumerate(contigs)}
    if str(src).endswith(".sam"):
        dst = dst_stem + ".sam"
        with open(dst, "wt") as fh:
            fh.write("@HD\tVN:1.6\n")
            for name, ln in contigs:
                fh.write("@SQ\tSN:%s\tLN:%d\n" % (name, ln))
            for k, r in enumerate(records):
                fh.write("r%d\t%d\t%s\t%d\t60\t%s\t*\t0\t0\t%s\t*\n"
                         % (k, r.flag, r.rname, r.pos, _cigar_text(r.cigars), r.seq))
        return dst
    dst = dst_stem + ".bam"
    recs = []
    for r in records:
        words = [] if r.cigars[0][1] is None else [(n << 4) | _OPS.index(op) for n, op in r.cigars]
        recs.append((index.get(r.rname, -1), r.pos - 1, r.flag, words, r.seq))
    bamio.write_bam(dst, contigs, recs, level=9)
    return dst


def table_of(aln):
    from helpers import reference_alignment_to_table

    return reference_alignment_to_table(aln)


def frame_arrays(df, prefix):
    out = {}
    for col in df.columns:
        v = df[col].to_numpy()
        out[prefix + col] = v.astype("U") if v.dtype == object or v.dtype.kind in "UT" else v
    out[prefix + "__columns"] = np.array(list(df.columns))
    return out


def golden_for_file(k, path):
    """Everything the reference says about one alignment file -> dict of arrays + manifest entry."""
    alns = k.parse_bam(path)
    arrays, manifest = {}, {"contigs": []}
    for c, (name, aln) in enumerate(alns.items()):
        t, ins = table_of(aln)
        arrays["c%d_counts" % c] = t.astype(np.int32)
        arrays["c%d_consensus_depth" % c] = np.asarray(aln.consensus_depth, dtype=np.int64)
        arrays["c%d_clip_start_depth" % c] = np.asarray(aln.clip_start_depth, dtype=np.int64)
        arrays["c%d_clip_end_depth" % c] = np.asarray(aln.clip_end_depth, dtype=np.int64)
        arrays["c%d_clip_depth" % c] = np.asarray(aln.clip_depth, dtype=np.int64)
        manifest["contigs"].append({
            "name": name, "ref_len": len(aln.weights),
            "insertions": [[i, list(d.items())] for i, d in enumerate(ins) if d],
        })
    runs = {}
    for tag, realign, kw in (("plain", False, {}), ("realign", True, {}),
                             ("opts", False, {"min_depth": 5, "trim_ends": True, "uppercase": True})):
        res = k.bam_to_consensus(path, realign, kw.get("min_depth", 1), 7, 0.1, 50, kw.get("trim_ends", False),
                                 kw.get("uppercase", False))
        runs[tag] = {
            "fasta": [[r.name, r.sequence] for r in res.consensuses],
            "changes": {name: ["-" if c is None else c for c in ch] for name, ch in res.refs_changes.items()},
            "reports": res.refs_reports,
        }
        for name, ch in runs[tag]["changes"].items():
            runs[tag]["changes"][name] = "".join(ch)
    manifest["runs"] = runs
    arrays.update(frame_arrays(k.weights(path), "w_"))
    arrays.update(frame_arrays(k.weights(path, True, True, 0.05), "wrel_"))
    try:
        arrays.update(frame_arrays(k.features(path), "f_"))
        manifest["features_error"] = None
    except Exception as exc:  # reference bug on multi-contig input
        manifest["features_error"] = type(exc).__name__
    return arrays, manifest


def digest_for_file(k, path, rel_input, source):
    import hashlib

    alns = k.parse_bam(path)
    entry = {"input": rel_input, "source": source, "contigs": []}
    for name, aln in alns.items():
        t, ins = table_of(aln)
        t = np.ascontiguousarray(t, dtype=np.int32)
        seq, changes = k.consensus_sequence(aln.weights, aln.insertions, aln.deletions, None, False, 1, False)
        entry["contigs"].append({
            "name": name, "ref_len": len(aln.weights),
            "table_sha256": hashlib.sha256(t.tobytes()).hexdigest(),
            "column_sums": [int(x) for x in t.sum(axis=1)],
            "insertions": [[i, list(d.items())] for i, d in enumerate(ins) if d],
            "fasta_sha256": hashlib.sha256(seq.encode()).hexdigest(), "fasta_len": len(seq),
            "changes_sha256": hashlib.sha256("".join("-" if c is None else c for c in changes).encode()).hexdigest(),
        })
    return entry


# ---- synthetic edge cases, as SAM text over a 20 bp contig ---------------
def edge_cases():
    L = 20
    cases = []

    def case(name, reads, ref_len=L):
        cases.append({"name": name, "ref_len": ref_len, "reads": reads})

    case("A1_worked_example", [(5, "3S4M2I3M2D2M3S", "acgTTTTGGAAACCttt")])
    case("A3_refskip_is_noop", [(1, "3M5N3M", "ACGTAC")])
    case("A4_hard_clip_pad", [(3, "2H4M1P2M2H", "ACGTAC")])
    case("A5_H_then_S_is_right_clip", [(5, "2H3S4M", "GGGACGT")])
    case("A6_mid_cigar_S", [(1, "2M2S2M", "ACGTAC")])
    case("A7_left_clip_overhang", [(2, "5S3M", "ACGTAACG")])
    case("A8_right_clip_overhang", [(17, "3M5S", "ACGTTGCA")])
    case("A9_clip_starts_wraps", [(1, "2I3S", "ACGTA")])
    case("A9_pos_zero_wraps", [(0, "4M", "ACGT")])
    case("A9_pos_zero_left_clip", [(0, "2S3M", "ACGTA")])
    case("A9_pos_zero_deletion", [(0, "2D3M", "ACG")])
    case("A9_pos_zero_insertion", [(0, "2I3M", "ACGTA")])
    case("A10_M_past_end", [(18, "5M", "ACGTA")])
    case("A10_D_past_end", [(18, "2M3D1M", "ACG")])
    case("A10_D_to_last_slot", [(18, "2M1D", "ACG")])
    case("A10_iupac_in_M", [(3, "4M", "ACRT")])
    case("A10_iupac_in_I_is_fine", [(3, "2M2I2M", "ACRYGT")])
    case("A10_iupac_in_left_clip", [(6, "3S3M", "AYGACG")])
    case("A10_iupac_in_left_clip_overhang_ignored", [(1, "3S3M", "YYYACG")])
    case("A10_iupac_in_right_clip", [(3, "3M3S", "ACGAYG")])
    case("A10_iupac_in_right_clip_past_end_ignored", [(18, "3M3S", "ACGAYG"[:3] + "YYY")])
    case("A10_equals_base", [(3, "3M", "A=G")])
    case("A10_seq_shorter_than_cigar", [(3, "6M", "ACGT")])
    case("A10_seq_short_in_right_clip", [(3, "3M4S", "ACGTA")])
    case("A10_seq_short_right_clip_stalled", [(17, "3M4S", "ACGT")])
    case("A10_seq_short_left_clip", [(8, "4S2M", "ACG")])
    case("A10_insertion_slice_past_seq", [(3, "2M5I", "ACGT")])
    case("A10_start_beyond_contig_I", [(25, "2I", "AC")])
    case("A10_insertion_at_contig_end", [(19, "2M2I", "ACGT")])
    case("A10_clip_ends_beyond", [(25, "2S", "AC")])
    case("A11_flags_count", [(1, "4M", "ACGT", 256), (1, "4M", "ACGT", 2048), (1, "4M", "ACGT", 1024),
                             (1, "4M", "ACGT", 512), (1, "4M", "ACGT", 4), (1, "4M", "A", 0), (1, "*", "*", 0)])
    case("eq_and_X_ops", [(2, "2=1X2M", "ACGTN")])
    case("two_insertions_same_slot", [(4, "2M2I3N1I2M", "ACGGTAC"), (4, "2M1I2M", "ACTAC"), (4, "2M2I2M", "ACGGAC")])
    case("first_error_wins", [(1, "4M", "ACGT"), (18, "5M", "ACGTA"), (3, "4M", "ACRT")])
    case("key_error_before_index_error_in_one_read", [(17, "4M", "ARGT")])
    case("index_error_before_key_error_in_one_read", [(18, "4M", "ACGR")])
    case("lowercase_and_N", [(2, "6M", "acgtnN")])
    case("tiny_contig", [(1, "1M1S", "AC")], ref_len=1)
    case("empty_after_filter", [(1, "4M", "ACGT", 4)])
    # a busier mixed pileup on a 60 bp contig so the vote sees D / N / I / tie branches together
    mixed = []
    for s in range(1, 40, 3):
        mixed.append((s, "10M", "ACGTACGTAC"))
    mixed += [(5, "3M2D5M", "ACGTTGCA")] * 9 + [(12, "4M3I4M", "ACGTGGGACGT")] * 8 + [(12, "4M3I4M", "ACGTCCCACGT")] * 8
    mixed += [(30, "5S6M4S", "TTTTTACGTACGGGG")] * 5 + [(44, "6M", "NNNNNN")] * 2
    case("mixed_vote_branches", mixed, ref_len=60)
    return cases


def sam_text(case):
    lines = ["@HD\tVN:1.6", "@SQ\tSN:ctg\tLN:%d" % case["ref_len"]]
    for k, r in enumerate(case["reads"]):
        pos, cig, seq = r[0], r[1], r[2]
        flag = r[3] if len(r) > 3 else 0
        lines.append("r%d\t%d\tctg\t%d\t60\t%s\t*\t0\t0\t%s\t*" % (k, flag, pos, cig, seq))
    return "\n".join(lines) + "\n"


def golden_for_edges(k, tmpdir):
    out = []
    for case in edge_cases():
        path = os.path.join(tmpdir, case["name"] + ".sam")
        with open(path, "wt") as fh:
            fh.write(sam_text(case))
        entry = {"name": case["name"], "sam": sam_text(case)}
        try:
            alns = k.parse_bam(path)
        except Exception as exc:
            entry["raises"] = [type(exc).__name__, [str(a) for a in exc.args] if isinstance(exc, KeyError) else []]
            out.append(entry)
            continue
        entry["raises"] = None
        entry["contigs"] = list(alns.keys())
        if alns:
            aln = alns["ctg"]
            t, ins = table_of(aln)
            entry["counts"] = t.tolist()
            entry["insertions"] = [[i, list(d.items())] for i, d in enumerate(ins) if d]
            for md in (1, 3):
                res = k.bam_to_consensus(path, False, md, 7, 0.1, 50, False, False)
                entry["fasta_min_depth_%d" % md] = [[r.name, r.sequence] for r in res.consensuses]
                entry["changes_min_depth_%d" % md] = "".join("-" if c is None else c for c in res.refs_changes["ctg"])
        out.append(entry)
    return out


def main():
    import tempfile

    if not refload.available():
        raise SystemExit("the reference tree is not present; goldens can only be generated in the build container")
    k = refload.load_reference()
    os.makedirs(INPUTS, exist_ok=True)
    manifest = {"reference": "bede/kindel v1.2.1 @ 14d727b", "files": {}}
    for name, rel in FIXTURES:
        src = os.path.join(refload.REFERENCE_ROOT, "tests", rel)
        dst = reencode(src, os.path.join(INPUTS, name))
        # the re-encoded file must mean the same to the reference as the original
        a = k.bam_to_consensus(src, False, 1, 7, 0.1, 50, False, False)
        b = k.bam_to_consensus(dst, False, 1, 7, 0.1, 50, False, False)
        assert [(r.name, r.sequence) for r in a.consensuses] == [(r.name, r.sequence) for r in b.consensuses], name
        arrays, entry = golden_for_file(k, dst)
        entry["input"] = os.path.relpath(dst, GOLDEN)
        entry["source"] = "tests/" + rel
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **arrays)
        manifest["files"][name] = entry
        print("golden", name, os.path.getsize(dst), os.path.getsize(os.path.join(GOLDEN, name + ".npz")))
    manifest["digests"] = {}
    for name, rel in DIGEST_FIXTURES:
        src = os.path.join(refload.REFERENCE_ROOT, "tests", rel)
        dst = reencode(src, os.path.join(INPUTS, name))
        manifest["digests"][name] = digest_for_file(k, dst, os.path.relpath(dst, GOLDEN), "tests/" + rel)
        print("digest", name, os.path.getsize(dst))
    with tempfile.TemporaryDirectory() as tmp:
        manifest["edge_cases"] = golden_for_edges(k, tmp)
    with open(os.path.join(GOLDEN, "manifest.json"), "wt") as fh:
        json.dump(manifest, fh, indent=0, sort_keys=True)
    print("edge cases:", len(manifest["edge_cases"]),
          "raising:", sum(1 for e in manifest["edge_cases"] if e["raises"]))


if __name__ == "__main__":
    main()
