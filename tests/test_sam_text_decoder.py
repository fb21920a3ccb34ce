"""SAM text through the C++ decoder (bam_host.cpp: lines -> BAM records in threads, shared classification / layout)
equals the Python text reader (bamio.read_sam), and whatever the strict C++ parser refuses falls back to it --
same arrays or the same exception, case by case."""
import numpy as np
import pytest

from fuzz_cases import random_case
from kindel_b200 import bamio

FIELDS = ("contig_len", "contig_read_off", "contig_slot", "ref_start", "seq_len", "l_seq", "seq_off", "seq4", "cig_off",
          "cigar", "complex_idx", "hard_idx")


def same(a, b, tag):
    assert a.contig_names == b.contig_names and a.n_records == b.n_records and a.n_slots == b.n_slots, tag
    for f in FIELDS:
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg="%s %s" % (tag, f))
    assert (a.n_events, a.reads_sorted, a.max_simple_len, a.reach_right, a.reach_left, a.aligned_bases) == \
           (b.n_events, b.reads_sorted, b.max_simple_len, b.reach_right, b.reach_left, b.aligned_bases), tag


def outcome(fn, path):
    try:
        return fn(path), None
    except Exception as exc:  # noqa: BLE001 -- the point is to compare whatever is raised
        return None, (type(exc), exc.args)


HDR = "@HD\tVN:1.6\n@SQ\tSN:a\tLN:50\n@SQ\tSN:b\tLN:30\n"
REC = "r\t0\ta\t5\t60\t8M\t*\t0\t0\tACGTACGT\t*\n"
CASES = {
    "plain": HDR + REC + "r\t16\tb\t1\t60\t2S4M1I1M\t*\t0\t0\tacgtnACG\t*\n",
    "crlf": (HDR + REC + REC).replace("\n", "\r\n"),
    "no_trailing_newline": (HDR + REC + REC).rstrip("\n"),
    "unmapped_and_star": HDR + "r\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t*\n" + "r\t4\ta\t7\t0\t*\t*\t0\t0\tAC?T\t*\n" + REC,
    "seq_star_and_short": HDR + "r\t0\ta\t3\t60\t*\t*\t0\t0\t*\t*\n" + "r\t0\tb\t3\t60\t1M\t*\t0\t0\tA\t*\n" + REC,
    "cigar_star": HDR + "r\t0\ta\t3\t60\t*\t*\t0\t0\tACGT\t*\n",
    "pos_zero": HDR + "r\t0\ta\t0\t60\t4M\t*\t0\t0\tACGT\t*\n" + REC,
    "short_lines_skipped": HDR + "garbage line\n" + "a\tb\tc\n" + REC,
    "contig_seen_only_filtered": HDR + "r\t4\tb\t3\t60\t4M\t*\t0\t0\tACGT\t*\n" + REC,
    "unknown_op_letters": HDR + "r\t0\ta\t3\t60\t2M1Z2M3\t*\t0\t0\tACGTA\t*\n",
    "iupac_and_equals": HDR + "r\t0\ta\t3\t60\t6M\t*\t0\t0\tAC=RYN\t*\n",
    "many_fields": HDR + "r\t0\ta\t5\t60\t8M\t*\t0\t0\tACGTACGT\t*\tNM:i:0\tXX:Z:a\tb\n",
    "empty_body": HDR,
    "no_header": REC,
    # refused by the C++ parser: the Python reader decides
    "unknown_rname": HDR + "r\t0\tzzz\t5\t60\t8M\t*\t0\t0\tACGTACGT\t*\n",
    "unknown_rname_filtered_read": HDR + "r\t4\tzzz\t5\t60\t8M\t*\t0\t0\tACGTACGT\t*\n" + REC,
    "bad_base": HDR + "r\t0\ta\t5\t60\t8M\t*\t0\t0\tACGTAC?T\t*\n",
    "bad_base_in_unused_read": HDR + "r\t4\ta\t5\t60\t8M\t*\t0\t0\tAC??ACGT\t*\n" + REC,
    "spacey_integers": HDR + "r\t 0\ta\t 5 \t60\t8M\t*\t0\t0\tACGTACGT\t*\n",
    "non_integer_pos": HDR + "r\t0\ta\tfive\t60\t8M\t*\t0\t0\tACGTACGT\t*\n",
    "header_between_records": HDR + REC + "@CO\tlate\n" + REC,
    "sq_without_ln": "@SQ\tSN:a\n@SQ\tSN:b\tLN:30\n" + "r\t0\tb\t5\t60\t8M\t*\t0\t0\tACGTACGT\t*\n",
    "duplicate_sn": "@SQ\tSN:a\tLN:50\n@SQ\tSN:a\tLN:60\n" + REC,
    "huge_flag": HDR + "r\t70000\ta\t5\t60\t8M\t*\t0\t0\tACGTACGT\t*\n",
    "lone_cr_inside_a_line": HDR + "r1\t0\ta\t3\t60\t7M\t*\t0\rD\t0\tCAACCTC\t*\n" + REC,   # text mode breaks the line at the CR
    "non_ascii_qname": HDR + "r\u00e9\t0\ta\t5\t60\t8M\t*\t0\t0\tACGTACGT\t*\n",
}


@pytest.mark.parametrize("name", list(CASES))
def test_sam_text_cases(tmp_path, name):
    p = tmp_path / (name + ".sam")
    p.write_bytes(CASES[name].encode())
    want, werr = outcome(bamio.read_sam, p)
    got, gerr = outcome(bamio.read_alignment, p)
    if werr is not None:
        assert gerr == werr, (name, gerr, werr)
    else:
        assert gerr is None, (name, gerr)
        same(got, want, name)


def test_sam_text_equals_python_reader_on_fuzz_cases(tmp_path):
    for seed in range(0, 400, 3):
        p = tmp_path / ("f%d.sam" % seed)
        p.write_text(random_case(seed))
        want, werr = outcome(bamio.read_sam, p)
        if werr is None:
            for threads in (1, 4):
                same(bamio.read_bam(p, threads=threads), want, "seed %d, %d threads" % (seed, threads))  # no fallback needed
        got, gerr = outcome(bamio.read_alignment, p)
        assert gerr == werr, seed


def test_large_sam_text_in_threads(tmp_path):
    """Enough lines for several byte ranges per thread; also gzip-compressed text."""
    import gzip

    from kindel_b200 import synth

    batch = synth.mixed_reads(3, [60_000, 20_000], 40, 0.3)
    contigs, recs = synth.to_records(batch)
    ops = "MIDNSHP=X"
    lines = ["@HD\tVN:1.6"] + ["@SQ\tSN:%s\tLN:%d" % c for c in contigs]
    for k, (ref, pos, flag, cig, seq) in enumerate(recs):
        lines.append("r%d\t%d\t%s\t%d\t60\t%s\t*\t0\t0\t%s\t*" % (
            k, flag, contigs[ref][0], pos + 1, "".join("%d%s" % (w >> 4, ops[w & 15]) for w in cig) or "*", seq))
    text = "\n".join(lines) + "\n"
    assert len(text) > 3 << 20
    p = tmp_path / "big.sam"
    p.write_text(text)
    want = bamio.read_sam(p)
    for threads in (1, 3, 8):
        same(bamio.read_bam(p, threads=threads), want, "%d threads" % threads)
    same(want, batch, "synth")
    gz = tmp_path / "big.sam.gz"
    with gzip.open(gz, "wb", compresslevel=1) as fh:
        fh.write(text.encode())
    same(bamio.read_alignment(gz), want, "gzip")
