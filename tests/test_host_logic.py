"""Host-side halves of the API (string assembly, --realign reassembly, report, weights/features float
tails, CLI) checked against the reference's golden outputs.  The count tables fed in here come from
the CPU oracle (test infrastructure), so this runs without a GPU; the `-m gpu` tests feed the same
host code from the CUDA engine."""
import numpy as np
import pandas as pd
import pytest

from conftest import golden_input
from kindel_b200 import bamio, cli
from kindel_b200 import kindel as K
from oracle import coracle


def oracle_run(path):
    batch = bamio.read_alignment(path)
    counts, events = coracle.pileup(batch)
    return K.PileupRun.from_host_tables(batch, counts, coracle.derive(counts), events), counts


def frame_from_golden(g, prefix):
    cols = [str(c) for c in g[prefix + "__columns"]]
    return pd.DataFrame({c: g[prefix + c] for c in cols}, columns=cols)


def assert_frame_matches(df, g, prefix):
    want = frame_from_golden(g, prefix)
    assert list(df.columns) == list(want.columns)
    assert len(df) == len(want)
    for c in want.columns:
        a, b = df[c].to_numpy(), want[c].to_numpy()
        if b.dtype.kind in "US":
            assert a.astype(str).tolist() == b.astype(str).tolist(), c
        else:
            assert a.dtype == b.dtype, (c, a.dtype, b.dtype)
            np.testing.assert_array_equal(a, b, err_msg=prefix + c)  # NaN == NaN here; bit-exact otherwise


def test_unit_consensus():  # reference tests/test_kindel.py:25-32
    w = {"A": 1, "C": 2, "G": 3, "T": 4, "N": 5}
    assert K.consensus(w) == ("N", 5, 0.33, False)
    assert K.consensus({"A": 5, "C": 5, "G": 3, "T": 4, "N": 1})[3] is True
    assert K.consensus({"A": 0, "T": 0, "G": 0, "C": 0, "N": 0}) == ("N", 0, 0, False)
    assert K.consensus({"A": 2, "T": 2, "G": 0, "C": 0, "N": 0}) == ("A", 2, 0.5, True)
    assert K.consensus({"A": 0, "T": 0, "G": 3, "C": 3, "N": 0})[0] == "G"
    assert K.consensus({}) == ("N", 0, 0, False)


def test_unit_merge_by_lcs():  # reference tests/test_kindel.py:35-53
    one = ("AACTGCCGCTAGGGGCGCGTTCGGGCTCGCCAACATCTTCAGTCCGGG",
           "GCCGCTAGGGGCGCGTTCGGGCTCGCCAACATCTTCAGTCCGGGCGCTAAGCAGAACA")
    two = ("AACTGCCGCTAGGGGCGCGTTCGGGCTCGCCAACATCTTCAGTCCGGGCGCTAAGCAGAACATC",
           "GCAGATACCTACACCACCGGGGGAACTGCCGCTAGGGGCGCGTTCGGGCTCGCCAACATCTTCAGTCCGGGCGCTAAGCAGAACA")
    want = "AACTGCCGCTAGGGGCGCGTTCGGGCTCGCCAACATCTTCAGTCCGGGCGCTAAGCAGAACA"
    assert K.merge_by_lcs(*one, min_overlap=7) == want
    assert K.merge_by_lcs(*two, min_overlap=7) == want
    assert K.merge_by_lcs("AT", "CG", min_overlap=7) is None


def test_cdrp_consensuses_known_strings(manifest):  # reference tests/test_kindel.py:92-111
    run, _ = oracle_run(golden_input(manifest["files"]["bwa_1_1"]))
    aln = run.alignment(0)
    cdrps = K.cdrp_consensuses(aln.weights, aln.deletions, aln.clip_start_weights, aln.clip_end_weights,
                               aln.clip_start_depth, aln.clip_end_depth, 0.1, 10)
    assert cdrps[0][0].seq == "AACTGCCGCTAGGGGCGCGTTCGGGCTCGCCAACATCTTCAGTCCGGGCGCTAAGCAGAACATCCAGCTGATCAACA"
    assert cdrps[0][1].seq == ("AGCGTCGATGCAGATACCTACACCACCGGGGGAACTGCCGCTAGGGGCGCGTTCGGGCTCGCCAACATCTTCAGTCCGGG"
                               "CGCTAAGCAGAACA")


def test_known_answer_integers_through_the_views(manifest):  # reference tests/test_kindel.py:63-89
    a = oracle_run(golden_input(manifest["files"]["bwa_1_1"]))[0].alignment(0)
    b = oracle_run(golden_input(manifest["files"]["ext_3_bc75"]))[0].alignment(0)
    assert a.ref_id == "ENA|EU155341|EU155341.2" and len(a.weights) == 9306
    assert a.weights[0]["A"] == 22 and a.weights[23]["A"] == 57
    assert b.weights[68]["G"] == 1 and b.weights[2368]["T"] == 13
    assert [b.deletions[i] for i in (399, 402, 411, 1048, 1049, 1050)] == [14, 14, 15, 14, 14, 14]
    assert b.clip_ends[1748] == 12 and a.clip_starts[525] == 16 and a.clip_starts[1437] == 84
    assert sum(b.insertions[452 + 1].values()) == 14 and sum(b.insertions[456 + 1].values()) == 14
    # namedtuple behaviour of the reference's `alignment`
    ref_id, weights, insertions, deletions, *rest = a
    assert ref_id == a.ref_id and len(rest) == 8 and a[1] is a.weights
    assert list(a.weights[0].keys()) == ["A", "T", "G", "C", "N"]
    assert len(a.weights[9300:]) == 6 and len(a.insertions) == 9307 and len(a.deletions) == 9307


def test_consensus_report_changes_against_golden(manifest):
    for name, entry in manifest["files"].items():
        path = golden_input(entry)
        run, counts = oracle_run(path)
        for tag, realign, md, trim, upper in (("plain", False, 1, False, False), ("realign", True, 1, False, False),
                                              ("opts", False, 5, True, True)):
            res = K.consensus_from_run(run, coracle.vote(counts, md), "X", realign, md, 7, 0.1, 50, trim, upper)
            want = entry["runs"][tag]
            assert [[r.name, r.sequence] for r in res.consensuses] == want["fasta"], (name, tag)
            for ctg, ch in res.refs_changes.items():
                assert "".join("-" if c is None else c for c in ch) == want["changes"][ctg]
            for ctg, rep in res.refs_reports.items():
                # the report echoes the input path; everything else must be identical
                exp = want["reports"][ctg].splitlines()
                got = rep.splitlines()
                assert len(exp) == len(got)
                assert [l for l in got if not l.startswith("- bam_path")] == [l for l in exp if not l.startswith("- bam_path")]


def check_clip_case(case, got, tables):
    """One case of tests/golden/clip_cases.json against a `result` and {contig: int32 [19, L+1] table}."""
    import hashlib

    seed = case["seed"]
    assert list(tables) == case["contigs"], seed
    for ctg, t in tables.items():
        sha = hashlib.sha256(np.ascontiguousarray(t, dtype=np.int32).tobytes()).hexdigest()
        assert sha == case["table_sha256"][ctg], (seed, ctg)
    assert [[r.name, r.sequence] for r in got.consensuses] == case["fasta"], (seed, case["options"])
    for ctg, ch in got.refs_changes.items():
        assert "".join("-" if c is None else c for c in ch) == case["changes"][ctg], (seed, ctg)
    for ctg, rep in got.refs_reports.items():
        assert [l for l in rep.splitlines() if not l.startswith("- bam_path")] == case["reports"][ctg], (seed, ctg)


def test_clip_heavy_cases_against_golden(clip_golden, tmp_path):
    """--realign machinery on synthetic clip-dominant regions (tests/clip_cases.py): oracle tables + host
    code == what the unmodified reference returned (oracle/make_clip_golden.py)."""
    from clip_cases import clip_case

    import helpers as H

    assert len(clip_golden["cases"]) >= 90
    for case in clip_golden["cases"]:
        path = tmp_path / ("clip%d.sam" % case["seed"])
        path.write_text(clip_case(case["seed"]))
        run, counts = oracle_run(path)
        o = case["options"]
        got = K.consensus_from_run(run, coracle.vote(counts, o[1]), str(path), *o)
        tables = {name: H.contig_view(run.batch, counts, c) for c, name in enumerate(run.batch.contig_names)}
        check_clip_case(case, got, tables)


def test_weights_and_features_frames_against_golden(manifest, golden_npz):
    for name, entry in manifest["files"].items():
        run, _ = oracle_run(golden_input(entry))
        g = golden_npz(name)
        assert_frame_matches(K.weights_from_run(run), g, "w_")
        assert_frame_matches(K.weights_from_run(run, True, True, 0.05), g, "wrel_")
        if entry["features_error"]:
            with pytest.raises(IndexError):
                K.features_from_run(run)
        else:
            assert_frame_matches(K.features_from_run(run), g, "f_")


def test_assemble_with_cdr_patches_and_skips():
    calls = np.array([0, 1, 2, 3, 0x14, 0x24, 0x31, 0, 1, 2], dtype=np.uint8)  # A C G T D N I+C A C G
    look = lambda p: ("GG", False)
    seq, ch = K.assemble_consensus(calls, look)
    assert seq == "ACGTNggCACG" and ch == [None, None, None, None, "D", "N", "I", None, None, None]
    R = K.Region
    seq, ch = K.assemble_consensus(calls, look, [R(2, 5, "TTTT", None)])
    assert seq == "ACttttNggCACG" and ch[2:5] == [None, None, None]
    seq, _ = K.assemble_consensus(calls, look, [R(2, 5, None, None)])  # falsy seq: ignored
    assert seq == "ACGTNggCACG"
    seq, _ = K.assemble_consensus(calls, look, [R(2, 2, "TT", None)])  # zero span: reference never recovers
    assert seq == "ACtt"
    seq, _ = K.assemble_consensus(np.array([4, 4, 0, 4], dtype=np.uint8), look, None, True, False)
    assert seq == "A"
    seq, _ = K.assemble_consensus(calls, lambda p: ("GG", True), None, False, True)
    assert seq == "ACGTNNCACG"


def test_cli_flag_surface():
    p = cli.build_parser()
    a = p.parse_args(["consensus", "x.bam"])
    assert (a.realign, a.min_depth, a.min_overlap, a.clip_decay_threshold, a.mask_ends, a.trim_ends, a.uppercase) == (
        False, 1, 7, 0.1, 50, False, False)
    a = p.parse_args(["consensus", "-r", "--min-depth", "3", "--min-overlap", "9", "-c", "0.2", "--mask-ends", "10",
                      "-t", "-u", "x.bam"])
    assert (a.realign, a.min_depth, a.min_overlap, a.clip_decay_threshold, a.mask_ends, a.trim_ends, a.uppercase) == (
        True, 3, 9, 0.2, 10, True, True)
    a = p.parse_args(["weights", "-r", "--confidence-alpha", "0.05", "x.bam"])
    assert a.relative and a.confidence and a.confidence_alpha == 0.05
    assert cli.version() == "kindel 1.2.1"
    import kindel
    assert kindel.__version__ == "1.2.1" and kindel.kindel.bam_to_consensus is K.bam_to_consensus


def test_bam_writer_reader_roundtrip(tmp_path):
    from kindel_b200 import synth

    batch = synth.complex_reads(11, 5000, 30)
    contigs, recs = synth.to_records(batch)
    recs.insert(3, (-1, -1, 4, [], "ACGT"))         # unmapped, rname *
    recs.insert(5, (0, 10, 4, [], "ACGT"))          # unmapped but placed: filtered
    recs.insert(7, (0, 10, 256, [(4 << 4)], "*"))   # SEQ * : filtered
    path = tmp_path / "rt.bam"
    bamio.write_bam(path, contigs, recs)
    back = bamio.read_alignment(path)
    assert back.n_records == len(recs) and back.n_reads == batch.n_reads
    for f in ("ref_start", "cig_off", "cigar", "contig_len", "contig_read_off", "contig_slot", "complex_idx", "hard_idx", "l_seq", "seq_len", "seq_off", "seq4"):
        np.testing.assert_array_equal(getattr(back, f), getattr(batch, f), err_msg=f)
    c0, e0 = coracle.pileup(batch)
    c1, e1 = coracle.pileup(back)
    np.testing.assert_array_equal(c0, c1)
    np.testing.assert_array_equal(e0, e1)


def test_bam_gather_flags_exotic_bases(tmp_path):
    """The C++ gather marks reads holding a base outside A,C,G,T,N; they become complex reads, so the
    fast kernel never sees them and the general walk reproduces the reference's KeyError semantics."""
    recs = [(0, 3, 0, [6 << 4], "ACGTNA"), (0, 4, 0, [6 << 4], "ACRTAC"), (0, 5, 0, [5 << 4], "ACG=A"),
            (0, 6, 0, [9 << 4], "ACGTACGTN"), (0, 7, 0, [(2 << 4), (2 << 4) | 1, (2 << 4)], "ACYRGT")]
    p = tmp_path / "x.bam"
    bamio.write_bam(p, [("c", 40)], recs)
    b = bamio.read_alignment(p)
    assert b.n_reads == 5
    assert (b.l_seq < 0).tolist() == [False, True, True, False, True]   # R, '=' -> complex; indel read too
    with pytest.raises(KeyError) as exc:
        coracle.pileup(b)
    assert exc.value.args == ("R",)
    # the same reads through SAM text classify identically (numpy check instead of the C++ one)
    sam = tmp_path / "x.sam"
    sam.write_text("@SQ\tSN:c\tLN:40\n" + "".join(
        "r%d\t0\tc\t%d\t60\t%s\t*\t0\t0\t%s\t*\n" % (k, r[1] + 1, "".join("%d%s" % (w >> 4, "MIDNSHP=X"[w & 15]) for w in r[3]), r[4])
        for k, r in enumerate(recs)))
    b2 = bamio.read_alignment(sam)
    np.testing.assert_array_equal(b2.l_seq, b.l_seq)
    np.testing.assert_array_equal(b2.seq4, b.seq4)


def test_variants_extension(tmp_path):
    """`variants` (extension, no reference code to pin against): thresholds, consensus exclusion, deletions."""
    sam = ["@HD\tVN:1.6", "@SQ\tSN:v\tLN:8"]
    reads = ["ACGTACGT"] * 6 + ["ACGAACGT"] * 3 + ["ACGTACTT"] * 1     # pos 4: T6 A3; pos 7: G9 T1
    for k, seq in enumerate(reads):
        sam.append("r%d\t0\tv\t1\t60\t8M\t*\t0\t0\t%s\t*" % (k, seq))
    sam.append("d0\t0\tv\t1\t60\t2M2D4M\t*\t0\t0\tACACGT\t*")              # deletion over pos 3-4
    sam.append("d1\t0\tv\t1\t60\t2M2D4M\t*\t0\t0\tACACGT\t*")
    path = tmp_path / "v.sam"
    path.write_text("\n".join(sam) + "\n")
    run, _ = oracle_run(path)
    df = K.variants_from_run(run)
    assert list(df.columns) == ["chrom", "pos", "depth", "consensus", "A", "C", "G", "T", "N", "deletions"]
    assert df["depth"].tolist() == [12] * 8 and "".join(df["consensus"]) == "ACGTACGT"
    row4, row7, row3 = df.iloc[3], df.iloc[6], df.iloc[2]
    assert row4["A"] == round(3 / 12, 4) and row4["deletions"] == round(2 / 12, 4) and row4["T"] == 0
    assert row7["T"] == 0                      # a single read does not exceed abs_threshold = 1
    assert row3["deletions"] == round(2 / 12, 4) and row3["G"] == 0
    only = K.variants_from_run(run, only_variants=True, absolute=True)
    assert only["pos"].tolist() == [3, 4] and only.iloc[1]["A"] == 3 and only.iloc[1]["deletions"] == 2
    assert K.variants_from_run(run, abs_threshold=0, rel_threshold=0.05, only_variants=True)["pos"].tolist() == [3, 4, 7]
    assert K.variants_from_run(run, rel_threshold=0.2, only_variants=True)["pos"].tolist() == [4]
    a = cli.build_parser().parse_args(["variants", "-a", "2", "-r", "0.1", "-o", "--absolute", "x.bam"])
    assert (a.abs_threshold, a.rel_threshold, a.only_variants, a.absolute) == (2, 0.1, True, True)


def test_bam_with_a_long_reference_dictionary(tmp_path):
    """A draft-assembly / metagenome style header: tens of thousands of contigs, several MB of binary dictionary
    (read_bam used to parse it from a truncated copy of the buffer's front)."""
    from kindel_b200 import bamio

    n = 60_000
    contigs = [("contig_%06d_with_a_rather_long_descriptive_name_%s" % (k, "x" * (k % 37)), 1000 + k % 50) for k in range(n)]
    recs = [(k, 5, 0, [(20 << 4)], "ACGTACGTACGTACGTACGT") for k in (0, 17, n // 2, n - 1)]
    path = tmp_path / "many.bam"
    bamio.write_bam(path, contigs, recs, level=1)
    b = bamio.read_alignment(path)
    assert b.n_reads == 4 and b.contig_names == [contigs[k][0] for k in (0, 17, n // 2, n - 1)]
    assert list(b.contig_len) == [contigs[k][1] for k in (0, 17, n // 2, n - 1)]


def test_bam_decoder_finds_records_across_task_ranges(tmp_path):
    """The C++ decoder walks byte ranges of the inflated stream in parallel from GUESSED record boundaries and then
    verifies the chain: records far longer than a range (nothing starts inside some ranges), ranges that begin inside
    quality strings of 0xff bytes, unmapped and filtered records, several contigs -- the result must not depend on
    the number of threads and must equal an independent sequential decode (oracle/samdecode.py)."""
    from oracle import samdecode

    rng = np.random.default_rng(5)
    contigs = [("a", 900_000), ("b", 5_000), ("c", 70_000)]
    recs = []
    def rnd(n):
        return "".join("ACGTN"[i] for i in rng.integers(0, 5, size=n))
    for k in range(6000):
        ref = int(rng.integers(0, 3))
        L = contigs[ref][1]
        n = int(rng.integers(2, 200))
        pos = int(rng.integers(0, L - n))
        kind = k % 7
        if kind == 0:
            recs.append((-1, -1, 4, [], rnd(n)))                                   # unmapped, rname *
        elif kind == 1 and n > 4:
            recs.append((ref, pos, 0, [(3 << 4) | 4, ((n - 3) << 4)], rnd(n)))     # 3S(n-3)M
        elif kind == 2 and n > 3:
            recs.append((ref, pos, 0, [(1 << 4), (1 << 4) | 1, ((n - 2) << 4)], rnd(n)))
        else:
            recs.append((ref, pos, 0, [(n << 4)], rnd(n)))
    for at, n in ((100, 420_000), (101, 300_000), (3000, 650_000)):                 # records of 0.45 .. 1 MB
        recs.insert(at, (0, 1000, 0, [(n << 4)], rnd(n)))
    path = tmp_path / "ranges.bam"
    bamio.write_bam(path, contigs, recs, level=1)
    ref_recs = [r for r in samdecode.read_alignment_file(str(path))[1]]
    base = None
    for threads in (1, 2, 3, 8, 32):
        b = bamio.read_bam(path, threads=threads)
        assert b.n_records == len(recs)
        if base is None:
            base = b
            kept = [r for r in ref_recs if r.mapped and len(r.seq) > 1 and r.rname != "*"]
            assert b.n_reads == len(kept)
            # same reads in the reference's iteration order: grouped by contig in first-seen order, file order inside
            order = []
            for r in ref_recs:
                if r.rname != "*" and r.rname not in order:
                    order.append(r.rname)
            assert b.contig_names == order
            want = [r for name in order for r in kept if r.rname == name]
            np.testing.assert_array_equal(b.ref_start, np.array([r.pos - 1 for r in want], dtype=np.int32))
            np.testing.assert_array_equal(b.seq_len, np.array([len(r.seq) for r in want], dtype=np.int32))
        else:
            for f in ("ref_start", "cig_off", "cigar", "contig_len", "contig_read_off", "complex_idx", "hard_idx", "l_seq",
                      "seq_len", "seq_off", "seq4"):
                np.testing.assert_array_equal(getattr(b, f), getattr(base, f), err_msg="%s, %d threads" % (f, threads))
    # a truncated stream is refused, whatever range the cut falls into
    raw = bamio.inflate_bam(path)
    cut = tmp_path / "cut.bam"
    with open(cut, "wb") as fh:
        fh.write(bamio._bgzf_block(bytes(raw[: len(raw) - 37]), 1) if len(raw) < 60000 else b"".join(
            bamio._bgzf_block(bytes(raw[s:min(s + 60000, len(raw) - 37)]), 1) for s in range(0, len(raw) - 37, 60000)))
        fh.write(bamio._bgzf_block(b"", 1))
    with pytest.raises(ValueError):
        bamio.read_bam(cut, threads=8)


def test_long_cigar_in_the_cg_tag(tmp_path):
    """A read whose real CIGAR sits in the CG:B,I tag behind the `<l_seq>S<span>N` placeholder (SAM spec 4.2.2; what
    samtools -- and so the reference's reader -- expands) decodes like the same read with the CIGAR in place; other
    tags before it (all aux types) are skipped correctly, and a placeholder without the tag stays what it says."""
    import struct

    def record(ref_id, pos0, cig, seq, aux=b""):
        qname = b"q\x00"
        packed = bamio.words_to_bam_bytes(bamio.encode_seq(seq), len(seq))
        core = struct.pack("<iiBBHHHiiii", ref_id, pos0, len(qname), 60, 4680, len(cig), 0, len(seq), -1, -1, 0)
        data = core + qname + struct.pack("<%dI" % len(cig), *cig) + packed + b"\xff" * len(seq) + aux
        return struct.pack("<i", len(data)) + data

    def bam(path, records):
        ht = b"@SQ\tSN:c\tLN:400\n"
        body = b"BAM\x01" + struct.pack("<i", len(ht)) + ht + struct.pack("<i", 1) + struct.pack("<i", 2) + b"c\x00" + struct.pack("<i", 400)
        body += b"".join(records)
        with open(path, "wb") as fh:
            fh.write(bamio._bgzf_block(body, 1) + bamio._bgzf_block(b"", 1))

    rng = np.random.default_rng(2)
    seq = "".join("ACGT"[i] for i in rng.integers(0, 4, size=60))
    real = [(5 << 4) | 4, (20 << 4), (3 << 4) | 1, (10 << 4), (4 << 4) | 2, (22 << 4)]       # 5S20M3I10M4D22M
    span = 20 + 10 + 4 + 22
    placeholder = [(60 << 4) | 4, (span << 4) | 3]                                              # 60S56N
    aux_before = (b"NMi" + struct.pack("<i", 3) + b"XAA" + b"x" + b"XcC" + b"\x07" + b"XsS" + struct.pack("<H", 9) +
                  b"MDZ" + b"20A35\x00" + b"XhH" + b"1AE3\x00" + b"XbBc" + struct.pack("<I", 3) + b"\x01\x02\x03" +
                  b"XfBf" + struct.pack("<I", 2) + struct.pack("<ff", 1.0, 2.0) + b"Xff" + struct.pack("<f", 0.5))
    cg = b"CGBI" + struct.pack("<I", len(real)) + struct.pack("<%dI" % len(real), *real)
    plain = [record(0, 100, [(60 << 4)], seq)]
    a = tmp_path / "inline.bam"
    bam(a, plain + [record(0, 30, real, seq)])
    b = tmp_path / "tagged.bam"
    bam(b, plain + [record(0, 30, placeholder, seq, aux_before + cg + b"ZZi" + struct.pack("<i", 1))])
    c = tmp_path / "untagged.bam"
    bam(c, plain + [record(0, 30, placeholder, seq, aux_before)])
    ba, bb, bc = bamio.read_bam(a), bamio.read_bam(b), bamio.read_bam(c)
    for f in ("ref_start", "seq_len", "l_seq", "seq_off", "seq4", "cig_off", "cigar", "complex_idx", "hard_idx"):
        np.testing.assert_array_equal(getattr(bb, f), getattr(ba, f), err_msg=f)
    assert bb.cigar[bb.cig_off[1]:bb.cig_off[2]].tolist() == real
    assert bc.cigar[bc.cig_off[1]:bc.cig_off[2]].tolist() == placeholder       # no tag: the placeholder is the CIGAR
    ca, _ = coracle.pileup(ba)
    cb, _ = coracle.pileup(bb)
    np.testing.assert_array_equal(ca, cb)
