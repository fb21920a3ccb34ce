"""Pins the oracle: the C restatement (oracle/kindel_oracle.c) against
  (a) the golden vectors the UNMODIFIED reference produced (tests/golden/, oracle/make_golden.py), and
  (b) -- only where /root/reference exists (the build container) -- the reference itself on every
      fixture of its own test-suite, including the known-answer integers of
      reference tests/test_kindel.py:63-89.
No GPU involved."""
import os
from collections import OrderedDict

import numpy as np
import pytest

import helpers as H
from conftest import golden_input
from kindel_b200 import bamio
from oracle import coracle, refload, samdecode


def test_golden_tables_insertions_and_vote(manifest, golden_npz):
    for name, entry in manifest["files"].items():
        batch = bamio.read_alignment(golden_input(entry))
        counts, events = coracle.pileup(batch)
        derived = coracle.derive(counts)
        ins = H.events_to_dicts(batch, events)
        g = golden_npz(name)
        assert batch.contig_names == [c["name"] for c in entry["contigs"]]
        calls = coracle.vote(counts, 1)
        for c, meta in enumerate(entry["contigs"]):
            L = meta["ref_len"]
            s0 = int(batch.contig_slot[c])
            assert int(batch.contig_len[c]) == L
            np.testing.assert_array_equal(H.contig_view(batch, counts, c), g["c%d_counts" % c], err_msg=name)
            np.testing.assert_array_equal(derived[0, s0:s0 + L], g["c%d_consensus_depth" % c])
            np.testing.assert_array_equal(derived[1, s0:s0 + L], g["c%d_clip_start_depth" % c])
            np.testing.assert_array_equal(derived[2, s0:s0 + L], g["c%d_clip_end_depth" % c])
            np.testing.assert_array_equal(derived[3, s0:s0 + L], g["c%d_clip_depth" % c])
            want = {s0 + i: [tuple(kv) for kv in items] for i, items in meta["insertions"]}
            got = {s: list(d.items()) for s, d in ins.items() if s0 <= s <= s0 + L}
            assert got == want, name
            changes = "".join("-" if c_ is None else c_ for c_ in H.calls_to_changes(calls[s0:s0 + L]))
            assert changes == entry["runs"]["plain"]["changes"][meta["name"]], name


def _case_batch(case, tmp_path):
    p = tmp_path / (case["name"] + ".sam")
    p.write_text(case["sam"])
    return bamio.read_alignment(p), p


def test_edge_cases_match_the_reference(manifest, tmp_path):
    """SURVEY.md Appendix A behaviours, each pinned by the reference's own output."""
    assert len(manifest["edge_cases"]) >= 40
    for case in manifest["edge_cases"]:
        batch, _ = _case_batch(case, tmp_path)
        if case["raises"]:
            kind, args = case["raises"]
            with pytest.raises({"IndexError": IndexError, "KeyError": KeyError}[kind]) as exc:
                coracle.pileup(batch)
            if kind == "KeyError":
                assert [str(a) for a in exc.value.args] == args, case["name"]
            continue
        counts, events = coracle.pileup(batch)
        assert batch.contig_names == case["contigs"]
        np.testing.assert_array_equal(H.contig_view(batch, counts, 0), np.array(case["counts"]), err_msg=case["name"])
        got = {s: list(d.items()) for s, d in H.events_to_dicts(batch, events).items()}
        assert got == {i: [tuple(kv) for kv in items] for i, items in case["insertions"]}, case["name"]
        L = int(batch.contig_len[0])
        for md in (1, 3):
            calls = coracle.vote(counts, md)[:L]
            changes = "".join("-" if c is None else c for c in H.calls_to_changes(calls))
            assert changes == case["changes_min_depth_%d" % md], case["name"]


def test_oracle_decoder_and_product_decoder_agree(manifest):
    """Two independent BAM/SAM decoders (stdlib struct vs C++ gather) see the same records."""
    for name, entry in manifest["files"].items():
        path = golden_input(entry)
        _, records = samdecode.read_alignment_file(path)
        batch = bamio.read_alignment(path)
        kept = [r for r in records if r.mapped and len(r.seq) > 1 and r.rname != "*"]
        assert batch.n_records == len(records)
        assert batch.n_reads == len(kept)
        order = OrderedDict()
        for r in records:
            if r.rname != "*":
                order.setdefault(r.rname, []).append(r)
        assert list(order) == batch.contig_names
        flat = [r for rs in order.values() for r in rs if r.mapped and len(r.seq) > 1]
        for k in list(range(0, len(flat), max(1, len(flat) // 200))):
            r = flat[k]
            assert int(batch.ref_start[k]) == r.pos - 1
            words = batch.cigar[int(batch.cig_off[k]):int(batch.cig_off[k + 1])].tolist()
            assert [(w >> 4, "MIDNSHP=X"[w & 15]) for w in words] == [c for c in r.cigars if c[1] is not None]
            assert H.event_string(batch, k, 0, len(r.seq)) == r.seq.upper() or int(batch.l_seq[k]) >= 0


needs_reference = pytest.mark.skipif(not refload.available(), reason="reference tree only exists in the build container")


@needs_reference
def test_reference_known_answers_through_the_oracle_loader():
    """reference tests/test_kindel.py:63-89, executed against the unmodified module."""
    k = refload.load_reference()
    root = os.path.join(refload.REFERENCE_ROOT, "tests")
    aln = list(k.parse_bam(os.path.join(root, "data_bwa_mem", "1.1.sub_test.bam")).values())[0]
    aln2 = list(k.parse_bam(os.path.join(root, "data_ext", "3.issue23.bc75.sam")).values())[0]
    assert aln.ref_id == "ENA|EU155341|EU155341.2" and len(aln.weights) == 9306
    assert aln.weights[0]["A"] == 22 and aln.weights[23]["A"] == 57
    assert aln2.weights[68]["G"] == 1 and aln2.weights[2368]["T"] == 13
    assert [aln2.deletions[i] for i in (399, 402, 411, 1048, 1049, 1050)] == [14, 14, 15, 14, 14, 14]
    assert aln2.clip_ends[1748] == 12
    assert aln.clip_starts[525] == 16 and aln.clip_starts[1437] == 84
    assert sum(aln2.insertions[453].values()) == 14 and sum(aln2.insertions[457].values()) == 14


@needs_reference
def test_c_oracle_equals_reference_on_every_reference_fixture():
    """All 17 BAM/SAM fixtures of the reference's test-suite (bact.tiny excluded: 50 s in the
    reference): tables, insertion dicts with their first-seen order, and the vote."""
    import glob

    k = refload.load_reference()
    root = os.path.join(refload.REFERENCE_ROOT, "tests")
    files = sorted(glob.glob(root + "/data_*/*.bam") + glob.glob(root + "/data_*/*.sam"))
    files = [f for f in files if "bact" not in f]
    assert len(files) == 17
    for path in files:
        batch = bamio.read_alignment(path)
        counts, events = coracle.pileup(batch)
        calls = coracle.vote(counts, 1)
        ins = H.events_to_dicts(batch, events)
        alns = k.parse_bam(path)
        assert list(alns) == batch.contig_names
        for c, aln in enumerate(alns.values()):
            t, ref_ins = H.reference_alignment_to_table(aln)
            np.testing.assert_array_equal(H.contig_view(batch, counts, c), t, err_msg=path)
            s0 = int(batch.contig_slot[c])
            for i, d in enumerate(ref_ins):
                assert list(d.items()) == list(ins.get(s0 + i, {}).items())
            _, changes = k.consensus_sequence(aln.weights, aln.insertions, aln.deletions, None, False, 1, False)
            assert H.calls_to_changes(calls[s0:s0 + len(aln.weights)]) == changes


@needs_reference
def test_golden_fasta_files_of_the_reference_suite():
    """The reference's own golden FASTA files (plain + realign), via the oracle loader: 21 of 22,
    the 22nd belongs to a test the reference has commented out (tests/test_kindel.py:281-299)."""
    import glob

    k = refload.load_reference()
    root = os.path.join(refload.REFERENCE_ROOT, "tests")

    def fasta(path):
        recs, name = {}, None
        for line in open(path):
            line = line.strip()
            if line.startswith(">"):
                name = line[1:]
                recs[name] = ""
            elif name:
                recs[name] += line
        return recs

    ok, bad = 0, []
    for d, ext in (("data_bwa_mem", ".bam"), ("data_minimap2", ".bam"), ("data_ext", ".sam")):
        for p in sorted(glob.glob(os.path.join(root, d, "*" + ext))):
            for realign, suf in ((False, ".fa"), (True, ".realign.fa")):
                res = k.bam_to_consensus(p, realign, 1, 7, 0.1, 50, False, False)
                if {r.name: r.sequence for r in res.consensuses} == fasta(os.path.splitext(p)[0] + suf):
                    ok += 1
                else:
                    bad.append(os.path.basename(p) + suf)
    assert ok == 21 and bad == ["3.issue23.bc75.sam.realign.fa"]


def test_python_port_against_golden(manifest, golden_npz):
    """oracle/py_oracle.py (the reference-shaped CPython loop timed by `bench.py --impl reference`)
    reproduces the reference's tables, insertion dicts and consensus on the small golden files."""
    from oracle import py_oracle

    for name in ("mm2_multi", "ext_3_bc75", "mm2_gp120"):
        entry = manifest["files"][name]
        header, records = samdecode.read_alignment_file(golden_input(entry))
        lens = {sn[3:]: int(next(f for f in fields if f.startswith("LN:"))[3:]) for sn, fields in header["@SQ"].items()}
        groups = OrderedDict()
        for r in records:
            if r.rname != "*":
                groups.setdefault(r.rname, []).append(py_oracle.Rec(r.pos, r.mapped, r.seq, r.cigars))
        g = golden_npz(name)
        assert list(groups) == [c["name"] for c in entry["contigs"]]
        for c, (ctg, recs) in enumerate(groups.items()):
            p = py_oracle.pileup(lens[ctg], recs)
            t = g["c%d_counts" % c]
            L = lens[ctg]
            for k, b in enumerate("ACGTN"):
                assert [w[b] for w in p.weights] == t[k, :L].tolist()
                assert [w[b] for w in p.clip_start_weights] == t[9 + k, :L].tolist()
                assert [w[b] for w in p.clip_end_weights] == t[14 + k, :L].tolist()
            assert p.deletions == t[5].tolist() and p.clip_starts == t[7].tolist() and p.clip_ends == t[8].tolist()
            assert list(p.consensus_depth) == g["c%d_consensus_depth" % c].tolist()
            want_ins = {i: [tuple(kv) for kv in items] for i, items in entry["contigs"][c]["insertions"]}
            assert {i: list(d.items()) for i, d in enumerate(p.insertions) if d} == want_ins
            seq, changes = py_oracle.vote(p, 1)
            assert seq == dict(map(tuple, entry["runs"]["plain"]["fasta"]))[ctg + "_cns"]
            assert "".join("-" if c_ is None else c_ for c_ in changes) == entry["runs"]["plain"]["changes"][ctg]


def _digest_check(table_i32, seq, changes, meta):
    import hashlib

    assert hashlib.sha256(np.ascontiguousarray(table_i32, dtype=np.int32).tobytes()).hexdigest() == meta["table_sha256"]
    assert [int(x) for x in table_i32.sum(axis=1)] == meta["column_sums"]
    assert len(seq) == meta["fasta_len"] and hashlib.sha256(seq.encode()).hexdigest() == meta["fasta_sha256"]
    assert hashlib.sha256("".join("-" if c is None else c for c in changes).encode()).hexdigest() == meta["changes_sha256"]


def test_megabase_fixture_digest(manifest):
    """The 6.1 Mb fixture of the reference (`bact.tiny`, secondary alignments with SEQ `*`, 8x depth):
    the oracle's dense table, consensus and changes hash to what the reference produced."""
    from kindel_b200 import kindel as K

    for name, entry in manifest["digests"].items():
        batch = bamio.read_alignment(golden_input(entry))
        counts, events = coracle.pileup(batch)
        calls = coracle.vote(counts, 1)
        ins = H.events_to_dicts(batch, events)
        assert batch.contig_names == [c["name"] for c in entry["contigs"]]
        for c, meta in enumerate(entry["contigs"]):
            s0, L = int(batch.contig_slot[c]), meta["ref_len"]
            want_ins = {s0 + i: [tuple(kv) for kv in items] for i, items in meta["insertions"]}
            assert {s: list(d.items()) for s, d in ins.items()} == want_ins
            seq, changes = K.assemble_consensus(calls[s0:s0 + L],
                                                lambda p: K.dict_consensus(ins.get(s0 + p, {})))
            _digest_check(H.contig_view(batch, counts, c), seq, changes, meta)
