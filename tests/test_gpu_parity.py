"""Parity of the CUDA engine, through the C ABI, with
  * the golden vectors produced by the unmodified reference (tests/golden/),
  * the CPU oracle on seeded synthetic inputs of the BASELINE.json shapes (full size for configs 2, 3
    and -- marked slow only by its 10^9 bases -- config 4), plus size-independent invariants.
Bit-exact everywhere: this path is integer / byte work."""
import os

import numpy as np
import pytest

import helpers as H
from conftest import golden_input
from test_host_logic import assert_frame_matches

pytestmark = pytest.mark.gpu


def engine_tables(batch):
    import torch

    from kindel_b200 import engine

    db = engine.upload(batch)
    counts, events = engine.pileup(db)
    torch.cuda.synchronize()
    return db, counts, events


def test_native_library_is_the_one_running():
    from kindel_b200 import _ffi, engine

    lib = _ffi.load()
    before = lib.kdl_launch_count()
    from kindel_b200 import synth

    b = synth.simple_reads(0, [2000], 20)
    engine_tables(b)
    assert lib.kdl_launch_count() > before
    assert os.path.basename(_ffi.lib_path()) == "libkindel_b200.so"


def test_golden_files_full_api(manifest, golden_npz):
    from kindel_b200 import kindel as K

    for name, entry in manifest["files"].items():
        path = golden_input(entry)
        alns = K.parse_bam(path)
        g = golden_npz(name)
        assert list(alns) == [c["name"] for c in entry["contigs"]]
        for c, (ctg, aln) in enumerate(alns.items()):
            np.testing.assert_array_equal(aln.table, g["c%d_counts" % c], err_msg=name)
            np.testing.assert_array_equal(aln.consensus_depth, g["c%d_consensus_depth" % c])
            np.testing.assert_array_equal(np.array(aln.clip_start_depth), g["c%d_clip_start_depth" % c])
            np.testing.assert_array_equal(np.array(aln.clip_end_depth), g["c%d_clip_end_depth" % c])
            np.testing.assert_array_equal(np.array(aln.clip_depth), g["c%d_clip_depth" % c])
            want = {i: [tuple(kv) for kv in items] for i, items in entry["contigs"][c]["insertions"]}
            for i in range(len(aln.insertions)):
                d = aln.insertions[i]
                assert list(d.items()) == want.get(i, []), (name, i)
        for tag, realign, md, trim, upper in (("plain", False, 1, False, False), ("realign", True, 1, False, False),
                                              ("opts", False, 5, True, True)):
            res = K.bam_to_consensus(path, realign, md, 7, 0.1, 50, trim, upper)
            want = entry["runs"][tag]
            assert [[r.name, r.sequence] for r in res.consensuses] == want["fasta"], (name, tag)
            for ctg, ch in res.refs_changes.items():
                assert "".join("-" if c is None else c for c in ch) == want["changes"][ctg]
            for ctg, rep in res.refs_reports.items():  # incl. min/max depth reduced on the device
                exp = [l for l in want["reports"][ctg].splitlines() if not l.startswith("- bam_path")]
                assert [l for l in rep.splitlines() if not l.startswith("- bam_path")] == exp, (name, tag)
        assert_frame_matches(K.weights(path), g, "w_")
        assert_frame_matches(K.weights(path, True, True, 0.05), g, "wrel_")
        if entry["features_error"]:
            with pytest.raises(IndexError):
                K.features(path)
        else:
            assert_frame_matches(K.features(path), g, "f_")


def test_known_answer_integers(manifest):  # reference tests/test_kindel.py:63-89, through the GPU
    from kindel_b200 import kindel as K

    a = list(K.parse_bam(golden_input(manifest["files"]["bwa_1_1"])).values())[0]
    b = list(K.parse_bam(golden_input(manifest["files"]["ext_3_bc75"])).values())[0]
    assert a.ref_id == "ENA|EU155341|EU155341.2" and len(a.weights) == 9306
    assert a.weights[0]["A"] == 22 and a.weights[23]["A"] == 57
    assert b.weights[68]["G"] == 1 and b.weights[2368]["T"] == 13
    assert [b.deletions[i] for i in (399, 402, 411, 1048, 1049, 1050)] == [14, 14, 15, 14, 14, 14]
    assert b.clip_ends[1748] == 12 and a.clip_starts[525] == 16 and a.clip_starts[1437] == 84
    assert sum(b.insertions[453].values()) == 14 and sum(b.insertions[457].values()) == 14


def test_edge_cases_and_exceptions(manifest, tmp_path):
    from kindel_b200 import kindel as K

    for case in manifest["edge_cases"]:
        p = tmp_path / (case["name"] + ".sam")
        p.write_text(case["sam"])
        if case["raises"]:
            kind, args = case["raises"]
            with pytest.raises({"IndexError": IndexError, "KeyError": KeyError}[kind]) as exc:
                K.parse_bam(p)
            if kind == "KeyError":
                assert [str(a) for a in exc.value.args] == args, case["name"]
            continue
        alns = K.parse_bam(p)
        assert list(alns) == case["contigs"]
        aln = alns["ctg"]
        np.testing.assert_array_equal(aln.table, np.array(case["counts"]), err_msg=case["name"])
        want = {i: [tuple(kv) for kv in items] for i, items in case["insertions"]}
        for i in range(len(aln.insertions)):
            assert list(aln.insertions[i].items()) == want.get(i, []), case["name"]
        for md in (1, 3):
            res = K.bam_to_consensus(p, False, md, 7, 0.1, 50, False, False)
            assert [[r.name, r.sequence] for r in res.consensuses] == case["fasta_min_depth_%d" % md], case["name"]
            # public consensus_sequence on the views == same answer
            seq, ch = K.consensus_sequence(aln.weights, aln.insertions, aln.deletions, None, False, md, False)
            assert seq == case["fasta_min_depth_%d" % md][0][1]
            assert "".join("-" if c is None else c for c in ch) == case["changes_min_depth_%d" % md]


def test_parse_records_and_plain_dict_inputs():
    """parse_records on record objects; consensus_sequence on plain lists of dicts (a user's own tables)."""
    from kindel_b200 import kindel as K
    from oracle import samdecode

    recs = [samdecode.Record("a", 0, "c", 5, "acgTTTTGGAAACCttt", ((3, "S"), (4, "M"), (2, "I"), (3, "M"), (2, "D"), (2, "M"), (3, "S"))),
            samdecode.Record("b", 4, "c", 5, "ACGT", ((4, "M"),)), samdecode.Record("c", 0, "c", 1, "A", ((1, "M"),))]
    aln = K.parse_records("c", 20, recs)
    assert aln.weights[4]["T"] == 1 and aln.weights[8]["A"] == 1 and aln.deletions[11] == 1
    assert aln.insertions[8] == {"GG": 1} and aln.clip_ends[4] == 1 and aln.clip_starts[14] == 1
    weights = [{"A": 10, "T": 0, "G": 0, "C": 0, "N": 0}, {"A": 0, "T": 0, "G": 0, "C": 10, "N": 0},
               {"A": 0, "T": 0, "G": 4, "C": 0, "N": 6}, {"A": 5, "T": 5, "G": 0, "C": 0, "N": 0},
               {"A": 0, "T": 0, "G": 0, "C": 0, "N": 0}, {"A": 1, "T": 0, "G": 0, "C": 0, "N": 0},
               {"A": 0, "T": 0, "G": 0, "C": 0, "N": 3}, {"A": 10, "T": 0, "G": 0, "C": 0, "N": 0}]
    ins = [{}, {"GG": 6}, {}, {"T": 3, "C": 3}, {}, {}, {}, {"AC": 1}, {}]
    dele = [6, 0, 0, 0, 0, 1, 0, 0, 0]
    seq, ch = K.consensus_sequence(weights, ins, dele, None, False, 1, False)
    assert seq == "ggCNNNNNacA" and ch == ["D", "I", None, "I", "N", "D", "N", "I"]


def _against_oracle(batch, min_depth=1):
    import torch

    from kindel_b200 import engine
    from oracle import coracle

    db, counts, events = engine_tables(batch)
    calls = engine.vote(counts, min_depth)
    derived = engine.derive(counts)
    oc, oe = coracle.pileup(batch)
    np.testing.assert_array_equal(counts.cpu().numpy(), oc)
    np.testing.assert_array_equal(events.cpu().numpy(), oe)
    np.testing.assert_array_equal(calls.cpu().numpy(), coracle.vote(oc, min_depth))
    np.testing.assert_array_equal(derived.cpu().numpy(), coracle.derive(oc))
    # invariants that do not need an oracle
    c = counts.cpu().numpy()
    assert int(c[0:5].sum()) == batch.aligned_bases
    assert int(c[6].sum()) == batch.n_events
    return c


def test_synthetic_config2_full_size():
    from kindel_b200 import synth

    _against_oracle(synth.simple_reads(1, [30000], 2000))


def test_synthetic_config3_full_size_with_edge_tail():
    from kindel_b200 import synth

    b = synth.complex_reads(3, 30000, 5000)
    assert b.n_reads > 1_000_000 and len(b.complex_idx) > 900_000
    _against_oracle(b, min_depth=3)


def test_synthetic_multi_contig():
    from kindel_b200 import synth

    # BASELINE config 5's shape (64 contigs x 100 kb), at 60x instead of 500x to keep the run short
    _against_oracle(synth.simple_reads(5, [100000] * 64, 60))


def test_unsorted_input_is_legal():
    from kindel_b200 import bamio, synth

    b = synth.simple_reads(7, [20000, 30000], 300)
    rng = np.random.default_rng(0)
    parts = []
    for c in range(2):
        lo, hi = int(b.contig_read_off[c]), int(b.contig_read_off[c + 1])
        parts.append(lo + rng.permutation(hi - lo))
    perm = np.concatenate(parts)
    n = b.n_reads
    shuffled = bamio.finalize(b.contig_names, b.contig_len, b.contig_read_off, b.ref_start[perm], b.seq_off[perm],
                              b.seq_len[perm], np.arange(n + 1), b.cigar[perm], b.seq4)
    assert not shuffled.reads_sorted
    _, c0, _ = engine_tables(b)
    _, c1, _ = engine_tables(shuffled)
    np.testing.assert_array_equal(c0.cpu().numpy(), c1.cpu().numpy())
    _against_oracle(shuffled)


def test_host_buffer_entry_point():
    """kdl_ctx_consensus: host pointers in, host buffers out, including the error path."""
    from kindel_b200 import bamio, engine, synth
    from oracle import coracle

    ctx = engine.HostContext(0)
    for b in (synth.complex_reads(21, 20000, 200), synth.simple_reads(22, [50000], 100)):
        counts = np.empty((19, b.n_slots), dtype=np.int32)
        events = np.empty((max(b.n_events, 1), 4), dtype=np.int32)
        calls = ctx.consensus(b, 2, counts_out=counts, events_out=events)
        oc, oe = coracle.pileup(b)
        np.testing.assert_array_equal(counts, oc)
        np.testing.assert_array_equal(events[: b.n_events], oe)
        np.testing.assert_array_equal(calls, coracle.vote(oc, 2))
        t = ctx.last_timing()
        assert t["h2d_ms"] > 0 and t["kernel_ms"] > 0
    bad = synth.simple_reads(23, [5000], 20)
    bad.seq4[7] = (int(bad.seq4[7]) & 0x0FFFFFFF) | 0x30000000  # nibble 3 = 'M' (IUPAC): KeyError('M')
    bad = bamio.finalize(bad.contig_names, bad.contig_len, bad.contig_read_off, bad.ref_start, bad.seq_off,
                         bad.seq_len, bad.cig_off, bad.cigar, bad.seq4)  # re-classify: that read is now hard
    with pytest.raises(KeyError) as exc:
        ctx.consensus(bad, 1)
    assert exc.value.args == ("M",)
    ctx.close()


def test_synthetic_config4_full_size():
    """5 Mb x 200x, 10^9 aligned bases: engine == oracle bit for bit, plus invariants."""
    from kindel_b200 import synth

    b = synth.simple_reads(4, [5_000_000], 200)
    assert b.aligned_bases >= 999_000_000
    c = _against_oracle(b)
    depth = c[0:5].sum(axis=0)
    assert abs(float(depth[1000:-1000].mean()) - 200.0) < 1.0


def test_multi_gpu_exchange_modes_match_one_gpu():
    """N = 2 (or more) ranks over NCCL: fused peer-memory vote and all_reduce+vote == oracle."""
    import subprocess
    import sys

    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(os.path.dirname(os.path.abspath(__file__)), "dist_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "dist parity ok" in res.stdout


def test_public_api_on_two_gpus(manifest, tmp_path):
    """bam_to_consensus / weights / parse_bam with devices=2 (one process per GPU, reads or whole contigs sharded,
    counts exchanged over NVLink) == devices=1, which the golden tests pin to the reference; including --realign
    (needs the reduced 19-column table), a multi-contig file (contig partition) and the exception path."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    import pandas as pd

    from kindel_b200 import bamio, synth
    from kindel_b200 import kindel as K

    paths = [golden_input(manifest["files"][k]) for k in ("bwa_1_1", "mm2_multi", "ext_3_bc75") if k in manifest["files"]]
    multi = tmp_path / "multi.bam"
    contigs, recs = synth.to_records(synth.mixed_reads(41, [3000, 5000, 2500, 4000], 30, 0.3))
    bamio.write_bam(multi, contigs, recs)
    paths.append(str(multi))
    for path in paths:
        for realign in (False, True):
            one = K.bam_to_consensus(path, realign=realign, min_depth=2)
            two = K.bam_to_consensus(path, realign=realign, min_depth=2, devices=2)
            assert [r.sequence for r in one.consensuses] == [r.sequence for r in two.consensuses], (path, realign)
            assert one.refs_changes == two.refs_changes and one.refs_reports == two.refs_reports
        pd.testing.assert_frame_equal(K.weights(path), K.weights(path, devices=2), check_exact=True)
        a1, a2 = K.parse_bam(path), K.parse_bam(path, devices=2)
        assert list(a1) == list(a2)
        for k in a1:
            np.testing.assert_array_equal(a1[k].table, a2[k].table)
            nz = np.flatnonzero(a1[k].table[6])
            assert all(list(a1[k].insertions[int(i)].items()) == list(a2[k].insertions[int(i)].items()) for i in nz)
    # a read that walks off its contig: IndexError on one GPU and on two
    bad = tmp_path / "bad.sam"
    bad.write_text("@SQ\tSN:c\tLN:100\n" + "".join("r%d\t0\tc\t%d\t60\t50M\t*\t0\t0\t%s\t*\n" % (k, p, "ACGTA" * 10)
                                                     for k, p in enumerate([1, 10, 40, 80, 90])))
    for dv in (1, 2):
        with pytest.raises(IndexError):
            K.bam_to_consensus(str(bad), devices=dv)


def test_count_table_reuse_without_memset():
    """CountTable: a reused table is never memset; kernels overwrite the weights and zero the other
    columns only when an earlier pileup dirtied them.  Every result must equal a fresh oracle run."""
    import torch

    from kindel_b200 import distributed as D
    from kindel_b200 import engine, synth
    from oracle import coracle

    seq = [synth.complex_reads(41, 30000, 80), synth.simple_reads(42, [30000], 60), synth.simple_reads(43, [30000], 5),
           synth.complex_reads(44, 30000, 20, edge_tail=False), synth.simple_reads(45, [30000], 700)]
    table = None
    for b in seq:
        db = engine.upload(b)
        if table is None:
            table = engine.CountTable(b.n_slots, db.device)
        assert b.n_slots == table.n_slots
        counts, events = engine.pileup(db, table=table)
        torch.cuda.synchronize()
        oc, oe = coracle.pileup(b)
        np.testing.assert_array_equal(counts.cpu().numpy(), oc)
        np.testing.assert_array_equal(events.cpu().numpy(), oe)
    # range-restricted: a shard only touches (and only cleans) its footprint
    full = synth.simple_reads(46, [200000], 50)
    table = engine.CountTable(full.n_slots, torch.device("cuda", torch.cuda.current_device()))
    want = coracle.pileup(full)[0]
    for rank in (0, 1, 2, 1):
        shard = D.shard_batch(full, rank, 3)
        counts, _ = engine.pileup(engine.upload(shard), table=table, slot_range=D.footprint(shard))
        np.testing.assert_array_equal(counts.cpu().numpy(), coracle.pileup(shard)[0])
    assert int(want[0:5].sum()) == full.aligned_bases


def test_sparse_and_unaligned_layouts():
    """Low depth (most tiles empty), tiny contigs sharing a tile, deep pile-ups (multi-flush, sub-chunks)."""
    from kindel_b200 import synth

    _against_oracle(synth.simple_reads(51, [700_000], 0.4))
    _against_oracle(synth.simple_reads(52, [151, 200, 333, 152, 1000, 77777], 25, read_len=150))
    _against_oracle(synth.simple_reads(53, [5000], 9000))
    _against_oracle(synth.simple_reads(54, [40000], 300, read_len=37))
    _against_oracle(synth.simple_reads(55, [60000], 50, read_len=1203))


def test_tile_index_power_of_32_read_counts():
    """K0's 32-ary search with read counts that fill all 32 buckets exactly (regression: the last
    element of the last bucket being smaller than the key must yield 'end', not bucket -1)."""
    from kindel_b200 import synth

    for L, depth in ((15360, 10), (4800, 32), (153600, 32)):  # 1024, 1024, 32768 reads of 150 bp
        b = synth.simple_reads(61, [L], depth)
        assert b.n_reads in (1024, 32768)
        _against_oracle(b)
    b = synth.simple_reads(62, [2000, 90000], 8)  # reads only at the very start of a long slot space
    _against_oracle(b)


def test_cli_end_to_end(manifest, tmp_path, capsys):
    """`kindel consensus|weights|features|version` == the reference CLI's output (reference
    tests/test_kindel.py:114-238 shell out to `kindel consensus <path>`).  The command functions are
    driven in-process through the argument parser (no interpreter start-up per case); one real
    `python -m kindel` subprocess proves the module entry point."""
    import subprocess
    import sys

    from kindel_b200 import cli

    for name, tag, extra in (("mm2_multi", "plain", []), ("ext_3_bc75", "realign", ["-r"]),
                             ("mm2_gp120", "opts", ["--min-depth", "5", "-t", "-u"]), ("bwa_1_1", "realign", ["--realign"])):
        entry = manifest["files"][name]
        capsys.readouterr()
        assert cli.main(["consensus", *extra, golden_input(entry)]) == 0
        out = capsys.readouterr()
        lines = out.out.strip().split("\n")
        got = [[lines[i][1:], lines[i + 1] if i + 1 < len(lines) else ""] for i in range(0, len(lines), 2)]
        assert got == entry["runs"][tag]["fasta"], (name, tag)
        assert "========================= REPORT ===========================" in out.err
    entry = manifest["files"]["ext_3_bc75"]
    assert cli.main(["weights", golden_input(entry)]) == 0
    text = capsys.readouterr().out
    assert text.split("\n", 1)[0].split("\t") == [
        "chrom", "pos", "A", "C", "G", "T", "N", "insertions", "deletions", "clip_starts", "clip_ends", "depth",
        "consensus", "shannon", "lower_ci", "upper_ci"]
    assert len(text.strip().split("\n")) == 1 + entry["contigs"][0]["ref_len"]
    assert cli.main(["features", golden_input(entry)]) == 0
    assert capsys.readouterr().out.split("\n", 1)[0].split("\t")[:4] == ["chrom", "pos", "A", "C"]
    assert cli.main(["version"]) == 0
    assert capsys.readouterr().out.strip() == "kindel 1.2.1"
    # the module entry point, once, in a fresh interpreter
    env = dict(os.environ, PYTHONPATH=H.ROOT)
    entry = manifest["files"]["mm2_multi"]
    res = subprocess.run([sys.executable, "-m", "kindel", "consensus", golden_input(entry)], capture_output=True,
                         text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = res.stdout.strip().split("\n")
    assert [[lines[i][1:], lines[i + 1]] for i in range(0, len(lines), 2)] == entry["runs"]["plain"]["fasta"]


def test_very_long_complex_read_among_short_reads(tmp_path):
    """A 200 kb read with an indel (complex: K1g) sitting in the middle of sorted short reads: its bases
    sit in the same packed array, larger than one staging buffer of the tile kernel, which must skip
    them (sub-chunk logic) while still counting every short read around it."""
    from kindel_b200 import bamio
    from kindel_b200 import kindel as K

    rng = np.random.default_rng(77)
    L = 400_000
    starts = np.sort(rng.integers(0, L - 150, size=3000))
    lines = ["@HD\tVN:1.6\tSO:coordinate", "@SQ\tSN:big\tLN:%d" % L]
    long_pos = 100_000
    long_seq = "".join(rng.choice(list("ACGT"), size=200_010))
    placed = False
    for k, s in enumerate(starts.tolist()):
        if not placed and s >= long_pos:
            lines.append("long\t0\tbig\t%d\t60\t100000M10I100000M\t*\t0\t0\t%s\t*" % (long_pos + 1, long_seq))
            placed = True
        seq = "".join(rng.choice(list("ACGTN"), size=150, p=[0.245, 0.245, 0.245, 0.245, 0.02]))
        lines.append("r%d\t0\tbig\t%d\t60\t150M\t*\t0\t0\t%s\t*" % (k, s + 1, seq))
    p = tmp_path / "long.sam"
    p.write_text("\n".join(lines) + "\n")
    batch = bamio.read_alignment(p)
    assert batch.reads_sorted and len(batch.complex_idx) == 1 and batch.max_simple_len == 150
    _against_oracle(batch)
    aln = K.parse_bam(p)["big"]
    assert aln.insertions[long_pos + 100_000] == {long_seq[100_000:100_010]: 1}


def test_tile_kernel_shapes():
    """K1 on the shapes that stress its paths: deep piles (several items per tile, mid-window flushes), long reads
    (staging capacity cuts items), sparse and multi-contig tiles, complex-heavy reads at several depths (piece
    lists, item cuts by piece capacity, sparse REDs, insertion events), mixed simple / complex."""
    from kindel_b200 import bamio, synth

    _against_oracle(synth.simple_reads(71, [300_000], 150))
    _against_oracle(synth.simple_reads(73, [4000], 6000))
    _against_oracle(synth.simple_reads(74, [500_000], 0.5))
    _against_oracle(synth.simple_reads(75, [151, 200, 90_000, 333], 40))
    _against_oracle(synth.simple_reads(76, [60_000], 40, read_len=1203))
    _against_oracle(synth.simple_reads(83, [9000], 600, read_len=6000))
    _against_oracle(synth.complex_reads(72, 30_000, 400))
    _against_oracle(synth.complex_reads(77, 3000, 4000, edge_tail=False))
    _against_oracle(synth.complex_reads(78, 200_000, 60, read_len=900))
    _against_oracle(mixed_reads(79, 400_000, 150, 0.05))


def mixed_reads(seed, L, depth, complex_frac):
    """A coordinate-sorted batch of simple reads with a fraction of clip / indel reads mixed in (what a real
    short-read BAM looks like; SURVEY.md 8d config 4: "mostly 150M with ~1 % indel/clip reads")."""
    from kindel_b200 import synth

    return synth.mixed_reads(seed, [L], depth, complex_frac)


@pytest.mark.parametrize("mode", ["atomics", "pieces"])
def test_complex_read_modes(monkeypatch, mode):
    """Tile-eligible complex reads either way kdl_pileup_range can take them (it picks by their share of the batch;
    KDL_CX forces one): as masked pieces through K1's bit-sliced counters, or -- when rare -- with K1 treating them
    as inert and K1e counting their bases with REDs.  Fresh-table mode included (CountTable)."""
    import torch

    from kindel_b200 import engine, synth
    from oracle import coracle

    monkeypatch.setenv("KDL_CX", mode)
    for b in (mixed_reads(79, 400_000, 150, 0.05), synth.complex_reads(72, 30_000, 400), mixed_reads(80, 300_000, 40, 0.5)):
        _against_oracle(b)
        db = engine.upload(b)
        table = engine.CountTable(b.n_slots, db.device)
        table.t.fill_(7)                       # garbage everywhere: the first pileup must overwrite / zero it all
        table.dirty, table.dirty_rest = (0, b.n_slots), True
        for _ in range(2):
            counts, events = engine.pileup(db, table=table)
        torch.cuda.synchronize()
        oc, oe = coracle.pileup(b)
        np.testing.assert_array_equal(counts.cpu().numpy(), oc)
        np.testing.assert_array_equal(events.cpu().numpy(), oe)


@pytest.mark.parametrize("split", [2, 7])
def test_depth_split_of_small_references(monkeypatch, split):
    """Fewer tiles than CTA slots: `split` CTAs share a tile by read range and flush with REDs (kdl_pileup_range
    picks the split from the batch; KDL_SPLIT forces it)."""
    from kindel_b200 import synth

    monkeypatch.setenv("KDL_SPLIT", str(split))
    _against_oracle(synth.simple_reads(1, [30000], 500))
    _against_oracle(synth.complex_reads(3, 30000, 300))
    _against_oracle(synth.simple_reads(75, [151, 200, 9000, 333], 300))


def test_big_bam_file_end_to_end(tmp_path):
    """A 200k-read coordinate-sorted BAM through the whole public path (BGZF inflate, C++ gather, flatten,
    K0/K1f/K2, host assembly): FASTA and changes equal what the oracle's tables give."""
    from kindel_b200 import bamio, synth
    from kindel_b200 import kindel as K
    from oracle import coracle

    b = synth.simple_reads(81, [250_000, 50_000], 100)
    path = tmp_path / "big.bam"
    synth.write_simple_bam(path, b)
    back = bamio.read_alignment(path)
    np.testing.assert_array_equal(back.seq4, b.seq4)
    np.testing.assert_array_equal(back.ref_start, b.ref_start)
    res = K.bam_to_consensus(path, min_depth=3)
    oc, _ = coracle.pileup(b)
    calls = coracle.vote(oc, 3)
    for c, name in enumerate(b.contig_names):
        s0, L = int(b.contig_slot[c]), int(b.contig_len[c])
        want_seq, want_changes = K.assemble_consensus(calls[s0:s0 + L], lambda p: ("", False))
        assert res.consensuses[c].name == name + "_cns"
        assert res.consensuses[c].sequence == want_seq
        assert res.refs_changes[name] == want_changes
        assert "min, max observed depth" in res.refs_reports[name]


def test_exchange_protocol_emulated_on_one_gpu():
    """The multi-GPU exchange kernels (K2x vote over footprint-clipped peer tables, K2g pull of the call
    slices, flag protocol; also K2p) driven for 3 emulated ranks inside ONE process on one GPU: the
    "peer" blocks are ordinary allocations of the same device, and the ranks' kernels are issued in an
    order that never has to spin.  Result of every rank == one-GPU oracle."""
    import ctypes as C

    import torch

    from kindel_b200 import _ffi
    from kindel_b200 import distributed as D
    from kindel_b200 import engine, synth
    from oracle import coracle

    lib = _ffi.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    st = int(torch.cuda.current_stream(dev).cuda_stream)
    world = 3
    for full in (synth.simple_reads(91, [120_000], 80), synth.complex_reads(92, 25_000, 200),
                 synth.simple_reads(93, [30_000, 50_000, 9_000], 40)):
        S = full.n_slots
        oc, _ = coracle.pileup(full)
        want = coracle.vote(oc, 2)
        table_bytes, calls_off, flags_off = 19 * S * 4, 19 * S * 4, 19 * S * 4 + S
        blocks = [torch.zeros(flags_off + 256, dtype=torch.uint8, device=dev) for _ in range(world)]
        shards = [D.shard_batch(full, r, world) for r in range(world)]
        feet = [D.footprint(s) for s in shards]
        slices = D.owner_slices(S, world)
        xs = []
        for r in range(world):
            x = _ffi.KdlExchange()
            x.n_ranks, x.rank = world, r
            for p, blk in enumerate(blocks):
                base = blk.data_ptr()
                x.tables[p], x.calls[p] = base, base + calls_off
                x.ready[p], x.done[p] = base + flags_off, base + flags_off + 64
                x.foot_lo[p], x.foot_hi[p] = feet[p]
                x.slice_lo[p], x.slice_hi[p] = slices[p]
            x.counter = blocks[r].data_ptr() + flags_off + 128
            xs.append(x)
        tables = [blk[:table_bytes].view(torch.int32).view(19, S) for blk in blocks]
        for epoch in (1, 2):  # two steps: the second reuses the tables (lazy zeroing) and the flags
            for r in range(world):
                tab = engine.CountTable(S, dev, tensor=tables[r])
                if epoch == 2:
                    tab.dirty, tab.dirty_rest = feet[r], shards[r].n_complex > 0
                    tab.dirty = engine._tile_align(*feet[r], S)
                engine.pileup(engine.upload(shards[r], dev), check=False, table=tab, slot_range=feet[r])
            for r in range(world):
                _ffi.check(lib.kdl_exchange_signal(C.byref(xs[r]), epoch, st), "signal")
            for r in range(world):
                _ffi.check(lib.kdl_exchange_vote(C.byref(xs[r]), S, 2, epoch, st), "vote")
            for r in range(world):
                _ffi.check(lib.kdl_exchange_wait(C.byref(xs[r]), epoch, st), "wait")
            torch.cuda.synchronize()
            for r in range(world):
                got = blocks[r][calls_off:calls_off + S].cpu().numpy()
                np.testing.assert_array_equal(got, want, err_msg="rank %d epoch %d" % (r, epoch))
        # K2p: the same reduction + vote through kdl_vote_peers_sparse (used with NCCL barrier/all_gather)
        ptrs = (C.c_void_p * world)(*[blk.data_ptr() for blk in blocks])
        flo = (C.c_int64 * world)(*[f[0] for f in feet])
        fhi = (C.c_int64 * world)(*[f[1] for f in feet])
        calls = torch.zeros(S, dtype=torch.uint8, device=dev)
        reduced = torch.zeros((7, S), dtype=torch.int32, device=dev)
        _ffi.check(lib.kdl_vote_peers_sparse(ptrs, flo, fhi, world, S, 0, S, 2, calls.data_ptr(), reduced.data_ptr(), st),
                   "vote_peers")
        torch.cuda.synchronize()
        np.testing.assert_array_equal(calls.cpu().numpy(), want)
        np.testing.assert_array_equal(reduced.cpu().numpy(), oc[:7])


def test_megabase_fixture_digest(manifest):
    """The reference's 6.1 Mb fixture (`bact.tiny`; 8x depth, mostly empty tiles, secondary records with
    SEQ `*`): the engine's dense table, consensus FASTA and changes hash to what the reference produced
    (README.md:39 of the reference warns about megabase genomes; here it is one tile pass)."""
    from test_oracle_pin import _digest_check

    from kindel_b200 import kindel as K

    for name, entry in manifest["digests"].items():
        path = golden_input(entry)
        alns = K.parse_bam(path)
        res = K.bam_to_consensus(path, False, 1, 7, 0.1, 50, False, False)
        assert list(alns) == [c["name"] for c in entry["contigs"]]
        for c, (ctg, aln) in enumerate(alns.items()):
            meta = entry["contigs"][c]
            want_ins = {i: [tuple(kv) for kv in items] for i, items in meta["insertions"]}
            nz = np.flatnonzero(aln.table[6])
            assert {int(i): list(aln.insertions[int(i)].items()) for i in nz} == want_ins
            _digest_check(aln.table, res.consensuses[c].sequence, res.refs_changes[ctg], meta)


def test_clip_heavy_cases_through_the_engine(clip_golden, tmp_path):
    """The deterministic clip-heavy cases (tests/clip_cases.py) end to end through the public API on the GPU:
    tables, --realign consensus, changes and report equal the unmodified reference's
    (tests/golden/clip_cases.json, generated by oracle/make_clip_golden.py)."""
    from clip_cases import clip_case
    from test_host_logic import check_clip_case

    from kindel_b200 import kindel as K

    for case in clip_golden["cases"]:
        path = tmp_path / ("clip%d.sam" % case["seed"])
        path.write_text(clip_case(case["seed"]))
        alns = K.parse_bam(str(path))
        got = K.bam_to_consensus(str(path), *case["options"])
        check_clip_case(case, got, {name: aln.table for name, aln in alns.items()})
