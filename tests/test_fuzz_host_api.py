"""Differential fuzzing of the HOST halves of the API against the unmodified reference (build container only).

For random clip-heavy alignments (tests/clip_cases.py) and the generic fuzz cases (tests/fuzz_cases.py):
`bam_to_consensus` under random option sets (realign on/off, min_depth, min_overlap, clip_decay_threshold,
mask_ends, trim_ends, uppercase) -- sequences, changes and report text -- the CDR helper functions called
the way the reference's tests call them, and the `weights` / `features` frames.  Count tables come from the
CPU oracle (`PileupRun.from_host_tables`), so the code under test is exactly the host code the GPU engine
feeds; no GPU is needed."""
import random

import pandas as pd
import pytest

from clip_cases import clip_case
from fuzz_cases import random_case
from kindel_b200 import bamio
from kindel_b200 import kindel as K
from oracle import coracle, refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree only exists in the build container")

N_CLIP = 160
N_GENERIC = 400  # most generic cases abort in the pileup (by design); the survivors are tiny odd contigs


def _run(path):
    batch = bamio.read_alignment(path)
    counts, events = coracle.pileup(batch)
    return K.PileupRun.from_host_tables(batch, counts, coracle.derive(counts), events), counts


def _options(seed):
    rng = random.Random(seed)
    return dict(realign=rng.random() < 0.8, min_depth=rng.choice([1, 1, 2, 5]), min_overlap=rng.choice([1, 3, 7, 9]),
                clip_decay_threshold=rng.choice([0.0, 0.1, 0.1, 0.3, 0.9]), mask_ends=rng.choice([0, 1, 5, 20, 50]),
                trim_ends=rng.random() < 0.5, uppercase=rng.random() < 0.3)


def _same_regions(a, b):
    return [tuple(r) for r in a] == [tuple(r) for r in b]


def _compare_consensus(k, path, seed, stats):
    o = _options(seed)
    args = (o["realign"], o["min_depth"], o["min_overlap"], o["clip_decay_threshold"], o["mask_ends"], o["trim_ends"],
            o["uppercase"])
    try:
        want = k.bam_to_consensus(str(path), *args)
        ref_err = None
    except (IndexError, KeyError) as exc:
        ref_err = type(exc)
    try:
        run, counts = _run(path)
    except (IndexError, KeyError, ValueError) as exc:
        assert ref_err is not None, (seed, exc)
        return None
    if ref_err is not None:
        # the pileup succeeded on our side, so the reference failed later (host logic): same exception here
        with pytest.raises(ref_err):
            K.consensus_from_run(run, coracle.vote(counts, o["min_depth"]), str(path), *args)
        stats["raised"] += 1
        return None
    got = K.consensus_from_run(run, coracle.vote(counts, o["min_depth"]), str(path), *args)
    assert [(r.name, r.sequence) for r in got.consensuses] == [(r.name, r.sequence) for r in want.consensuses], (seed, o)
    assert dict(got.refs_changes) == dict(want.refs_changes), (seed, o)
    assert dict(got.refs_reports) == dict(want.refs_reports), (seed, o)
    if o["realign"]:
        for rep in want.refs_reports.values():  # reports that list at least one merged clip-dominant region
            line = next(l for l in rep.splitlines() if l.startswith("- clip-dominant regions"))
            stats["patched"] += line.strip() != "- clip-dominant regions:"
    return run


def _compare_cdr_functions(k, path, run, seed, stats):
    """cdr_start / cdr_end / cdrp_consensuses / merge_cdrps on the reference's own list-of-dict tables and on
    the engine's views must give the reference's Regions."""
    o = _options(seed)
    alns = k.parse_bam(str(path))
    for c, (name, aln) in enumerate(alns.items()):
        mine = run.alignment(c)
        a = (aln.weights, aln.deletions, aln.clip_start_weights, aln.clip_start_depth, o["clip_decay_threshold"], o["mask_ends"])
        want_f = k.cdr_start_consensuses(*a)
        assert _same_regions(K.cdr_start_consensuses(*a), want_f), (seed, name, "fwd/lists")
        assert _same_regions(K.cdr_start_consensuses(mine.weights, mine.deletions, mine.clip_start_weights,
                                                     mine.clip_start_depth, a[4], a[5]), want_f), (seed, name, "fwd/views")
        b = (aln.weights, aln.deletions, aln.clip_end_weights, aln.clip_end_depth, o["clip_decay_threshold"], o["mask_ends"])
        want_r = k.cdr_end_consensuses(*b)
        assert _same_regions(K.cdr_end_consensuses(*b), want_r), (seed, name, "rev/lists")
        assert _same_regions(K.cdr_end_consensuses(mine.weights, mine.deletions, mine.clip_end_weights,
                                                   mine.clip_end_depth, b[4], b[5]), want_r), (seed, name, "rev/views")
        p = (aln.weights, aln.deletions, aln.clip_start_weights, aln.clip_end_weights, aln.clip_start_depth,
             aln.clip_end_depth, o["clip_decay_threshold"], o["mask_ends"])
        want_p = k.cdrp_consensuses(*p)
        got_p = K.cdrp_consensuses(*p)
        assert [(tuple(f), tuple(r)) for f, r in got_p] == [(tuple(f), tuple(r)) for f, r in want_p], (seed, name)
        assert _same_regions(K.merge_cdrps(got_p, o["min_overlap"]), k.merge_cdrps(want_p, o["min_overlap"])), (seed, name)
        stats["regions"] += len(want_f) + len(want_r)
        stats["pairs"] += len(want_p)


def _compare_frames(k, path, run, seed):
    rng = random.Random(seed)
    relative, confidence, alpha = rng.random() < 0.5, rng.random() < 0.7, rng.choice([0.01, 0.05, 0.2])
    want = k.weights(str(path), relative, confidence, alpha)
    got = K.weights_from_run(run, relative, confidence, alpha)
    pd.testing.assert_frame_equal(got, want, check_exact=True)
    try:
        want_f = k.features(str(path))
    except IndexError:
        with pytest.raises(IndexError):
            K.features_from_run(run)
        return
    pd.testing.assert_frame_equal(K.features_from_run(run), want_f, check_exact=True)


def test_clip_heavy_cases_against_reference(tmp_path):
    k = refload.load_reference()
    stats = dict(raised=0, patched=0, regions=0, pairs=0)
    for seed in range(N_CLIP):
        path = tmp_path / ("clip%d.sam" % seed)
        path.write_text(clip_case(seed))
        run = _compare_consensus(k, path, seed, stats)
        if run is None:
            continue
        _compare_cdr_functions(k, path, run, seed, stats)
        if seed % 4 == 0:
            _compare_frames(k, path, run, seed)
    # the generator really reaches the realign machinery
    assert stats["regions"] > 100 and stats["pairs"] > 20 and stats["patched"] > 10, stats


def test_generic_cases_host_api_against_reference(tmp_path):
    k = refload.load_reference()
    stats = dict(raised=0, patched=0, regions=0, pairs=0)
    done = 0
    for seed in range(N_GENERIC):
        path = tmp_path / ("gen%d.sam" % seed)
        path.write_text(random_case(seed))
        try:
            run = _compare_consensus(k, path, seed, stats)
        except ValueError:
            continue  # a base outside the BAM alphabet cannot be packed (documented deviation)
        if run is None:
            continue
        done += 1
        _compare_cdr_functions(k, path, run, seed, stats)
        if seed % 3 == 0:
            _compare_frames(k, path, run, seed)
    assert done > 20
