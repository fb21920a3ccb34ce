"""INTEGRATION.md's ctypes stub -- the binding a maintainer of the reference would add -- is executed as written:
the decode half on the CPU against bamio.read_bam, the pileup + vote half on the GPU against the C oracle."""
import os
import re

import numpy as np
import pytest

from helpers import ROOT
from kindel_b200 import _ffi, bamio, synth
from oracle import coracle


def _stub_namespace():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    code = next(b for b in blocks if "kindel/_b200.py" in b)
    code = code.replace('C.CDLL("libkindel_b200.so")', "C.CDLL(%r)" % _ffi.lib_path())
    _ffi.load()  # builds the library if it is not there
    ns = {}
    exec(compile(code, "INTEGRATION.md:_b200.py", "exec"), ns)
    return ns


def _bam(tmp_path):
    batch = synth.mixed_reads(7, [4000, 2500], 25, 0.3)
    contigs, recs = synth.to_records(batch)
    recs.insert(2, (-1, -1, 4, [], "ACGT"))
    path = tmp_path / "stub.bam"
    bamio.write_bam(path, contigs, recs, level=1)
    return str(path)


def test_stub_structs_match_the_abi():
    ns = _stub_namespace()
    import ctypes as C

    for mine, theirs in ((ns["Batch"], _ffi.KdlBatch), (ns["Diag"], _ffi.KdlDiag)):
        assert C.sizeof(mine) == C.sizeof(theirs)
        assert [(n, getattr(mine, n).offset, getattr(mine, n).size) for n, _ in mine._fields_] == \
               [(n, getattr(theirs, n).offset, getattr(theirs, n).size) for n, _ in theirs._fields_]


def test_decode_stub_equals_read_bam(tmp_path):
    ns = _stub_namespace()
    path = _bam(tmp_path)
    flat = ns["decode"](path, threads=3)
    want = bamio.read_bam(path)
    assert flat.contig_names == want.contig_names and flat.n_slots == want.n_slots
    for f in ("contig_len", "contig_slot", "contig_read_off", "ref_start", "seq_off", "l_seq", "seq4", "complex_idx", "hard_idx"):
        np.testing.assert_array_equal(getattr(flat, f), getattr(want, f), err_msg=f)
    assert (flat.n_events, flat.reads_sorted, flat.reach_right, flat.reach_left, flat.max_simple_len) == \
           (want.n_events, want.reads_sorted, want.reach_right, want.reach_left, want.max_simple_len)


@pytest.mark.gpu
def test_pileup_and_vote_stub_equals_the_oracle(tmp_path):
    ns = _stub_namespace()
    path = _bam(tmp_path)
    flat = ns["decode"](path)
    calls, counts, events = ns["pileup_and_vote"](flat, 2)
    want = bamio.read_bam(path)
    c0, e0 = coracle.pileup(want)
    np.testing.assert_array_equal(counts, c0)
    np.testing.assert_array_equal(events[: len(e0)], e0)
    np.testing.assert_array_equal(calls, coracle.vote(c0, 2))
