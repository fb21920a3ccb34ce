"""K0 + the tile-owner kernels (K1f, K1w and the experimental K1f-lean, K1x, K1w2), SOURCE-level, on the CPU.

tests/emu/ compiles kindel_b200/csrc/pileup_tiled.cu and pileup_wide.cu for the host and runs them under a
functional model of the CUDA execution model (tests/emu/cuda_emu.h); the tables must equal the C oracle's.
This checks what a numpy model of the arithmetic (tests/k1f_model.py) cannot: the kernels' own indexing,
sentinels, sub-chunk splitting, flush conditions, coverage scan, fresh/accumulate modes and staging protocol.
It does not model timing or the memory model; the `-m gpu` parity tests remain the proof on the device."""
import numpy as np
import pytest

import emu_harness as E
from kindel_b200 import synth
from oracle import coracle

pytestmark = pytest.mark.skipif(not E.available(), reason="needs g++ and the CUDA headers")

CASES = {
    "shallow":      lambda: synth.simple_reads(5, [3000], 40),
    "deep":         lambda: synth.simple_reads(71, [2000], 3000),            # > 31 blocks per window: mid-window flushes
    "long_reads":   lambda: synth.simple_reads(73, [9000], 600, read_len=6000),  # staging capacity splits sub-chunks
    "sparse":       lambda: synth.simple_reads(74, [50_000], 0.5),           # mostly empty tiles and windows
    "multi_contig": lambda: synth.simple_reads(75, [151, 200, 9000, 333, 160, 700], 40),  # several contigs per tile
    "len_1203":     lambda: synth.simple_reads(76, [20_000], 40, read_len=1203),
    "cfg4_like":    lambda: synth.simple_reads(77, [60_000], 200),
    "short_reads":  lambda: synth.simple_reads(78, [5000], 60, read_len=9),
}


@pytest.mark.parametrize("variant", [E.K1F, E.K1X, E.K1F_LEAN, E.K1W, E.K1W2], ids=["K1f", "K1x", "K1f-lean", "K1w", "K1w2"])
@pytest.mark.parametrize("name", list(CASES))
def test_tile_owner_kernel_source_equals_oracle(name, variant):
    batch = CASES[name]()
    want, _ = coracle.pileup(batch)
    for fresh in (False, True):
        got = E.run_pileup(batch, variant, fresh)
        np.testing.assert_array_equal(got[:5], want[:5], err_msg="%s fresh=%s" % (name, fresh))
        assert not got[5:].any()


@pytest.mark.parametrize("variant", [E.K1F, E.K1X, E.K1F_LEAN, E.K1W, E.K1W2], ids=["K1f", "K1x", "K1f-lean", "K1w", "K1w2"])
def test_accumulate_and_slot_ranges(variant):
    """Two batches added into one table (accumulate mode), then a fresh pass over a tile sub-range only."""
    a = synth.simple_reads(81, [20_000], 30)
    b = synth.simple_reads(81, [20_000], 50, read_seed=9)
    wa, _ = coracle.pileup(a)
    wb, _ = coracle.pileup(b)
    t = E.run_pileup(a, variant, False)
    t = E.run_pileup(b, variant, False, counts=t)
    np.testing.assert_array_equal(t[:5], (wa + wb)[:5])
    # fresh overwrite of tiles [7, 19) with batch a: inside the range a's counts, outside untouched
    lo, n = 7, 12
    t2 = E.run_pileup(a, variant, True, tile_lo=lo, n_tiles=n, counts=t.copy())
    s0, s1 = lo * 512, (lo + n) * 512
    np.testing.assert_array_equal(t2[:5, s0:s1], wa[:5, s0:s1])
    np.testing.assert_array_equal(t2[:5, :s0], t[:5, :s0])
    np.testing.assert_array_equal(t2[:5, s1:], t[:5, s1:])


def test_emulator_catches_a_staging_bug():
    """The emulator is not vacuous: a kernel that skips the wait for its bulk copy reads garbage here."""
    lib = E.load()
    assert lib.emu_selftest_missing_wait() == 1   # data read before mbar_wait differs from the source
    assert lib.emu_selftest_missing_wait_fixed() == 0


# ---- the rest of the kernels: K1s, K1g, diagnose, K2 (vote), derive, K2p (peer vote) ------------------------
def _oracle_outcome(batch):
    try:
        return None, coracle.pileup(batch)
    except IndexError:
        return ("IndexError",), None
    except KeyError as exc:
        return ("KeyError", exc.args[0]), None


def test_whole_pipeline_on_the_fuzz_cases(tmp_path):
    """Every kernel kdl_pileup / kdl_vote / kdl_derive / kdl_diagnose launches, source-level, on the 400 random
    alignments of tests/fuzz_cases.py: tables, insertion events, calls, derived columns and the exception
    (type and KeyError argument) equal the oracle's."""
    from fuzz_cases import random_case

    from kindel_b200 import bamio

    raised = done = 0
    for seed in range(400):
        path = tmp_path / ("fuzz%d.sam" % seed)
        path.write_text(random_case(seed))
        try:
            batch = bamio.read_alignment(path)
        except (ValueError, KeyError):
            continue
        err, res = _oracle_outcome(batch)
        if err:
            raised += 1
            with pytest.raises({"IndexError": IndexError, "KeyError": KeyError}[err[0]]) as exc:
                E.pileup_pipeline(batch)
            if err[0] == "KeyError":
                assert exc.value.args[0] == err[1], seed
            continue
        counts, events = res
        got_c, got_e = E.pileup_pipeline(batch)
        np.testing.assert_array_equal(got_c, counts, err_msg=str(seed))
        np.testing.assert_array_equal(got_e, events, err_msg=str(seed))
        if seed % 4 == 0:
            np.testing.assert_array_equal(E.vote(got_c, 2), coracle.vote(counts, 2), err_msg=str(seed))
            np.testing.assert_array_equal(E.derive(got_c), coracle.derive(counts), err_msg=str(seed))
        done += 1
    assert raised > 20 and done > 20


@pytest.mark.parametrize("variant", [E.K1F, E.K1F_LEAN, E.K1W2], ids=["K1f", "K1f-lean", "K1w2"])
def test_mixed_batch_pipeline(variant):
    """Config-3 shape (clips, indels, edge-case tail; most reads complex, some simple): the tile-owner kernel takes
    the simple reads, K1g the rest, into one table."""
    batch = synth.complex_reads(72, 6000, 60)
    want_c, want_e = coracle.pileup(batch)
    got_c, got_e = E.pileup_pipeline(batch, variant)
    np.testing.assert_array_equal(got_c, want_c)
    np.testing.assert_array_equal(got_e, want_e)
    np.testing.assert_array_equal(E.vote(got_c, 1), coracle.vote(want_c, 1))
    np.testing.assert_array_equal(E.vote(got_c, 7), coracle.vote(want_c, 7))
    np.testing.assert_array_equal(E.derive(got_c), coracle.derive(want_c))


def test_peer_vote_over_footprints():
    """K2p: the vote of the sum of per-rank tables, each read only inside its footprint, on an owner slice."""
    from kindel_b200 import distributed as D

    batch = synth.complex_reads(73, 5000, 40)
    full, _ = coracle.pileup(batch)
    shards = [D.shard_batch(batch, r, 3) for r in range(3)]
    tables = [coracle.pileup(s)[0] for s in shards]
    feet = [D.footprint(s) for s in shards]
    np.testing.assert_array_equal(sum(tables), full)
    want = coracle.vote(full, 2)
    n_slots = full.shape[1]
    for lo, hi in D.owner_slices(n_slots, 3):
        calls, reduced = E.vote_peers(tables, feet, lo, hi, 2, want_reduced=True)
        np.testing.assert_array_equal(calls[lo:hi], want[lo:hi])
        np.testing.assert_array_equal(reduced[:, lo:hi], full[:7, lo:hi])


@pytest.mark.parametrize("world", [2, 3])
def test_fused_exchange_epochs(world):
    """K2x + K2g for all ranks of a read-sharded pileup, two epochs with different data (flags compare epochs):
    every rank ends up with the complete call bytes of the summed table."""
    from kindel_b200 import distributed as D

    flags = {k: [np.zeros(16, dtype=np.int32) for _ in range(world)] for k in ("ready", "done")}
    flags["counter"] = [np.zeros(1, dtype=np.int32) for _ in range(world)]
    calls = None
    for epoch, seed in ((1, 74), (2, 75)):
        batch = synth.complex_reads(seed, 7000, 30)
        full, _ = coracle.pileup(batch)
        shards = [D.shard_batch(batch, r, world) for r in range(world)]
        tables = [coracle.pileup(s)[0] for s in shards]
        feet = [D.footprint(s) for s in shards]
        n_slots = full.shape[1]
        slices = D.owner_slices(n_slots, world)
        if calls is None:
            calls = [np.full(n_slots, 0xEE, dtype=np.uint8) for _ in range(world)]
        E.exchange_epoch(tables, feet, slices, calls, flags, epoch, min_depth=2)
        want = coracle.vote(full, 2)
        for r in range(world):
            np.testing.assert_array_equal(calls[r], want, err_msg="epoch %d rank %d" % (epoch, r))
            assert (flags["ready"][r][:world] == epoch).all() and (flags["done"][r][:world] == epoch).all()


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 1024, 1025, 5000, 262_145, 300_001])
def test_seq_off_scan(n):
    """K-1: exclusive prefix sum of ceil(l_seq / 8) with the complex flag (bit 31) ignored."""
    rng = np.random.default_rng(n)
    l = rng.integers(0, 400, size=n).astype(np.int64)
    flagged = np.where(rng.random(n) < 0.3, l | 0x80000000, l).astype(np.uint32).view(np.int32)
    words = (l + 7) >> 3
    want = np.concatenate(([0], np.cumsum(words)[:-1])).astype(np.uint32)
    np.testing.assert_array_equal(E.seq_off_scan(flagged), want)


def test_dense_layout_predicate():
    from kindel_b200 import engine

    from kindel_b200 import distributed as D

    strided = synth.complex_reads(5, 3000, 30)      # fixed stride per read, shorter reads leave gaps
    assert not engine.seq_is_dense(strided)
    b = D.shard_batch(strided, 1, 2)                 # a shard is re-packed back to back
    assert engine.seq_is_dense(b) and engine.seq_is_dense(synth.simple_reads(6, [2000], 20))
    np.testing.assert_array_equal(E.seq_off_scan(b.l_seq), b.seq_off)
    st, keep = engine.host_struct(b, derive_seq_off=True)
    assert st.seq_off is None and st.l_seq
    b.seq_off[5] += 1
    assert not engine.seq_is_dense(b)
    with pytest.raises(ValueError):
        engine.host_struct(b, derive_seq_off=True)


@pytest.mark.parametrize("schedule,seed", [("reverse", 0), ("random", 1), ("random", 2)])
@pytest.mark.parametrize("variant", [E.K1F, E.K1W, E.K1W2], ids=["K1f", "K1w", "K1w2"])
def test_other_thread_interleavings(variant, schedule, seed):
    """The staging protocols (bulk copy + mbarrier in K1f; full/empty mbarrier ring, producer barrier and
    cp.async prefetch in K1w / K1w2) under other thread orders than the emulator's default."""
    E.set_schedule(schedule, seed)
    try:
        for name in ("shallow", "multi_contig", "long_reads"):
            batch = CASES[name]()
            want, _ = coracle.pileup(batch)
            got = E.run_pileup(batch, variant, fresh=True, grid=4)
            np.testing.assert_array_equal(got[:5], want[:5], err_msg="%s %s/%d" % (name, schedule, seed))
    finally:
        E.set_schedule("forward")
