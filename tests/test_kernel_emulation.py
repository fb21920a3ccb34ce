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


def test_emulator_catches_a_staging_bug(tmp_path):
    """The emulator is not vacuous: a kernel that skips the wait for its bulk copy reads garbage here."""
    import ctypes as C

    lib = E.load()
    assert lib.emu_selftest_missing_wait() == 1   # data read before mbar_wait differs from the source
    assert lib.emu_selftest_missing_wait_fixed() == 0
