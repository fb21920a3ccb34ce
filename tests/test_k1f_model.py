"""The bit-level arithmetic of the tile-owner kernel, modelled in numpy (tests/k1f_model.py), against
the CPU oracle: funnel-shift extraction of one-hot nibbles, Harley-Seal vertical counters, bit-sliced
quarter sums, plane transposition, N from the coverage identity.  No GPU involved."""
import numpy as np

from k1f_model import Planes, csa, extract8, funnelshift_l, pileup_model, quarter_sum
from kindel_b200 import synth
from oracle import coracle


def test_funnelshift_and_csa_identities():
    rng = np.random.default_rng(0)
    for _ in range(200):
        lo, hi = int(rng.integers(0, 1 << 32)), int(rng.integers(0, 1 << 32))
        sh = int(rng.integers(0, 64))
        assert funnelshift_l(lo, hi, sh) == ((((hi << 32) | lo) << (sh & 31)) >> 32) & 0xFFFFFFFF
        a, b, c = (int(v) for v in rng.integers(0, 1 << 32, size=3))
        carry, s = csa(a, b, c)
        for bit in range(32):
            tot = ((a >> bit) & 1) + ((b >> bit) & 1) + ((c >> bit) & 1)
            assert ((s >> bit) & 1) + 2 * ((carry >> bit) & 1) == tot


def test_vertical_counters_count():
    rng = np.random.default_rng(1)
    streams = [Planes() for _ in range(4)]
    want = np.zeros(32, dtype=np.int64)
    for s in streams:
        for _ in range(25):  # 200 words per stream < 255
            x = [int(v) for v in rng.integers(0, 1 << 32, size=8)]
            s.add8(x)
            for w in x:
                want += (w >> np.arange(32)) & 1
    tot = quarter_sum(streams)
    for bit in range(4):
        got = extract8(tot, bit)
        for b in range(8):
            assert got[b] == want[4 * (7 - b) + bit]


def test_model_equals_oracle_on_sorted_simple_reads():
    for b in (synth.simple_reads(101, [3000], 60), synth.simple_reads(102, [700, 1500, 151], 25, sub_rate=0.2),
              synth.simple_reads(103, [2000], 30, read_len=37, sub_rate=0.3), synth.simple_reads(104, [2500], 12, read_len=301)):
        assert len(b.complex_idx) == 0 and b.reads_sorted
        want, _ = coracle.pileup(b)
        got = pileup_model(b)
        np.testing.assert_array_equal(got, want[:5])
        assert int(got[4].sum()) > 0  # N really occurs and is recovered from coverage
