"""Executable model of the K1f arithmetic (kindel_b200/csrc/pileup_tiled.cu) in numpy -- test infrastructure.

It mirrors, step by step, what one warp of the tile-owner kernel does for its 64-slot window, so the
bit-level tricks are checked on the CPU (tests/test_k1f_model.py) independently of any GPU run:

  * extraction: the 8 bases a read puts on a lane's 8 slots = funnel shift of two words of the read,
    words outside [0, n_words) read as zero;              (kernel: jb / predicated LDS / SHF.L.W)
  * counting: the word's 32 bits are 32 one-bit inputs to vertical counters, 8 reads per Harley-Seal
    carry-save block with a ripple from the 8s plane up;   (Planes::add8)
  * four read streams per window, summed bit-sliced;       (quarter_sum)
  * planes -> integers per nibble and bit;                 (extract8)
  * N is never counted: A+C+G+T (raw) = coverage + 3 N.    (flush_window)
Only simple, coordinate-sorted reads are modelled (what K1f sees); complex reads belong to K1g.
"""
from __future__ import annotations

import numpy as np

F_P = 8          # planes per stream
TILE = 512
WIN = 64
M32 = 0xFFFFFFFF


def funnelshift_l(lo: int, hi: int, shift: int) -> int:
    """High 32 bits of (hi:lo) << (shift & 31)  == __funnelshift_l(lo, hi, shift)."""
    sh = shift & 31
    return ((((hi << 32) | lo) << sh) >> 32) & M32


def csa(a: int, b: int, c: int):
    return ((a & b) | (c & (a | b))) & M32, (a ^ b ^ c) & M32  # carry, sum


class Planes:
    def __init__(self):
        self.p = [0] * F_P

    def add8(self, x):
        p = self.p
        ta, p[0] = csa(p[0], x[0], x[1])
        tb, p[0] = csa(p[0], x[2], x[3])
        fa, p[1] = csa(p[1], ta, tb)
        tc, p[0] = csa(p[0], x[4], x[5])
        td, p[0] = csa(p[0], x[6], x[7])
        fb, p[1] = csa(p[1], tc, td)
        e, p[2] = csa(p[2], fa, fb)
        for k in range(3, F_P):
            t = p[k] & e
            p[k] ^= e
            e = t
        assert e == 0, "a stream overflowed its planes: the kernel flushes every 31 blocks to prevent this"


def quarter_sum(streams):
    """Bit-sliced sum of the four streams' planes: F_P planes in, F_P + 2 out (two butterfly stages)."""
    def add(a, b):
        out, carry = [], 0
        for k in range(max(len(a), len(b))):
            x = a[k] if k < len(a) else 0
            y = b[k] if k < len(b) else 0
            carry, s = csa(x, y, carry)
            out.append(s)
        out.append(carry)
        return out

    s01 = add(streams[0].p, streams[1].p)
    s23 = add(streams[2].p, streams[3].p)
    tot = add(s01, s23)
    assert len(tot) == F_P + 2
    return tot


def extract8(planes, bit):
    """Counter of nibble bit `bit` for the lane's slots 0..7 (slot b sits in nibble 7 - b)."""
    out = []
    for b in range(8):
        pos = 4 * (7 - b) + bit
        out.append(sum(((pl >> pos) & 1) << k for k, pl in enumerate(planes)))
    return out


def pileup_model(batch):
    """weights columns [5, n_slots] of a batch of simple sorted reads, computed the K1f way."""
    n_slots = int(batch.n_slots)
    out = np.zeros((5, n_slots), dtype=np.int64)
    per_contig = np.diff(batch.contig_read_off)
    gstart = (np.repeat(batch.contig_slot, per_contig) + batch.ref_start.astype(np.int64))
    lens = batch.l_seq.astype(np.int64)
    assert (lens > 0).all(), "model covers simple reads only"
    maxlen = int(batch.max_simple_len)
    for tile in range(n_slots // TILE):
        t0 = tile * TILE
        lo = int(np.searchsorted(gstart, t0 - maxlen + 1, side="left"))   # K0: tile index
        hi = int(np.searchsorted(gstart, t0 + TILE, side="left"))
        if lo >= hi:
            continue
        gs = gstart[lo:hi] - t0
        ln = lens[lo:hi]
        # coverage of the tile's slots: +1 / -1 difference array, prefix sum
        diff = np.zeros(TILE + 1, dtype=np.int64)
        cs, ce = np.clip(gs, 0, TILE), np.clip(gs + ln, 0, TILE)
        np.add.at(diff, cs[cs < ce], 1)
        np.add.at(diff, ce[cs < ce], -1)
        cov = np.cumsum(diff)[:TILE]
        for warp in range(TILE // WIN):
            wlo = warp * WIN
            a = int(np.searchsorted(gs, wlo - maxlen + 1, side="left"))
            e = int(np.searchsorted(gs, wlo + WIN, side="left"))
            if a >= e:
                continue
            for lane8 in range(8):                       # the 8 lanes of a quarter (same slots in all quarters)
                p8b = (wlo >> 1) + 4 * lane8
                streams = [Planes() for _ in range(4)]
                base = a & ~7
                blocks = 0
                while base < e:
                    for q in range(4):                   # quarter q: reads base + 8q .. base + 8q + 7
                        x = []
                        for u in range(8):
                            i = base + 8 * q + u
                            if i >= hi - lo:             # sentinel
                                x.append(0)
                                continue
                            g = int(gs[i])
                            nb = ((int(ln[i]) + 7) >> 3) << 2
                            words = batch.seq4[int(batch.seq_off[lo + i]):int(batch.seq_off[lo + i]) + nb // 4]
                            jb = (p8b - (((g + 7) >> 3) << 2)) & M32
                            hw = int(words[jb >> 2]) if jb < nb else 0
                            lw = int(words[((jb + 4) & M32) >> 2]) if ((jb + 4) & M32) < nb else 0
                            x.append(funnelshift_l(lw, hw, ((-g) & 7) << 2))
                        streams[q].add8(x)
                    base += 32
                    blocks += 1
                    assert blocks < 31, "model does not implement the mid-window flush"
                tot = quarter_sum(streams)
                cols = [extract8(tot, bit) for bit in range(4)]
                for b in range(8):
                    slot = wlo + 8 * lane8 + b
                    raw = [cols[bit][b] for bit in range(4)]
                    n3 = sum(raw) - int(cov[slot])
                    assert n3 % 3 == 0 and n3 >= 0
                    n = n3 // 3
                    for bit in range(4):
                        out[bit, t0 + slot] += raw[bit] - n
                    out[4, t0 + slot] += n
    return out


# ---- wide lanes: 16 slots per lane, groups of 4 lanes, 8 read streams per warp ------------------------------
