"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement of the reference's hot path, in the
reference's own shape: one object per record, one dict of counts per reference position.

Why it exists next to the C restatement: the reference is CPython, so the honest "reference CPU
path" figure for `bench.py --impl reference` / `cpu_baseline` is a Python loop over record objects
with per-base dict updates -- the same work per aligned base the reference does
(reference kindel/kindel.py:29-96 for the pileup and its post-pass, :369-430 for the vote).  The
reference itself cannot travel to the GPU box, so this port is what is timed there.  It is pinned
against the goldens in tests/test_oracle_pin.py like the C port (small cases only: it is slow).

Not imported by the product; only tests/ and bench.py's CPU legs use it.
"""
from __future__ import annotations

from collections import defaultdict, namedtuple

Pileup = namedtuple("Pileup", "weights insertions deletions clip_starts clip_ends clip_start_weights "
                              "clip_end_weights consensus_depth clip_start_depth clip_end_depth clip_depth")


class Rec:
    __slots__ = ("pos", "mapped", "seq", "cigars")

    def __init__(self, pos, mapped, seq, cigars):
        self.pos, self.mapped, self.seq, self.cigars = pos, mapped, seq, cigars


def _fresh(n):
    return [{"A": 0, "T": 0, "G": 0, "C": 0, "N": 0} for _ in range(n)]


def base_call(w):
    """kindel.py:369-381 -> (key, count, tie)."""
    if not sum(w.values()):
        return "N", 0, False
    key = max(w, key=w.get)
    cnt = w[key]
    return key, cnt, any(v == cnt for k, v in w.items() if k != key)


def pileup(ref_len, records):
    """kindel.py:29-96: tables for one contig from record objects (1-based .pos)."""
    weights, csw, cew = _fresh(ref_len), _fresh(ref_len), _fresh(ref_len)
    clip_starts, clip_ends, deletions = [0] * (ref_len + 1), [0] * (ref_len + 1), [0] * (ref_len + 1)
    insertions = [defaultdict(int) for _ in range(ref_len + 1)]
    for rec in records:
        if not rec.mapped or len(rec.seq) <= 1:
            continue
        seq = rec.seq
        r, q = rec.pos - 1, 0
        for i, (n, op) in enumerate(rec.cigars):
            if op == "M" or op == "=" or op == "X":
                for _ in range(n):
                    weights[r][seq[q].upper()] += 1
                    r += 1
                    q += 1
            elif op == "I":
                insertions[r][seq[q:q + n].upper()] += 1
                q += n
            elif op == "D":
                for k in range(n):
                    deletions[r + k] += 1
                r += n
            elif op == "S":
                if i == 0:
                    clip_ends[r] += 1
                    for g in range(n):
                        b = seq[g].upper()
                        rel = r - n + g
                        if rel >= 0:
                            cew[rel][b] += 1
                    q += n
                else:
                    clip_starts[r - 1] += 1
                    for _ in range(n):
                        b = seq[q].upper()
                        if r < ref_len:
                            csw[r][b] += 1
                            r += 1
                            q += 1
    # post-pass, kindel.py:83-96
    consensus_depth = [w[base_call(w)[0]] if sum(w.values()) else 0 for w in weights]
    csd = [w["A"] + w["C"] + w["G"] + w["T"] for w in csw]
    ced = [w["A"] + w["C"] + w["G"] + w["T"] for w in cew]
    return Pileup(weights, insertions, deletions, clip_starts, clip_ends, csw, cew, consensus_depth, csd, ced,
                  [a + b for a, b in zip(csd, ced)])


def vote(p, min_depth=1):
    """kindel.py:384-430 without patches/trim: (sequence, changes)."""
    out, changes = [], [None] * len(p.weights)
    n = len(p.weights)
    for pos, w in enumerate(p.weights):
        ins = sum(p.insertions[pos].values()) if p.insertions[pos] else 0
        dele = p.deletions[pos]
        depth = w["A"] + w["C"] + w["G"] + w["T"]
        nxt = p.weights[pos + 1] if pos + 1 < n else None
        depth_next = (nxt["A"] + nxt["C"] + nxt["G"] + nxt["T"]) if nxt else 0
        if dele > depth * 0.5:
            changes[pos] = "D"
        elif depth < min_depth:
            out.append("N")
            changes[pos] = "N"
        else:
            if ins > min(depth * 0.5, depth_next * 0.5):
                key, _, tie = base_call(p.insertions[pos])
                out.append("N" if tie else key.lower())
                changes[pos] = "I"
            key, _, tie = base_call(w)
            out.append("N" if tie else key)
    return "".join(out), changes


_NIB = "=ACMGRSVTWYHKDBN"
_OPS = "MIDNSHP=X"


def records_of(batch, lo=0, hi=None):
    """Record objects for reads [lo, hi) of a flattened batch (decode is not part of the timed path,
    exactly as BAM decode is outside the reference's hot loops)."""
    import numpy as np

    hi = batch.ref_start.shape[0] if hi is None else hi
    lut = np.frombuffer(_NIB.encode(), dtype=np.uint8)
    out = []
    for r in range(lo, hi):
        words = batch.cigar[int(batch.cig_off[r]):int(batch.cig_off[r + 1])].tolist()
        lseq = int(batch.seq_len[r])
        base = int(batch.seq_off[r])
        w = batch.seq4[base:base + (lseq + 7) // 8].astype(np.uint32)
        nib = ((w[:, None] >> np.arange(28, -4, -4, dtype=np.uint32)[None, :]) & 15).reshape(-1)
        seq = lut[nib[:lseq]].tobytes().decode()
        out.append(Rec(int(batch.ref_start[r]) + 1, True, seq, tuple((w >> 4, _OPS[w & 15]) for w in words)))
    return out
