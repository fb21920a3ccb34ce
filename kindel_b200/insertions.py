"""Insertion strings from the event list the pileup kernel emits.

The reference keeps, per reference position, a dict {inserted string -> count} in first-seen order
(reference kindel/kindel.py:38,55-58).  The engine keeps the dense total per slot in count column 6
and one event row (slot, read, q_off, len) per I op, written in the reference's iteration order.
This module rebuilds, lazily and only where someone looks, the dict of a slot and its
`consensus()` (kindel.py:369-381: first maximum in first-seen order, tie = another key with the
same count) -- the vote needs that only at the few slots whose call carries change code 'I'
(kindel.py:419-422).
"""
from __future__ import annotations

import numpy as np

from .bamio import NIBBLES, ReadBatch

_LUT = np.frombuffer(NIBBLES.encode(), dtype=np.uint8)


def decode_events(batch: ReadBatch, rows: np.ndarray) -> list:
    """Upper-case inserted strings of event rows (slot, read, q_off, len), vectorised.
    A string is clipped at the end of SEQ exactly like Python slicing (kindel.py:56)."""
    if rows.shape[0] == 0:
        return []
    read = rows[:, 1].astype(np.int64)
    q0 = rows[:, 2].astype(np.int64)
    lseq = batch.seq_len[read].astype(np.int64)
    # simple reads store the op length in l_seq, but simple reads have no I ops, so lseq is SEQ's
    q1 = np.minimum(q0 + rows[:, 3].astype(np.int64), lseq)
    ln = np.maximum(q1 - q0, 0)
    total = int(ln.sum())
    if total == 0:
        return [""] * rows.shape[0]
    ends = np.cumsum(ln)
    starts = ends - ln
    within = np.arange(total, dtype=np.int64) - np.repeat(starts, ln)
    q = np.repeat(q0, ln) + within
    word = batch.seq4[np.repeat(batch.seq_off[read].astype(np.int64), ln) + (q >> 3)]
    nib = (word >> (28 - 4 * (q & 7)).astype(np.uint32)) & 0xF
    chars = _LUT[nib].tobytes().decode("ascii")
    return [chars[s:e] for s, e in zip(starts.tolist(), ends.tolist())]


class InsertionTable:
    """Events of one pileup, indexed by slot."""

    def __init__(self, batch: ReadBatch, events: np.ndarray):
        self.batch = batch
        events = np.ascontiguousarray(events, dtype=np.int32).reshape(-1, 4)
        order = np.argsort(events[:, 0], kind="stable")  # keep iteration order inside a slot
        self.events = events[order]
        self.slots = self.events[:, 0].astype(np.int64)

    def rows_at(self, slot: int) -> np.ndarray:
        lo = np.searchsorted(self.slots, slot, side="left")
        hi = np.searchsorted(self.slots, slot, side="right")
        return self.events[lo:hi]

    def dict_at(self, slot: int) -> dict:
        """{string: count} in first-seen order (what `insertions[pos]` is in the reference)."""
        out = {}
        for s in decode_events(self.batch, self.rows_at(slot)):
            out[s] = out.get(s, 0) + 1
        return out

    def consensus_at(self, slot: int):
        """(string, tie) = consensus(insertions[pos])[0], [3]  (kindel.py:369-381)."""
        return dict_consensus(self.dict_at(slot))


def dict_consensus(d: dict):
    if not d or not sum(d.values()):
        return "N", False
    best, freq = None, None
    for k, v in d.items():
        if freq is None or v > freq:
            best, freq = k, v
    tie = bool(freq) and any(v == freq for k, v in d.items() if k != best)
    return best, tie
