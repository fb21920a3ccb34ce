"""`alignment`-shaped views over the engine's count table.

The reference returns, per contig, a 12-field namedtuple of Python lists of dicts
(reference kindel/kindel.py:97-128).  Here the same 12 names (and positions, for tuple unpacking)
are views over one int32 table [19, L+1] copied back from the GPU; dicts are made on access, in
the reference's key order A,T,G,C,N (kindel.py:29 -- the order decides `consensus()` ties on the
realign path).  Call sites served: `weights[i]["A"]`, `len(weights)`, iteration,
slicing `weights[pos:]`, `w.values()`, `insertions[i].values()` / truthiness, integer lists.
"""
from __future__ import annotations

import numpy as np

from .insertions import InsertionTable

FIELDS = ("ref_id", "weights", "insertions", "deletions", "clip_starts", "clip_ends", "clip_start_weights",
          "clip_end_weights", "clip_start_depth", "clip_end_depth", "clip_depth", "consensus_depth")
_DICT_ORDER = (("A", 0), ("T", 3), ("G", 2), ("C", 1), ("N", 4))  # key order of kindel.py:29


class BaseCounts:
    """Sequence of {"A","T","G","C","N"} dicts over five int32 columns [5, L]."""

    __slots__ = ("cols",)

    def __init__(self, cols: np.ndarray):
        self.cols = cols

    def __len__(self):
        return self.cols.shape[1]

    def _row(self, i: int) -> dict:
        c = self.cols
        return {k: int(c[j, i]) for k, j in _DICT_ORDER}

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._row(k) for k in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("list index out of range")
        return self._row(i)

    def __iter__(self):
        for i in range(len(self)):
            yield self._row(i)


class Insertions:
    """Sequence (length L+1) of {string: count} dicts, rebuilt from the event list on access."""

    __slots__ = ("table", "slot0", "n", "totals")

    def __init__(self, table: InsertionTable, slot0: int, n: int, totals: np.ndarray):
        self.table, self.slot0, self.n, self.totals = table, slot0, n, totals

    def __len__(self):
        return self.n

    def _row(self, i: int) -> dict:
        if self.totals[i] == 0:
            return {}
        return self.table.dict_at(self.slot0 + i)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._row(k) for k in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError("list index out of range")
        return self._row(i)

    def __iter__(self):
        for i in range(self.n):
            yield self._row(i)


class Alignment:
    """Drop-in for the reference's `alignment` namedtuple (same field names, order and indexing)."""

    _fields = FIELDS

    def __init__(self, ref_id: str, table: np.ndarray, derived: np.ndarray, ins_table: InsertionTable,
                 slot0: int):
        # table: int32 [19, L+1]; derived: int32 [5, L+1]
        L = table.shape[1] - 1
        self.ref_id = ref_id
        self.table = table
        self.ref_len = L
        self.weights = BaseCounts(table[0:5, :L])
        self.clip_start_weights = BaseCounts(table[9:14, :L])
        self.clip_end_weights = BaseCounts(table[14:19, :L])
        self.insertions = Insertions(ins_table, slot0, L + 1, table[6])
        self._derived = derived
        self._lists = {}

    def _list(self, name, arr):
        v = self._lists.get(name)
        if v is None:
            v = self._lists[name] = arr.tolist()
        return v

    @property
    def deletions(self):
        return self._list("deletions", self.table[5])

    @property
    def clip_starts(self):
        return self._list("clip_starts", self.table[7])

    @property
    def clip_ends(self):
        return self._list("clip_ends", self.table[8])

    @property
    def clip_start_depth(self):
        return self._list("clip_start_depth", self._derived[1, : self.ref_len])

    @property
    def clip_end_depth(self):
        return self._list("clip_end_depth", self._derived[2, : self.ref_len])

    @property
    def clip_depth(self):
        return self._list("clip_depth", self._derived[3, : self.ref_len])

    @property
    def consensus_depth(self):
        # the reference returns a numpy array here (kindel.py:89)
        return self._derived[0, : self.ref_len].astype(np.int64)

    @property
    def acgt_depth(self):
        return self._derived[4, : self.ref_len]

    # namedtuple behaviour
    def __iter__(self):
        return (getattr(self, f) for f in FIELDS)

    def __len__(self):
        return len(FIELDS)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return tuple(getattr(self, f) for f in FIELDS[i])
        return getattr(self, FIELDS[i])

    def _asdict(self):
        return {f: getattr(self, f) for f in FIELDS}

    def __repr__(self):
        return "alignment(ref_id=%r, ref_len=%d)" % (self.ref_id, self.ref_len)
