"""kindel's Python API on top of the B200 engine (drop-in for `kindel.kindel` of bede/kindel 1.2.1).

Same names, signatures, defaults and return shapes as the reference module `kindel/kindel.py`
(SURVEY.md 8b).  What changed is where the two hot loops run:

  parse_records / parse_bam   (reference kindel/kindel.py:21-153)  -> K1 pileup kernels
  consensus_sequence          (reference kindel/kindel.py:384-430) -> K2 vote kernel + host string
                                                                      assembly

Everything the north star leaves on the host stays on the host and is restated here in numpy:
BAM/SAM decode (bamio.py), `--realign` clip-dominant-region reassembly (kindel.py:156-366),
report text (kindel.py:437-485) and the float tails of `weights` / `features` (kindel.py:558-664).
There is no CPU implementation of the pileup or the vote in this package.
"""
from __future__ import annotations

import logging
import os
from collections import OrderedDict, namedtuple

import numpy as np

from . import bamio, engine
from .insertions import InsertionTable, dict_consensus
from .views import Alignment, BaseCounts, Insertions

Region = namedtuple("Region", ["start", "end", "seq", "direction"])
result = namedtuple("result", ["consensuses", "refs_changes", "refs_reports"])

_BASE_CHARS = np.frombuffer(b"ACGTN", dtype=np.uint8)
_CHANGE_LUT = (None, "D", "N", "I")

try:  # the reference wraps consensus sequences in dnaio.Sequence (kindel.py:433-434)
    from dnaio import Sequence as _Sequence
except Exception:  # dnaio not installed: same three attributes

    class _Sequence:
        __slots__ = ("name", "sequence", "qualities")

        def __init__(self, name=None, sequence=None, qualities=None):
            self.name, self.sequence, self.qualities = name, sequence, qualities

        def __repr__(self):
            return "Sequence(name=%r, sequence=%r)" % (self.name, self.sequence)


# ------------------------------------------------------------------------------------ pileup
class PileupRun:
    """One file's pileup on the device: count table, events, and lazily the host copies."""

    def __init__(self, batch: bamio.ReadBatch, device=None):
        self.batch = batch
        self.dbatch = engine.upload(batch, device)
        self.counts, self.events = engine.pileup(self.dbatch)
        self.calls_device = None
        self._host_counts = None
        self._host_derived = None
        self._ins = None

    @classmethod
    def from_host_tables(cls, batch, counts, derived, events):
        """Wrap tables that already sit in host memory (results copied back by another path, e.g.
        the kdl_ctx_* host-buffer call or a multi-GPU reduction).  Does no computation."""
        run = cls.__new__(cls)
        run.batch, run.dbatch, run.counts, run.events, run.calls_device = batch, None, None, None, None
        run._host_counts = np.ascontiguousarray(counts, dtype=np.int32)
        run._host_derived = np.ascontiguousarray(derived, dtype=np.int32)
        run._ins = InsertionTable(batch, events)
        return run

    def vote(self, min_depth=1) -> np.ndarray:
        """K2 over the whole table -> call bytes on the host (the device copy is kept for K5)."""
        self.calls_device = engine.vote(self.counts, min_depth)
        return self.calls_device.cpu().numpy()

    @property
    def ins_table(self) -> InsertionTable:
        if self._ins is None:
            self._ins = InsertionTable(self.batch, self.events.cpu().numpy())
        return self._ins

    @property
    def host_counts(self) -> np.ndarray:
        if self._host_counts is None:
            self._host_counts = self.counts.cpu().numpy()
        return self._host_counts

    @property
    def host_derived(self) -> np.ndarray:
        if self._host_derived is None:
            self._host_derived = engine.derive(self.counts).cpu().numpy()
        return self._host_derived

    def contig_slice(self, c: int):
        s = int(self.batch.contig_slot[c])
        return s, s + int(self.batch.contig_len[c]) + 1

    def alignment(self, c: int) -> Alignment:
        s, e = self.contig_slice(c)
        return Alignment(self.batch.contig_names[c], self.host_counts[:, s:e], self.host_derived[:, s:e],
                         self.ins_table, s)

    def alignments(self) -> OrderedDict:
        return OrderedDict((self.batch.contig_names[c], self.alignment(c)) for c in range(self.batch.n_contigs))


def _op_word(length, op):
    code = bamio._OP_CODE.get(op, 15) if op is not None else 15
    return (int(length) << 4) | code


def flatten_records(ref_id, ref_len, records) -> bamio.ReadBatch:
    """Record objects (.pos 1-based, .mapped, .seq, .cigars) -> ReadBatch of one contig.
    The filter is the reference's (kindel.py:43-46)."""
    ref_start, l_seq, cig_off, cigar, seq_off, parts = [], [], [0], [], [], []
    words = 0
    n_rec = 0
    for rec in records:
        n_rec += 1
        if not rec.mapped or len(rec.seq) <= 1:
            continue
        ref_start.append(int(rec.pos) - 1)
        l_seq.append(len(rec.seq))
        cigar.extend(_op_word(ln, op) for ln, op in rec.cigars)
        cig_off.append(len(cigar))
        enc = bamio.encode_seq(rec.seq)
        seq_off.append(words)
        words += enc.size
        parts.append(enc)
    seq4 = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint32)
    return bamio.finalize([ref_id], np.array([ref_len], dtype=np.int64), np.array([0, len(ref_start)], dtype=np.int64),
                          np.array(ref_start, dtype=np.int64), np.array(seq_off, dtype=np.int64),
                          np.array(l_seq, dtype=np.int64), np.array(cig_off, dtype=np.int64),
                          np.array(cigar, dtype=np.int64), seq4, n_records=n_rec)


def parse_records(ref_id, ref_len, records):
    """Pileup of one contig's records -> `alignment` (reference kindel/kindel.py:21-128)."""
    return PileupRun(flatten_records(ref_id, ref_len, records)).alignment(0)


def _default_devices(devices):
    """`devices` = number of GPUs of this node to shard the pileup over (None: $KINDEL_GPUS, else 1)."""
    if devices is None:
        devices = int(os.environ.get("KINDEL_GPUS", "1") or 1)
    return max(1, int(devices))


def pileup_run(bam_path, devices=None, min_depth=1):
    """(PileupRun, calls) of an alignment file on `devices` GPUs.  devices > 1: one process per GPU, reads (or whole
    contigs) sharded, counts exchanged over NVLink in front of the vote (distributed.run_sharded); the result is
    bit-identical to one GPU."""
    batch = bamio.read_alignment(bam_path)
    devices = _default_devices(devices)
    if devices <= 1:
        run = PileupRun(batch)
        return run, None
    from . import distributed

    calls, counts, derived, events = distributed.run_sharded(batch, devices, min_depth)
    return PileupRun.from_host_tables(batch, counts, derived, events), calls


def parse_bam(bam_path, devices=None):
    """Alignment information for each reference sequence, first-seen order
    (reference kindel/kindel.py:131-153).  devices: extension, see pileup_run."""
    return pileup_run(bam_path, devices)[0].alignments()


# --------------------------------------------------------------------------------- consensus
def consensus(weight):
    """(base, frequency, proportion, tie) of one count dict (reference kindel/kindel.py:369-381):
    first maximum in dict order; ("N", 0) when empty/all zero; tie = another key shares it."""
    total = sum(weight.values())
    base, frequency = "N", 0
    if total:
        first = True
        for k, v in weight.items():
            if first or v > frequency:
                base, frequency, first = k, v, False
    tie = bool(frequency) and any(v == frequency for k, v in weight.items() if k != base)
    proportion = round(frequency / total, 2) if total else 0
    return (base, frequency, proportion, tie)


def _vote_columns(weights, insertions, deletions):
    """The 7 vote columns [7, L+1] from either engine views or plain lists of dicts."""
    L = len(weights)
    cols = np.zeros((7, L + 1), dtype=np.int32)
    if isinstance(weights, BaseCounts):
        cols[0:5, :L] = weights.cols
    else:
        for i, w in enumerate(weights):
            cols[0, i], cols[1, i], cols[2, i], cols[3, i], cols[4, i] = w["A"], w["C"], w["G"], w["T"], w["N"]
    dele = np.asarray(deletions[:L] if not isinstance(deletions, np.ndarray) else deletions[:L], dtype=np.int64)
    cols[5, : dele.shape[0]] = dele
    if isinstance(insertions, Insertions):
        cols[6, :L] = insertions.totals[:L]
    else:
        for i in range(L):
            d = insertions[i]
            cols[6, i] = sum(d.values()) if d else 0
    return cols


def _device_vote(cols: np.ndarray, min_depth) -> np.ndarray:
    import torch

    dev = engine.require_cuda()
    n = cols.shape[1]
    n_pad = (n + 3) // 4 * 4
    t = torch.zeros((7, n_pad), dtype=torch.int32, device=dev)
    t[:, :n] = torch.from_numpy(np.ascontiguousarray(cols)).to(dev)
    return engine.vote(t, min_depth).cpu().numpy()[:n]


def _emit_range(calls, lo, hi, ins_lookup, out, changes):
    """Append the consensus text of positions [lo, hi) to `out` (kindel.py:413-424)."""
    if hi <= lo:
        return
    seg = calls[lo:hi]
    change = (seg >> 4) & 3
    chars = _BASE_CHARS[seg & 7]
    for k in np.flatnonzero(change).tolist():
        changes[lo + k] = _CHANGE_LUT[change[k]]
    ins_pos = np.flatnonzero(change == 3)
    keep = change != 1
    if ins_pos.size == 0:
        out.append(chars[keep].tobytes().decode("ascii"))
        return
    prev = 0
    for k in ins_pos.tolist():
        out.append(chars[prev:k][keep[prev:k]].tobytes().decode("ascii"))
        s, tie = ins_lookup(lo + k)
        out.append("N" if tie else s.lower())
        prev = k
    out.append(chars[prev:][keep[prev:]].tobytes().decode("ascii"))


def assemble_consensus(calls, ins_lookup, cdr_patches=None, trim_ends=False, uppercase=False):
    """Call bytes of one contig (length L) -> (consensus string, changes list).

    Restates the sequential part of consensus_sequence (kindel.py:387-401, 425-430): CDR patches
    (first Region whose start == pos, provided some Region starting there has a truthy seq) emit
    their lower-cased sequence and skip `end - start - 1` further positions without looking at them.
    """
    L = calls.shape[0]
    changes = [None] * L
    out = []
    starts = sorted({r.start for r in cdr_patches if r.seq and 0 <= r.start < L}) if cdr_patches else []
    pos = 0
    for st in starts:
        if st < pos:
            continue  # lies inside a span that is being skipped
        _emit_range(calls, pos, st, ins_lookup, out, changes)
        patch = next(r for r in cdr_patches if r.start == st)
        out.append(patch.seq.lower())
        skip = (patch.end - patch.start) - 1
        if skip < 0:  # the reference's counter goes negative and never recovers: nothing more is emitted
            pos = L
            break
        pos = st + 1 + skip
    _emit_range(calls, pos, L, ins_lookup, out, changes)
    seq = "".join(out)
    if trim_ends:
        seq = seq.strip("N")
    if uppercase:
        seq = seq.upper()
    return seq, changes


def consensus_sequence(weights, insertions, deletions, cdr_patches, trim_ends, min_depth, uppercase):
    """Per-position vote -> (consensus string, changes) (reference kindel/kindel.py:384-430).
    The vote itself runs on the GPU (K2); strings are assembled here."""
    calls = _device_vote(_vote_columns(weights, insertions, deletions), min_depth)[: len(weights)]
    return assemble_consensus(calls, lambda p: dict_consensus(insertions[p]), cdr_patches, trim_ends, uppercase)


def consensus_seqrecord(consensus, ref_id):
    return _Sequence(name=f"{ref_id}_cns", sequence=consensus, qualities=None)


# ---------------------------------------------------------------- realign (host, kindel.py:156-366)
def _first_max_base(cols: np.ndarray) -> np.ndarray:
    """consensus(w)[0] for every column of a [5, n] A,C,G,T,N block: first max in A,T,G,C,N order."""
    order = np.array([0, 3, 2, 1, 4])
    stacked = cols[order]
    idx = order[np.argmax(stacked, axis=0)]
    idx = np.where(cols.sum(axis=0) == 0, 4, idx)
    return _BASE_CHARS[idx]


def _cols_of(base_counts) -> np.ndarray:
    if isinstance(base_counts, BaseCounts):
        return np.asarray(base_counts.cols, dtype=np.int64)
    n = len(base_counts)
    cols = np.zeros((5, n), dtype=np.int64)
    for i, w in enumerate(base_counts):
        cols[:, i] = (w["A"], w["C"], w["G"], w["T"], w["N"])
    return cols


def _masked(n: int, mask_ends: int) -> np.ndarray:
    m = np.zeros(n, dtype=bool)
    r = range(n)
    m[list(r[:mask_ends])] = True
    m[list(r[-mask_ends:])] = True  # mask_ends == 0 masks everything, like positions[-0:]
    return m


def _cdr_inputs(weights, deletions, clip_weights, clip_depth, clip_decay_threshold, mask_ends):
    w = _cols_of(weights)
    n = w.shape[1]
    depth = w.sum(axis=0)  # all five keys (sum(w.values()), kindel.py:182)
    dele = np.asarray(deletions, dtype=np.int64)[:n]
    cd = np.asarray(clip_depth, dtype=np.int64)[:n]
    dominant = (cd / (depth + dele + 1) > 0.5) & ~_masked(n, mask_ends)
    extend = cd > (depth + dele) * clip_decay_threshold
    bases = _first_max_base(_cols_of(clip_weights))
    return n, dominant, extend, bases


def _start_regions(n, dominant, extend, bases):
    """-> regions from the per-position predicates (reference kindel/kindel.py:156-213)."""
    stops = np.flatnonzero(~extend)
    regions = []
    for pos in np.flatnonzero(dominant).tolist():
        if any(r.start <= pos < r.end for r in regions):
            continue
        k = np.searchsorted(stops, pos)
        if k < stops.shape[0]:
            end = int(stops[k])
            seq_end = end
        else:  # ran to the contig end without decaying
            end = n - 1
            seq_end = n
        regions.append(Region(pos, end, bases[pos:seq_end].tobytes().decode("ascii"), "\u2192"))
    return regions


def _end_regions(n, dominant, extend, bases):
    """<- regions from the per-position predicates (reference kindel/kindel.py:216-275)."""
    stops = np.flatnonzero(~extend)
    regions = []
    for pos in np.flatnonzero(dominant)[::-1].tolist():
        if any(r.start <= pos < r.end for r in regions):
            continue
        # extension walks pos-1, pos-2, ... and stops at the first position that has decayed
        k = np.searchsorted(stops, pos) - 1  # last stop < pos
        if pos == 0:
            start, seq = 0, ""
        elif k >= 0:
            start = int(stops[k])
            seq = bases[start + 1:pos + 1].tobytes().decode("ascii") if start < pos - 1 else ""
        else:
            start = 0
            seq = bases[0:pos + 1].tobytes().decode("ascii")
        regions.append(Region(start, pos + 1, seq, "\u2190"))
    return regions


def cdr_start_consensuses(weights, deletions, clip_start_weights, clip_start_depth, clip_decay_threshold,
                          mask_ends):
    """Right-clipped (->) consensuses of clip-dominant regions (reference kindel/kindel.py:156-213)."""
    return _start_regions(*_cdr_inputs(weights, deletions, clip_start_weights, clip_start_depth, clip_decay_threshold,
                                       mask_ends))


def cdr_end_consensuses(weights, deletions, clip_end_weights, clip_end_depth, clip_decay_threshold, mask_ends):
    """Left-clipped (<-) consensuses of clip-dominant regions (reference kindel/kindel.py:216-275)."""
    return _end_regions(*_cdr_inputs(weights, deletions, clip_end_weights, clip_end_depth, clip_decay_threshold,
                                     mask_ends))


def _pair_regions(fwd, rev):
    pairs = []
    for f in fwd:
        for r in rev:
            if max(f.start, r.start) < min(f.end, r.end):
                pairs.append((f, r))
                break
    return pairs


def cdrps_from_device(counts, s, e, clip_decay_threshold, mask_ends):
    """cdrp_consensuses for the contig at slots [s, e) straight from the device table: K4 evaluates the
    clip-dominance and decay predicates and the clip consensus bases per position (2 bytes per position come
    back instead of the 76-byte table row); regions, pairing and merging are the same host code."""
    n = e - s
    flags, bases = engine.cdr_flags(counts, s, e, clip_decay_threshold)
    keep = ~_masked(n, mask_ends)
    fwd = _start_regions(n, ((flags & 1) != 0) & keep, (flags & 2) != 0, _BASE_CHARS[bases & 7])
    rev = _end_regions(n, ((flags & 4) != 0) & keep, (flags & 8) != 0, _BASE_CHARS[(bases >> 4) & 7])
    return _pair_regions(fwd, rev)


def cdrp_consensuses(weights, deletions, clip_start_weights, clip_end_weights, clip_start_depth, clip_end_depth,
                     clip_decay_threshold, mask_ends):
    """Pairs of overlapping -> / <- clip consensuses (reference kindel/kindel.py:278-320)."""
    fwd = cdr_start_consensuses(weights, deletions, clip_start_weights, clip_start_depth, clip_decay_threshold,
                                mask_ends)
    rev = cdr_end_consensuses(weights, deletions, clip_end_weights, clip_end_depth, clip_decay_threshold,
                              mask_ends)
    return _pair_regions(fwd, rev)


def merge_by_lcs(s1, s2, min_overlap):
    """Superstring of s1 and s2 about their longest common substring if it is at least
    min_overlap long, else None (reference kindel/kindel.py:323-347).  Among equally long common
    substrings the one ending first in s1 wins, as in the reference's row-major scan."""
    longest, x_longest = 0, 0
    if s1 and s2:
        b = np.frombuffer(s2.encode("utf-32-le"), dtype=np.uint32)
        prev = np.zeros(b.shape[0] + 1, dtype=np.int64)
        for x, ch in enumerate(s1, start=1):
            row = np.zeros_like(prev)
            hit = b == ord(ch)
            row[1:][hit] = prev[:-1][hit] + 1
            m = int(row.max())
            if m > longest:
                longest, x_longest = m, x
            prev = row
    lcs = s1[x_longest - longest:x_longest]
    if len(lcs) < min_overlap:
        return None
    return s1.split(lcs, 1)[0] + lcs + s2.split(lcs, 1)[1]


def merge_cdrps(cdrps, min_overlap):
    """Merged clip-dominant region pairs as Regions (reference kindel/kindel.py:350-366)."""
    merged = []
    for fwd_cdr, rev_cdr in cdrps:
        seq = merge_by_lcs(fwd_cdr.seq, rev_cdr.seq, min_overlap)
        if not seq:
            logging.warning(
                f"No overlap found for clip dominant region spanning positions {fwd_cdr.start}-{rev_cdr.end} (min_overlap = {min_overlap})"
            )
        merged.append(Region(fwd_cdr.start, rev_cdr.end, seq, None))
    return merged


# -------------------------------------------------------------------------------------- report
DepthRange = namedtuple("DepthRange", ["dmin", "dmax"])  # min / max ACGT depth of a contig (kindel.py:450,477-479)


def build_report(ref_id, weights, changes, cdr_patches, bam_path, realign, min_depth, min_overlap,
                 clip_decay_threshold, trim_ends, uppercase):
    """REPORT text block (reference kindel/kindel.py:437-485)."""
    if isinstance(weights, DepthRange):  # already reduced on the device: no table copy needed
        dmin, dmax = weights.dmin, weights.dmax
    elif isinstance(weights, BaseCounts):
        acgt = weights.cols[0:4].sum(axis=0)
        dmin, dmax = (int(acgt.min()), int(acgt.max()))
    else:
        depths = [w["A"] + w["C"] + w["G"] + w["T"] for w in weights]
        dmin, dmax = min(depths), max(depths)
    sites = getattr(changes, "sites", None)  # (a list built by _changes_list knows its sites already)
    if sites is None:
        sites = {"N": [], "I": [], "D": []}
        for pos, change in enumerate(changes, start=1):
            if change in sites:
                sites[change].append(str(pos))
    patches = ["{}-{}: {}".format(r.start, r.end, r.seq) for r in cdr_patches] if cdr_patches else ""
    lines = [
        "========================= REPORT ===========================",
        "reference: {}".format(ref_id),
        "options:",
        "- bam_path: {}".format(bam_path),
        "- min_depth: {}".format(min_depth),
        "- realign: {}".format(realign),
        "    - min_overlap: {}".format(min_overlap),
        "    - clip_decay_threshold: {}".format(clip_decay_threshold),
        "- trim_ends: {}".format(trim_ends),
        "- uppercase: {}".format(uppercase),
        "observations:",
        "- min, max observed depth: {}, {}".format(dmin, dmax),
        "- ambiguous sites: {}".format(", ".join(sites["N"])),
        "- insertion sites: {}".format(", ".join(sites["I"])),
        "- deletion sites: {}".format(", ".join(sites["D"])),
        "- clip-dominant regions: {}".format(", ".join(patches)),
    ]
    return "\n".join(lines) + "\n"


# --------------------------------------------------------------------------------- public API
def bam_to_consensus(bam_path, realign=False, min_depth=1, min_overlap=9, clip_decay_threshold=0.1,
                     mask_ends=50, trim_ends=False, uppercase=False, devices=None):
    """Consensus sequence(s) of an alignment file (reference kindel/kindel.py:488-555).

    Device work per file: one pileup (K1) and one vote (K2) over all contigs at once; only the
    call bytes, the insertion events and -- for --realign and the report -- count columns come
    back to the host.  `devices` (extension; default $KINDEL_GPUS or 1) shards the pileup over that many GPUs of
    the node."""
    run, calls = pileup_run(bam_path, devices, min_depth)
    if calls is None:
        calls = run.vote(min_depth)
    return consensus_from_run(run, calls, bam_path, realign, min_depth, min_overlap,
                              clip_decay_threshold, mask_ends, trim_ends, uppercase)


class _Changes(list):
    """The reference's `changes` list (None / 'D' / 'N' / 'I' per position) that also remembers where its few
    non-None entries are, so the report needs no pass over millions of Nones."""

    __slots__ = ("sites",)


def _changes_list(calls):
    """Per-position change codes (None / 'D' / 'N' / 'I') of one contig's call bytes."""
    change = (calls >> 4) & 3
    out = _Changes([None] * calls.shape[0])
    at = np.flatnonzero(change)
    out.sites = {"N": [], "I": [], "D": []}
    for k, code in zip(at.tolist(), change[at].tolist()):
        name = _CHANGE_LUT[code]
        out[k] = name
        out.sites[name].append(str(k + 1))
    return out


def _device_texts(run, calls_all):
    """K5: the consensus text of every contig assembled on the device (emitted-length scan + scatter); only the
    insertion strings of the 'I' sites are resolved on the host (from the event list) and handed over."""
    batch = run.batch
    is_ins = ((calls_all >> 4) & 3) == 3
    slots = np.flatnonzero(is_ins)
    if slots.size:  # positions only: the extra slot behind a contig never emits (kindel.py:390 loops over weights)
        c = np.searchsorted(batch.contig_slot, slots, side="right") - 1
        slots = slots[slots < batch.contig_slot[c] + batch.contig_len[c].astype(np.int64)]
    strings = []
    for sl in slots.tolist():
        text, tie = run.ins_table.consensus_at(sl)
        strings.append("N" if tie else text.lower())
    return engine.assemble(run.calls_device, batch, slots, strings)


def consensus_from_run(run, calls_all, bam_path, realign=False, min_depth=1, min_overlap=9,
                       clip_decay_threshold=0.1, mask_ends=50, trim_ends=False, uppercase=False):
    """Host half of bam_to_consensus: per contig, optional CDR patches, string assembly, report."""
    ins_table = run.ins_table
    consensuses, refs_changes, refs_reports = [], {}, {}
    on_device = run.counts is not None
    texts = _device_texts(run, calls_all) if (on_device and not realign and run.calls_device is not None) else None
    for c, ref_id in enumerate(run.batch.contig_names):
        s, e = run.contig_slice(c)
        if on_device:
            # the call bytes, the insertion events and, for the report, the min / max ACGT depth are all that is
            # needed: reduced on the device instead of copying 76 B per position back
            d = run.counts[0:4, s:e - 1].sum(dim=0)
            report_weights = DepthRange(int(d.min().item()), int(d.max().item())) if e - 1 > s else DepthRange(0, 0)
        else:
            aln = run.alignment(c)  # host tables (another path produced them)
            report_weights = aln.weights
        if realign:
            if on_device:
                cdrps = cdrps_from_device(run.counts, s, e - 1, clip_decay_threshold, mask_ends)
            else:
                cdrps = cdrp_consensuses(aln.weights, aln.deletions, aln.clip_start_weights, aln.clip_end_weights,
                                         aln.clip_start_depth, aln.clip_end_depth, clip_decay_threshold, mask_ends)
            cdr_patches = merge_cdrps(cdrps, min_overlap)
        else:
            cdr_patches = None
        if texts is not None:
            cons, changes = texts[c], _changes_list(calls_all[s:e - 1])
            if trim_ends:
                cons = cons.strip("N")
            if uppercase:
                cons = cons.upper()
        else:
            cons, changes = assemble_consensus(calls_all[s:e - 1], lambda p, s=s: ins_table.consensus_at(s + p),
                                               cdr_patches, trim_ends, uppercase)
        report = build_report(ref_id, report_weights, changes, cdr_patches, bam_path, realign, min_depth,
                              min_overlap, clip_decay_threshold, trim_ends, uppercase)
        consensuses.append(consensus_seqrecord(cons, ref_id))
        refs_reports[ref_id] = report
        refs_changes[ref_id] = changes
    return result(consensuses, refs_changes, refs_reports)


def weights(bam_path: "path to SAM/BAM file", relative: "output relative nucleotide frequencies" = False,
            confidence: "calculate confidence interval" = True, confidence_alpha: "confidence interval alpha" = 0.01,
            devices=None):
    """DataFrame of per-site nucleotide frequencies, depth, consensus, clip starts/ends, confidence
    interval and entropy (reference kindel/kindel.py:558-630).  Integer columns come from the GPU
    table; the float tail is the reference's arithmetic, vectorised.  devices: extension, see pileup_run."""
    return weights_from_run(pileup_run(bam_path, devices)[0], relative, confidence, confidence_alpha)


def weights_from_run(run, relative=False, confidence=True, confidence_alpha=0.01):
    """Host half of `weights`: DataFrame from the count table of a finished pileup."""
    import pandas as pd
    import scipy.stats

    tab = run.host_counts
    frames = []
    for c, chrom in enumerate(run.batch.contig_names):
        s, e = run.contig_slice(c)
        L = e - s - 1
        t = tab[:, s:e].astype(np.int64)
        frames.append(pd.DataFrame({
            "chrom": [chrom] * L, "pos": np.arange(1, L + 1, dtype=np.int64),
            "A": t[0, :L], "C": t[1, :L], "G": t[2, :L], "T": t[3, :L], "N": t[4, :L],
            "insertions": t[6, 1:L + 1],  # row i reports the insertions of slot i (kindel.py:581)
            "deletions": t[5, :L], "clip_starts": t[7, :L], "clip_ends": t[8, :L],
        }))
    cols = ["chrom", "pos", "A", "C", "G", "T", "N", "insertions", "deletions", "clip_starts", "clip_ends"]
    weights_df = pd.concat(frames, ignore_index=True) if frames else pd.DataFrame(columns=cols)
    six = ["A", "C", "G", "T", "N", "deletions"]
    weights_df["depth"] = weights_df[six].sum(axis=1)
    consensus_depths = weights_df[six].max(axis=1)
    weights_df["consensus"] = consensus_depths.divide(weights_df.depth)
    rel = pd.DataFrame()
    for nt in six:
        rel[[nt]] = weights_df[[nt]].divide(weights_df.depth, axis=0)
        rel = rel.round({k: 4 for k in six})
    acgt = rel[["A", "C", "G", "T"]].values
    with np.errstate(invalid="ignore", divide="ignore"):
        weights_df["shannon"] = scipy.stats.entropy(acgt, axis=1) if len(acgt) else []
    if confidence:
        cnt = consensus_depths.to_numpy()
        nobs = weights_df["depth"].to_numpy()
        # Jeffreys interval per (count, depth) pair (kindel.py:569-574,619-624): an elementwise function of two small
        # integers, and megabases of positions share a few thousand distinct pairs -- evaluate those, gather the rest
        # (the same scipy call on the same inputs: bit-identical to the per-row result)
        base = int(nobs.max()) + 1 if len(nobs) else 1
        pair, inverse = np.unique(cnt.astype(np.int64) * base + nobs.astype(np.int64), return_inverse=True)
        ucnt, unobs = pair // base, pair % base
        lower, upper = scipy.stats.beta.interval(1 - confidence_alpha, ucnt + 0.5, unobs - ucnt + 0.5)
        weights_df["lower_ci"] = np.asarray(lower)[inverse]
        weights_df["upper_ci"] = np.asarray(upper)[inverse]
    if relative:
        for nt in ["A", "C", "G", "T", "N"]:
            weights_df[[nt]] = rel[[nt]]
    return weights_df.round(dict(consensus=3, lower_ci=3, upper_ci=3, shannon=3))


def variants(bam_path: "path to SAM/BAM file", abs_threshold: "absolute frequency above which to call variants" = 1,
             rel_threshold: "relative frequency (0.0-1.0) above which to call variants" = 0.01,
             only_variants: "exclude invariant sites from output" = False,
             absolute: "report absolute variant frequencies" = False, devices=None):
    """EXTENSION -- not in the reference snapshot.  The reference's README (README.md:106-107) lists a `variants`
    sub-command ("Output variants exceeding specified absolute and relative frequency thresholds") but its code
    (kindel/kindel.py, kindel/cli.py) has no such function, so there is nothing to be bit-exact with: parity
    unpinned.  Defined here as a host-side filter over the same integer table `weights`
    reports: per site, every allele (A, C, G, T, N, deletion) other than the site's most frequent one whose count
    exceeds `abs_threshold` AND whose share of the depth (A+C+G+T+N+deletions, as in `weights`) exceeds
    `rel_threshold`.  Columns: chrom, pos, depth, consensus (allele letter, `-` = deletion), then one column per
    allele holding its relative (default) or absolute frequency where it is a variant and 0 elsewhere."""
    return variants_from_run(pileup_run(bam_path, devices)[0], abs_threshold, rel_threshold, only_variants, absolute)


def variants_from_run(run, abs_threshold=1, rel_threshold=0.01, only_variants=False, absolute=False):
    """Host half of `variants` (extension; see there)."""
    import pandas as pd

    tab = run.host_counts
    alleles = ["A", "C", "G", "T", "N", "deletions"]
    rows = [0, 1, 2, 3, 4, 5]
    frames = []
    for c, chrom in enumerate(run.batch.contig_names):
        s, e = run.contig_slice(c)
        L = e - s - 1
        t = tab[rows, s:s + L].astype(np.int64)                       # [6, L]
        depth = t.sum(axis=0)
        top = t.argmax(axis=0)                                        # first maximum in A,C,G,T,N,del order
        with np.errstate(invalid="ignore", divide="ignore"):
            share = np.where(depth > 0, t / np.maximum(depth, 1), 0.0)
        is_var = (t > abs_threshold) & (share > rel_threshold) & (np.arange(6)[:, None] != top[None, :])
        value = np.where(is_var, t if absolute else np.round(share, 4), 0)
        df = pd.DataFrame({"chrom": [chrom] * L, "pos": np.arange(1, L + 1, dtype=np.int64), "depth": depth,
                           "consensus": np.where(depth > 0, np.array(list("ACGTN-"))[top], "N")})
        for k, a in enumerate(alleles):
            df[a] = value[k]
        if only_variants:
            df = df[is_var.any(axis=0)]
        frames.append(df)
    cols = ["chrom", "pos", "depth", "consensus"] + alleles
    return pd.concat(frames, ignore_index=True) if frames else pd.DataFrame(columns=cols)


def features(bam_path: "path to SAM/BAM file", devices=None):
    """DataFrame of relative per-site nucleotide frequencies, indels and entropy
    (reference kindel/kindel.py:633-664), including its indexing of `i`/`d` by global row number
    into the LAST contig's tables (IndexError on most multi-contig files).
    devices: extension, see pileup_run."""
    return features_from_run(pileup_run(bam_path, devices)[0])


def features_from_run(run):
    """Host half of `features`."""
    import pandas as pd
    import scipy.stats

    tab = run.host_counts
    frames = []
    last = None
    for c, chrom in enumerate(run.batch.contig_names):
        s, e = run.contig_slice(c)
        L = e - s - 1
        t = tab[:, s:e].astype(np.int64)
        last = t
        frames.append(pd.DataFrame({"chrom": [chrom] * L, "pos": np.arange(1, L + 1, dtype=np.int64),
                                    "A": t[0, :L], "C": t[1, :L], "G": t[2, :L], "T": t[3, :L], "N": t[4, :L]}))
    df = pd.concat(frames, ignore_index=True) if frames else pd.DataFrame(
        columns=["chrom", "pos", "A", "C", "G", "T", "N"])
    n_rows = len(df)
    if n_rows:
        if n_rows > last.shape[1]:
            raise IndexError("list index out of range")  # aln.insertions[pos], kindel.py:645
        df["i"] = last[6, :n_rows]
        df["d"] = last[5, :n_rows]
    else:
        df["i"] = []
        df["d"] = []
    df["depth"] = df[["A", "C", "G", "T", "N", "d"]].sum(axis=1)
    consensus_depths = df[["A", "C", "G", "T", "N"]].max(axis=1)
    df["consensus"] = consensus_depths.divide(df.depth)
    for nt in ["A", "C", "G", "T", "N", "i", "d"]:
        df[[nt]] = df[[nt]].divide(df.depth, axis=0)
    vals = df[["A", "C", "G", "T", "i", "d"]].values
    with np.errstate(invalid="ignore", divide="ignore"):
        df["shannon"] = scipy.stats.entropy(vals.astype(np.float64), axis=1) if n_rows else []
    return df.round(3)


def plotly_clips(bam_path):
    """Plotly HTML of depth / clip / indel traces of the first contig (reference kindel/kindel.py:667-703)."""
    import plotly.graph_objs as go
    import plotly.offline as py

    aln = list(parse_bam(bam_path).items())[0][1]
    aligned_depth = np.asarray(aln.weights.cols).sum(axis=0).tolist()
    ins = np.asarray(aln.table[6]).tolist()
    x_axis = list(range(1, len(aligned_depth) + 1))
    traces = [
        go.Scattergl(x=x_axis, y=aligned_depth, mode="lines", name="Aligned depth"),
        go.Scattergl(x=x_axis, y=aln.clip_depth, mode="lines", name="Soft clip total depth"),
        go.Scattergl(x=x_axis, y=aln.clip_start_depth, mode="lines", name="Soft clip start depth"),
        go.Scattergl(x=x_axis, y=aln.clip_end_depth, mode="lines", name="Soft clip end depth"),
        go.Scattergl(x=x_axis, y=aln.clip_starts, mode="markers", name="Soft clip starts"),
        go.Scattergl(x=x_axis, y=aln.clip_ends, mode="markers", name="Soft clip ends"),
        go.Scattergl(x=x_axis, y=ins, mode="markers", name="Insertions"),
        go.Scattergl(x=x_axis, y=aln.deletions, mode="markers", name="Deletions"),
    ]
    fig = go.Figure(data=traces, layout=go.Layout(xaxis=dict(type="linear", autorange=True),
                                                  yaxis=dict(type="linear", autorange=True)))
    out_fn = os.path.splitext(os.path.split(bam_path)[1])[0]
    py.plot(fig, filename=out_fn + ".plot.html")
