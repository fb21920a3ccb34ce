"""Host-side logic of the read-sharded path on CPU: 2-process gloo group.
Sharding, footprints and slot ownership are numpy; the per-shard tables come from the CPU oracle
(test infrastructure) and are summed with a real `torch.distributed` all_reduce over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kindel_b200 import distributed as D
from kindel_b200 import synth
from oracle import coracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = synth.complex_reads(5, 6000, 120) if kind == "complex" else synth.simple_reads(6, [4000, 3000], 80)
        shard = D.shard_batch(full, rank, world)
        counts, _ = coracle.pileup(shard)
        lo, hi = D.footprint(shard)
        # the table is zero outside the shard's footprint (what lets K2p skip peers)
        assert not counts[:, :lo].any() and not counts[:, hi:].any()
        t = torch.from_numpy(counts[:7].copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        want, _ = coracle.pileup(full)
        assert np.array_equal(t.numpy(), want[:7])
        # emulate the fused reduce + vote of this rank's slot slice from footprint-clipped tables
        feet = [None] * world
        dist.all_gather_object(feet, (lo, hi))
        tabs = [None] * world
        dist.all_gather_object(tabs, counts[:7])
        a, b = D.owner_slices(full.n_slots, world)[rank]
        acc = np.zeros((7, full.n_slots), dtype=np.int64)
        for (flo, fhi), tab in zip(feet, tabs):
            acc[:, flo:fhi] += tab[:, flo:fhi]
        assert np.array_equal(acc[:, a:b], want[:7, a:b])
        calls = coracle.vote(acc.astype(np.int32), 1)[a:b]
        assert np.array_equal(calls[:-1], coracle.vote(want, 1)[a:b][:-1])  # last slot needs depth_next of b
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["simple", "complex"])
def test_two_rank_sharded_reduction_gloo(kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) == "ok"


def test_shards_partition_the_reads():
    full = synth.complex_reads(9, 5000, 60)
    parts = [D.shard_batch(full, r, 3) for r in range(3)]
    assert sum(p.n_reads for p in parts) == full.n_reads
    assert sum(p.aligned_bases for p in parts) == full.aligned_bases
    assert sum(p.n_events for p in parts) == full.n_events
    total = sum(coracle.pileup(p)[0].astype(np.int64) for p in parts)
    assert np.array_equal(total, coracle.pileup(full)[0])
    for w in (1, 2, 3, 8):
        sl = D.owner_slices(full.n_slots, w)
        assert sl[0][0] == 0 and sl[-1][1] == full.n_slots
        assert all(a[1] == b[0] for a, b in zip(sl, sl[1:])) and all(lo % 512 == 0 for lo, _ in sl)


def test_plans_partitions_and_event_merge():
    """Host logic of the public multi-GPU entry: plan choice, whole-contig partition, slices cut along footprints,
    and the merge of per-shard insertion events back into the single-GPU order."""
    multi = synth.mixed_reads(11, [3000, 5000, 2500, 4000, 3500, 4500], 25, 0.3)
    single = synth.complex_reads(12, 6000, 80)
    assert D.choose_plan(multi, 2) == "contigs" and D.choose_plan(multi, 8) == "reads" and D.choose_plan(single, 2) == "reads"
    want_c, want_e = coracle.pileup(multi)
    for world, plan in ((2, "contigs"), (3, "contigs"), (3, "reads")):
        idx = [D.shard_indices(multi, r, world, plan) for r in range(world)]
        assert sorted(np.concatenate(idx).tolist()) == list(range(multi.n_reads))
        shards = [D.select_reads(multi, i) for i in idx]
        tabs, evs = zip(*(coracle.pileup(s) for s in shards))
        np.testing.assert_array_equal(sum(t.astype(np.int64) for t in tabs), want_c)
        np.testing.assert_array_equal(D.merge_events(evs, idx), want_e)
        feet = [D.footprint(s) for s in shards]
        for (lo, hi), t in zip(feet, tabs):
            assert not t[:, :lo].any() and not t[:, hi:].any()
        sl = D.footprint_slices(feet, multi.n_slots)
        assert sl[0][0] == 0 and sl[-1][1] == multi.n_slots and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        if plan == "contigs":  # nothing is shared: every slice's slots are covered by its own table only
            for r, (a, b) in enumerate(sl):
                others = sum(tabs[p][:, a:b].astype(np.int64) for p in range(world) if p != r)
                assert not np.any(others)
    assert D.footprint_slices([(0, 0), (0, 0)], 4096) == D.owner_slices(4096, 2)


def test_batch_save_load_roundtrip(tmp_path):
    from kindel_b200 import bamio

    b = synth.complex_reads(13, 4000, 30)
    bamio.save_batch(str(tmp_path / "b"), b)
    back = bamio.load_batch(str(tmp_path / "b"))
    for f in bamio._SAVE_FIELDS:
        np.testing.assert_array_equal(getattr(back, f), getattr(b, f), err_msg=f)
    assert back.contig_names == b.contig_names and back.n_events == b.n_events and back.reach_right == b.reach_right
    np.testing.assert_array_equal(coracle.pileup(back)[0], coracle.pileup(b)[0])
