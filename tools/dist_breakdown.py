"""Diagnostic (not a test): per-stage CUDA-event breakdown of the fused multi-GPU step.

Default: every iteration starts from a barrier + device sync (stages timed in isolation).
BACK_TO_BACK=1: no barrier / sync between iterations, like bench.py's timed loop -- the mode in which
~0.1 ms per step beyond K1 + K2x + K2g is still unaccounted for."""
import ctypes as C
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kindel_b200 import _ffi, distributed as D, engine, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
full = synth.simple_reads(4, [5_000_000], 200)
shard = D.shard_batch(full, rank, world)
sc = D.ShardedConsensus(shard, dev, mode="fused")
lib = _ffi.load()
names = ["zero", "pileup", "signal", "vote", "wait"]
acc = {n: 0.0 for n in names}
wall = 0.0
evs = []
enq = []
for it in range(13):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    if not os.environ.get("BACK_TO_BACK"):
        dist.barrier()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev[0].record()
    ev[1].record()
    engine.pileup(sc.dbatch, check=False, table=sc.table, slot_range=sc.foot)
    ev[2].record()
    sc.epoch += 1
    lo, hi = sc.slices[sc.rank]
    st = int(torch.cuda.current_stream(dev).cuda_stream)
    lib.kdl_exchange_signal(C.byref(sc.xstruct), sc.epoch, st)
    ev[3].record()
    lib.kdl_exchange_vote(C.byref(sc.xstruct), sc.n_slots, 1, sc.epoch, st)
    ev[4].record()
    lib.kdl_exchange_wait(C.byref(sc.xstruct), sc.epoch, st)
    ev[5].record()
    t1 = time.perf_counter()
    enq.append((t1 - t0) * 1e3)
    if not os.environ.get("BACK_TO_BACK") or it == 12:
        torch.cuda.synchronize()
    t2 = time.perf_counter()
    evs.append(ev)
    if it >= 3 and not os.environ.get("BACK_TO_BACK"):
        for k, n in enumerate(names):
            acc[n] += ev[k].elapsed_time(ev[k + 1])
        wall += (t2 - t0) * 1e3
        if it == 12 and rank == 0:
            print("cpu enqueue ms", (t1 - t0) * 1e3)
gaps = 0.0
if os.environ.get("BACK_TO_BACK"):
    torch.cuda.synchronize()
    for ev in evs[3:]:
        for k, n in enumerate(names):
            acc[n] += ev[k].elapsed_time(ev[k + 1])
    wall = evs[3][0].elapsed_time(evs[-1][5])
    # device idle time BETWEEN iterations (last event of one to first event of the next): if this is where the
    # missing time sits, the loop is limited by the CPU enqueueing the launches, not by the exchange
    gaps = sum(evs[i][5].elapsed_time(evs[i + 1][0]) for i in range(3, len(evs) - 1))
    if rank == 0:
        print("cpu enqueue ms per step (back to back): %.4f" % (sum(enq[3:]) / len(enq[3:])), flush=True)
for r in range(world):
  dist.barrier()
  if rank == r:
    print("rank", rank, {n: round(v / 10, 4) for n, v in acc.items()}, "sum", round(sum(acc.values()) / 10, 4), "wall", round(wall / 10, 4), "gaps", round(gaps / 10, 4),
          "foot", sc.foot, "slots", sc.n_slots, flush=True)
dist.barrier()
sc.close()
dist.destroy_process_group()
