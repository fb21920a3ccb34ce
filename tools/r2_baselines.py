"""Diagnostic (not a test): device time of K0+K1 (+K1g) and K2 on the BASELINE shapes that round 1 never timed.

    python tools/r2_baselines.py      (on the GPU box)
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kindel_b200 import engine, synth  # noqa: E402


def timeit(name, b, reps=10):
    db = engine.upload(b)
    table = engine.CountTable(b.n_slots, db.device)
    calls = torch.empty(b.n_slots, dtype=torch.uint8, device=db.device)
    for _ in range(3):
        engine.pileup(db, check=False, table=table)
        engine.vote(table.t, 1, out=calls)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    k1 = k2 = 0.0
    for _ in range(reps):
        ev[0].record()
        engine.pileup(db, check=False, table=table)
        ev[1].record()
        engine.vote(table.t, 1, out=calls)
        ev[2].record()
        torch.cuda.synchronize()
        k1 += ev[0].elapsed_time(ev[1]) / reps
        k2 += ev[1].elapsed_time(ev[2]) / reps
    print("%-34s reads %9d complex %8d bases %.3e  K1 %.4f ms  K2 %.4f ms  -> %.3e bases/s" % (
        name, b.n_reads, b.n_complex, b.aligned_bases, k1, k2, b.aligned_bases / ((k1 + k2) * 1e-3)), flush=True)


t0 = time.time()
timeit("cfg2 30kb x2000 simple", synth.simple_reads(2, [30_000], 2000))
timeit("cfg3 30kb x5000 complex", synth.complex_reads(3, 30_000, 5000))
timeit("cfg3-like 30kb x5000 simple", synth.simple_reads(3, [30_000], 5000))
timeit("cfg5 64x100kb x500 simple", synth.simple_reads(5, [100_000] * 64, 500))
timeit("cfg4 5Mb x200 simple", synth.simple_reads(4, [5_000_000], 200))
timeit("cfg4 5Mb x200, 1% complex", synth.mixed_reads(4, [5_000_000], 200, 0.01))
timeit("cfg4 5Mb x200, 20% complex", synth.mixed_reads(4, [5_000_000], 200, 0.20))
timeit("weak shard 1/8 of 5Mb, 1600x", synth.simple_reads(4, [5_000_000], 200, start_frac=(0.0, 0.125), read_seed=[4, 0]))
timeit("weak shard 1/2 of 5Mb, 400x", synth.simple_reads(4, [5_000_000], 200, start_frac=(0.0, 0.5), read_seed=[4, 0]))
print("total %.1f s" % (time.time() - t0))
