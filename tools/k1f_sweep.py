"""Diagnostic (not a test): K1f time on cfg 4 for a few launch configurations."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kindel_b200 import engine, synth  # noqa: E402

b = synth.simple_reads(4, [5_000_000], 200)
db = engine.upload(b)
table = engine.CountTable(b.n_slots, db.device)
for mult in os.environ.get("SWEEP", "tiled,ws").split(","):
    os.environ["KDL_K1F"] = mult
    for _ in range(3):
        engine.pileup(db, check=False, table=table)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        engine.pileup(db, check=False, table=table)
    ev[1].record()
    torch.cuda.synchronize()
    print("kernel", mult, "K0+K1f ms", ev[0].elapsed_time(ev[1]) / 10, flush=True)
