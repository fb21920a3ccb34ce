"""Diagnostic (not a test): K0 + tile-owner kernel time on cfg 4 for the kernel variants, with a parity check.

    SWEEP=tiled,lean,ws2,wide,ws python tools/k1f_sweep.py        (on the GPU box)

Each variant (KDL_K1F=...) is run on the same batch into a reused CountTable; the first variant's weight
columns are the yardstick the others must equal bit for bit (and the first is the GPU-validated default)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kindel_b200 import engine, synth  # noqa: E402

b = synth.simple_reads(4, [5_000_000], int(os.environ.get("SWEEP_DEPTH", "200")))
db = engine.upload(b)
table = engine.CountTable(b.n_slots, db.device)
want = None
for variant in os.environ.get("SWEEP", "tiled,lean,ws2,wide,ws").split(","):
    os.environ["KDL_K1F"] = variant
    for _ in range(3):
        engine.pileup(db, check=False, table=table)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        engine.pileup(db, check=False, table=table)
    ev[1].record()
    torch.cuda.synchronize()
    got = table.t[0:5].clone()
    if want is None:
        want, verdict = got, "yardstick"
    else:
        verdict = "EQUAL" if torch.equal(got, want) else "DIFFERENT (%d slots)" % int((got != want).any(dim=0).sum())
    print("kernel", variant, "K0+K1 ms %.4f" % (ev[0].elapsed_time(ev[1]) / 10), verdict, flush=True)
