"""torchrun worker for the multi-GPU parity test: both exchange modes == one-GPU result == oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kindel_b200 import distributed as D  # noqa: E402
from kindel_b200 import synth  # noqa: E402
from oracle import coracle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for kind in ("simple", "complex", "multi"):
        if kind == "simple":
            full = synth.simple_reads(31, [300_000], 120)
        elif kind == "complex":
            full = synth.complex_reads(32, 40_000, 300)
        else:
            full = synth.simple_reads(33, [50_000] * 6, 90)
        oc, _ = coracle.pileup(full)
        want = coracle.vote(oc, 2)
        shard = D.shard_batch(full, rank, world)
        for mode in ("fused", "peer", "allreduce"):
            sc = D.ShardedConsensus(shard, dev, mode=mode)
            for _ in range(3):  # repeated steps: the tables are re-zeroed and re-read safely
                calls = sc.step(2)
            torch.cuda.synchronize()
            got = calls.cpu().numpy()
            assert np.array_equal(got, want), (kind, mode, rank, int((got != want).sum()))
            if mode == "allreduce":
                assert np.array_equal(sc.counts[:7].cpu().numpy(), oc[:7])
            sc.close()
    dist.barrier()
    if rank == 0:
        print("dist parity ok: world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
