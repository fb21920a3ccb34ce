"""DEVELOPMENT AID for the build container (no GPU): run the BODIES of `-m gpu` tests against the CPU oracle.

    KDL_SHIM_ENGINE=1 python -m pytest tests/test_gpu_parity.py -k "digest or clip_heavy or cli" -q

With KDL_SHIM_ENGINE=1 and no CUDA device, conftest.py installs oracle-backed stand-ins for
`kindel_b200.engine.upload / pileup / vote / derive` and lifts the `gpu` skip, so that the Python of a new GPU
test (fixtures, host-side API calls, comparisons) is exercised here before GPU minutes are spent on it.  It
proves nothing about the kernels and is never active when a CUDA device is present; the product package has no
such path (kindel_b200.engine raises without CUDA)."""
import numpy as np
import torch

from kindel_b200 import engine
from oracle import coracle


def install():
    if torch.cuda.is_available():
        raise RuntimeError("the shim is for machines without a GPU only")
    cpu = torch.device("cpu")

    def upload(host, device=None, non_blocking=False):
        return engine.DeviceBatch(host=host, device=cpu, tensors={}, struct=None)

    def pileup(dbatch, counts=None, check=True, table=None, slot_range=None):
        try:
            c, e = coracle.pileup(dbatch.host)
        except (IndexError, KeyError):
            raise
        t = torch.from_numpy(c)
        if counts is not None:
            counts += t
            t = counts
        if table is not None:
            table.t.copy_(t)
            t = table.t
        return t, torch.from_numpy(np.ascontiguousarray(e))

    def vote(counts, min_depth=1, out=None):
        c = counts.numpy()
        if c.shape[0] == 7:  # the 7 vote columns only (consensus_sequence on caller-made tables)
            full = np.zeros((19, c.shape[1]), dtype=np.int32)
            full[0:7] = c
            c = full
        calls = torch.from_numpy(coracle.vote(c, min_depth))
        if out is not None:
            out.copy_(calls)
            return out
        return calls

    def derive(counts):
        return torch.from_numpy(coracle.derive(counts.numpy()))

    engine.require_cuda = lambda device=None: cpu
    engine.upload, engine.pileup, engine.vote, engine.derive = upload, pileup, vote, derive
