#!/usr/bin/env python
"""bench.py -- aligned bases / second through pileup + consensus (BASELINE.json's metric).

    python bench.py --gpus 1 --steps 10 --warmup 3            # this engine, 1 GPU
    torchrun ... bench.py --gpus N ...                        # read-sharded over N GPUs (NCCL)
    python bench.py --impl reference ...                      # the reference's CPU path, same metric

A "step" is one pass of the hot path over one batch: zero the count table, K1 pileup over every read,
(N > 1: sum the vote columns across ranks), K2 vote over every position.

Workload (config.workload): BASELINE.json configs[3], the one the north star's targets are quoted
on -- synthetic 5 Mb contig, 200x, 150 bp `150M` reads, coordinate-sorted, 1 % substitutions
(6.67 M reads, 10^9 aligned bases).  For N > 1 every rank gets that full per-GPU workload on its own 1/N
slice of the coordinate range (weak scaling: a coordinate-sorted, N x deeper alignment cut into N contiguous read
blocks); `--scaling strong` cuts the N = 1 data set N ways instead.

Numbers on the JSON line:
  value      whole-job aligned bases/s with the flattened reads already resident in HBM
             (CUDA events on the launch stream, max over ranks).
  e2e        the same metric through the C-ABI host-buffer call kdl_ctx_consensus: pinned HOST buffers
             in, H2D + kernels + D2H of the call bytes inside the timed region.
  roofline   dominant kernel (K1) against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the reference's algorithm on this box's host CPU, bounded sample (see --impl reference).
Input (607 MB) is larger than L2 (126 MB), so no explicit L2 flush is needed between iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "aligned bases/sec through pileup+consensus"
UNIT = "aligned_bases/s"

WORKLOADS = {
    # name: (contig lengths, depth)
    "cfg4_5Mb_200x": ([5_000_000], 200),
    "cfg2_30kb_2000x": ([30_000], 2000),
    "cfg5_64x100kb_500x": ([100_000] * 64, 500),
    "tiny": ([200_000], 50),
}


def make_workload(name, rank=0, world=1, scaling="weak"):
    """This rank's reads and the whole job's aligned bases.

    weak   (default): every rank gets the full per-GPU workload -- `depth` x the contig lengths worth of
           reads -- placed on its own 1/N slice of the coordinate range (a coordinate-sorted N x deeper
           BAM cut into N contiguous blocks): per-GPU work is fixed as N grows.
    strong: the N = 1 data set cut into N contiguous read blocks: total work fixed."""
    from kindel_b200 import distributed, synth

    lens, depth = WORKLOADS[name]
    if world == 1:
        full = synth.simple_reads(4, lens, depth)
        return full, full.aligned_bases
    if scaling == "strong":
        full = synth.simple_reads(4, lens, depth)
        return distributed.shard_batch(full, rank, world), full.aligned_bases
    shard = synth.simple_reads(4, lens, depth, start_frac=(rank / world, (rank + 1) / world), read_seed=[4, rank])
    return shard, shard.aligned_bases * world  # every rank holds the same number of equally long reads


def algorithmic_bytes(batch):
    """per read ceil(l_seq/2) + 4*n_cigar + 12 read-side bytes (K1);
    per position 28 B read + 1 B written by the vote (K2)."""
    lseq = batch.seq_len.astype(np.int64)
    n_cig = np.diff(batch.cig_off.astype(np.int64))
    k1 = int(((lseq + 1) // 2).sum() + 4 * n_cig.sum() + 12 * batch.n_reads)
    k2 = int(batch.n_slots) * 29
    return k1, k2


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` capture of this same command (profiles/); None for configurations not captured."""
    if workload != "cfg4_5Mb_200x" or world != 1:
        return None  # the committed capture is of the default workload
    try:
        with open(os.path.join(ROOT, "profiles", "r02_k1_traffic.json")) as fh:
            d = json.load(fh)
        return int(d["dram_bytes_read"]) + int(d["dram_bytes_write"])
    except Exception:
        return None


# ------------------------------------------------------------------------------ CPU baselines
def cpu_port_sample(batch, seconds_target=12.0):
    """Reference-shaped Python port (oracle/py_oracle.py) on a bounded window of the workload:
    the reads whose start lies in the first `window` positions, against a contig truncated there."""
    from oracle import py_oracle

    read_len = 150
    L0 = int(batch.contig_len[0])
    window = min(L0, 250_000)
    hi = int(np.searchsorted(batch.ref_start[: int(batch.contig_read_off[1])], window - read_len, side="right"))
    recs = py_oracle.records_of(batch, 0, hi)
    bases = sum(n for r in recs for n, op in r.cigars if op in "M=X")
    t0 = time.perf_counter()
    p = py_oracle.pileup(window, recs)
    py_oracle.vote(p, 1)
    dt = time.perf_counter() - t0
    return {"value": bases / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle/py_oracle.py (reference-shaped CPython loop): %d reads / %d aligned bases over the "
                      "first %d positions of the workload, pileup+post-pass+vote, %.1f s" % (len(recs), bases, window, dt)}


def cpu_native_sample(batch):
    """The C restatement (oracle/kindel_oracle.c), 1 thread, whole shard: what a compiled CPU loop does."""
    from oracle import coracle

    t0 = time.perf_counter()
    counts, _ = coracle.pileup(batch)
    coracle.vote(counts, 1)
    dt = time.perf_counter() - t0
    return {"value": batch.aligned_bases / dt, "unit": UNIT, "cores": 1, "kind": "port-native",
            "sample": "oracle/kindel_oracle.c single thread, full workload, %.2f s" % dt}


def host_side_timings(batch):
    """The host work that surrounds the timed spans (reported separately; it stays on the host in
    both paths): BAM inflate + C++ gather + flatten of a 10^6-read slice of the workload written as a real
    BGZF-compressed BAM, on this box's cores."""
    import tempfile

    from kindel_b200 import bamio, synth

    n = min(batch.n_reads, 1_000_000)
    words = int(batch.seq_off[1] - batch.seq_off[0]) if batch.n_reads > 1 else 19
    sub = bamio.finalize(batch.contig_names, batch.contig_len, np.array([0, n]), batch.ref_start[:n],
                         np.arange(n, dtype=np.int64) * words, batch.l_seq[:n], np.arange(n + 1),
                         np.full(n, int(batch.l_seq[0]) << 4), batch.seq4[: n * words], n_records=n)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "slice.bam")
        synth.write_simple_bam(path, sub)
        size = os.path.getsize(path)
        t0 = time.perf_counter()
        back = bamio.read_bam(path)
        dt = time.perf_counter() - t0
    assert back.n_reads == n
    return {"bam_decode_flatten_reads_per_s": n / dt, "bam_decode_flatten_aligned_bases_per_s": back.aligned_bases / dt,
            "sample": "%d reads, %.0f MB BGZF BAM, inflate (zlib, thread pool) + C++ gather + numpy flatten: %.2f s"
                      % (n, size / 1e6, dt), "cores": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    batch, total = make_workload(args.workload)
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        vals.append(cpu_port_sample(batch))
    best = max(vals, key=lambda v: v["value"])
    v = statistics.median(x["value"] for x in vals)
    native = cpu_native_sample(batch)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": 0, "ms_per_step": None, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": args.workload, "note": "reference is single-threaded CPython (kindel/kindel.py:1-14)"},
        "cpu_baseline": dict(best, value=v),
        "cpu_native_port": native,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"nproc": os.cpu_count()},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------- GPU arm
def run_native(args):
    import torch
    import torch.distributed as dist

    from kindel_b200 import _ffi, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _ffi.load()

    batch, total_bases = make_workload(args.workload, rank, world, args.scaling)
    n_slots = batch.n_slots
    k1_bytes, k2_bytes = algorithmic_bytes(batch)
    if world == 1:
        db = engine.upload(batch, dev)
        table = engine.CountTable(n_slots, dev)
        counts = table.t
        calls_buf = torch.empty(n_slots, dtype=torch.uint8, device=dev)

        def step(timers=None):
            # a fresh pileup into a reused table: nothing is memset, K1f overwrites the weight columns
            if timers:
                timers[0].record()
            engine.pileup(db, check=False, table=table)
            if timers:
                timers[1].record()
            out = engine.vote(counts, 1, out=calls_buf)
            if timers and len(timers) > 2:
                timers[2].record()
            return out
    else:
        from kindel_b200 import distributed

        sc = distributed.ShardedConsensus(batch, dev, mode=args.exchange)
        db, counts = sc.dbatch, sc.counts

        def step(timers=None):
            return sc.step(1, timers)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    k1_ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.kdl_launch_count()
    torch.cuda.synchronize()
    ev0.record()
    for i in range(args.steps):
        calls = step(k1_ev[i])
    ev1.record()
    torch.cuda.synchronize()
    launches = lib.kdl_launch_count() - launches0
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    k1_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in k1_ev)
    k2_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in k1_ev) if world == 1 else None
    t = torch.tensor([ms_total, k1_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, k1_ms_max = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    value = total_bases / (ms_per_step * 1e-3)

    # ---- e2e through the C-ABI host-buffer entry point, pinned host memory ----------------------
    e2e = None
    if rank == 0 or world > 1:
        ctx = engine.HostContext(local)
        pinned = {}
        for f in engine._FIELDS:
            a = np.ascontiguousarray(getattr(batch, f))
            tpin = torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).pin_memory() if a.size else None
            pinned[f] = tpin
        ptr = {f: (int(tp.data_ptr()) if tp is not None else None) for f, tp in pinned.items()}
        struct = engine.make_struct(batch, ptr)
        calls_host = torch.empty(n_slots, dtype=torch.uint8).pin_memory()
        calls_np = calls_host.numpy()
        e2e_ms = []
        for i in range(max(2, args.warmup - 1) + args.steps):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            ctx.consensus(batch, 1, calls_out=calls_np, struct=struct)
            wall = (time.perf_counter() - t0) * 1e3
            tm = ctx.last_timing()
            e2e_ms.append((tm["h2d_ms"] + tm["kernel_ms"] + tm["d2h_ms"], wall, tm))
        e2e_ms = e2e_ms[-args.steps:]
        dev_ms = statistics.mean(x[0] for x in e2e_ms)
        wall_ms = statistics.mean(x[1] for x in e2e_ms)
        tt = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        # every rank copies its own shard in and its calls out; with N > 1 this leg runs the shards
        # concurrently but does not reduce across ranks (the reduction is in `value`'s step)
        e2e = {"value": total_bases / (float(tt[0]) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": batch.input_bytes(),
               "d2h_bytes_per_step": int(n_slots) + 16,
               "ms_per_step": float(tt[0]), "wall_ms_per_step": float(tt[1]),
               "breakdown_ms": {k: statistics.mean(x[2][k] for x in e2e_ms) for k in ("h2d_ms", "kernel_ms", "d2h_ms")},
               "api": "kdl_ctx_consensus (include/kindel_b200.h), pinned host buffers",
               "note": None if world == 1 else "every rank pushes its own shard through the host-buffer call "
                       "concurrently (N PCIe links); the cross-rank exchange is part of `value`, not of this leg"}
        ctx.close()

    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = k1_bytes / (k1_ms_max * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": args.workload if world == 1 or args.scaling == "strong" else
                       "%s per GPU (%dx the depth in total, cut into %d coordinate blocks)" % (args.workload, world, world),
                       "reads_per_rank": int(batch.n_reads),
                       "aligned_bases_total": int(total_bases),
                       "sharding": "contiguous blocks of the coordinate-sorted reads, one per rank" if world > 1 else "none",
                       "reduction": ("none" if world == 1 else
                                     "K2x: flags + reduce + vote + call scatter in one kernel over CUDA-IPC peer memory "
                                     "(NVLink), footprint-clipped; no NCCL on the data path" if args.exchange == "fused" else
                                     "K2p: vote over peer tables (NVLink), NCCL barrier + all_gather of call bytes"
                                     if args.exchange == "peer" else
                                     "NCCL all_reduce(int32 sum) of the 7 vote columns, vote replicated"),
                       "l2_policy": "inputs (%.0f MB) larger than L2 (126 MB); no flush" % (batch.input_bytes() / 1e6)},
            "roofline": {"bound": "hbm", "kernel": "K0 tile index + K1 tile-owner pileup",
                         "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.workload, world),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": k1_bytes, "kernel_ms": k1_ms_max},
            "kernels_ms": {"k0_k1_pileup": k1_ms_max, "k2_vote": k2_ms,
                           "k2_vote_gbs": (k2_bytes / (k2_ms * 1e-3) / 1e9) if k2_ms else None},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_port_sample(batch)
            line["cpu_native_port"] = cpu_native_sample(batch)
            if len(batch.contig_names) == 1 and batch.n_complex == 0:
                line["host"] = host_side_timings(batch)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cfg4_5Mb_200x")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="N > 1: weak = the full per-GPU workload on every rank (N x deeper in total); "
                         "strong = the N = 1 data set cut N ways")
    ap.add_argument("--exchange", choices=["fused", "peer", "allreduce"], default="fused",
                    help="N > 1: fused = flags + reduce + vote + call scatter over NVLink peer memory (no NCCL on "
                         "the data path); peer = same kernel with NCCL barrier/all_gather; allreduce = NCCL "
                         "all_reduce of the vote columns, vote replicated")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    raise SystemExit(main())
