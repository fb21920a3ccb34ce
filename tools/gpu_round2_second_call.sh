#!/bin/bash
# Second GPU call of the next round (TWO GPUs, ~6 minutes, charged 2x): where does the multi-GPU step lose
# ~0.1 ms?   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_round2_second_call.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $RUN tools/dist_breakdown.py > gpurun_out/r02_dist_isolated.log 2>&1
BACK_TO_BACK=1 timeout 300 $RUN tools/dist_breakdown.py > gpurun_out/r02_dist_back_to_back.log 2>&1
timeout 300 $RUN bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_n2_weak.json 2> gpurun_out/r02_bench_n2_weak.err
timeout 300 $RUN bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --scaling strong > gpurun_out/r02_bench_n2_strong.json 2> gpurun_out/r02_bench_n2_strong.err
grep -h "rank\|cpu enqueue" gpurun_out/r02_dist_isolated.log gpurun_out/r02_dist_back_to_back.log
