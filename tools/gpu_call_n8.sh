#!/bin/bash
# N-GPU bench line of the default workload (weak + strong beside it).  gpurun --gpus N -- bash tools/gpu_call_n8.sh N
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
N=${1:-8}
TAG=${2:-r02n}
nvidia-smi -L > gpurun_out/${TAG}_n${N}_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29573 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_n${N}_cfg4.json 2> gpurun_out/${TAG}_n${N}_cfg4.err
echo "rc=$?"; tail -c 400 gpurun_out/${TAG}_n${N}_cfg4.err; head -c 600 gpurun_out/${TAG}_n${N}_cfg4.json
