#!/bin/bash
# One multi-GPU call (gpurun --gpus N): the >= 2-GPU parity tests, then bench lines at N.  Lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
N=${1:-2}
TAG=${2:-r02m}
nvidia-smi -L | tee gpurun_out/${TAG}_n${N}_summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "multi_gpu or two_gpus" > gpurun_out/${TAG}_n${N}_tests.log 2>&1
echo "multi-gpu tests rc=$?" | tee -a gpurun_out/${TAG}_n${N}_summary.txt
tail -15 gpurun_out/${TAG}_n${N}_tests.log | tee -a gpurun_out/${TAG}_n${N}_summary.txt
run() {  # name, extra args
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
      bench.py --gpus $N --steps 20 --warmup 5 $2 > gpurun_out/${TAG}_n${N}_$1.json 2> gpurun_out/${TAG}_n${N}_$1.err
  echo "$1 rc=$?" | tee -a gpurun_out/${TAG}_n${N}_summary.txt
  tail -c 300 gpurun_out/${TAG}_n${N}_$1.err
}
run cfg4 ""
[ "${SKIP_ALLREDUCE:-0}" = 1 ] || run cfg4_allreduce "--exchange allreduce --no-strong"
run cfg5 "--workload cfg5_64x100kb_500x"
ls -la gpurun_out | grep ${TAG}_n${N}
