#!/bin/bash
# One GPU call: the -m gpu suite, the shape timings and bench lines.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
TAG=${1:-r02c}
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -x > gpurun_out/${TAG}_gputests.log 2>&1
echo "gpu tests rc=$?" | tee gpurun_out/${TAG}_summary.txt
tail -5 gpurun_out/${TAG}_gputests.log | tee -a gpurun_out/${TAG}_summary.txt
timeout 600 python tools/r2_baselines.py > gpurun_out/${TAG}_baselines.log 2>&1
cat gpurun_out/${TAG}_baselines.log | tee -a gpurun_out/${TAG}_summary.txt
for w in cfg4_5Mb_200x cfg4_5Mb_200x_simple; do
  timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --workload $w > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  tail -c 400 gpurun_out/${TAG}_bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 1 --workload cfg4_5Mb_200x > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -c 300 gpurun_out/${TAG}_bench_reference.err
