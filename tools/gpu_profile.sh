#!/bin/bash
# Profiles of bench workloads: ncu --set full of the tile kernel (one launch after warm-up), then a launch list.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
TAG=${1:-r02p}
shift
for W in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pileup_tile -s 4 -c 1 -f -o gpurun_out/${TAG}_k1_$W \
      python bench.py --steps 3 --warmup 3 --no-cpu --workload $W > gpurun_out/${TAG}_ncu_$W.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches_$W.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu --workload $W > gpurun_out/${TAG}_launches_$W.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
