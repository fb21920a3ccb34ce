#!/bin/bash
# Profiles of one bench workload: launch list (per-kernel durations) + ncu --set full of the tile kernel.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
W=${1:-cfg4_5Mb_200x}
TAG=${2:-r02p}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_$W.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --workload $W > gpurun_out/${TAG}_launches_$W.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pileup_tile -s 4 -c 1 -f -o gpurun_out/${TAG}_k1_$W \
    python bench.py --steps 3 --warmup 3 --no-cpu --workload $W > gpurun_out/${TAG}_ncu_$W.log 2>&1
ls -la gpurun_out | grep ${TAG}
