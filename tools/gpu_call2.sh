#!/bin/bash
# Launch lists + one ncu capture + bench lines of the non-default workloads.  Lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
TAG=${1:-r02h}
for W in cfg4_5Mb_200x cfg3_30kb_5000x cfg4_5Mb_200x_30pct_complex; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 48 --csv --log-file gpurun_out/${TAG}_launches_$W.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu --workload $W > gpurun_out/${TAG}_launches_$W.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pileup_tile -s 4 -c 1 -f -o gpurun_out/${TAG}_k1_cfg3 \
    python bench.py --steps 3 --warmup 3 --no-cpu --workload cfg3_30kb_5000x > gpurun_out/${TAG}_ncu_cfg3.log 2>&1
for W in cfg4_5Mb_200x cfg2_30kb_2000x cfg3_30kb_5000x cfg5_64x100kb_500x cfg4_5Mb_200x_30pct_complex cfg4_5Mb_200x_simple; do
  timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --workload $W > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
  tail -c 300 gpurun_out/${TAG}_bench_$W.err
done
ls -la gpurun_out | grep ${TAG}
