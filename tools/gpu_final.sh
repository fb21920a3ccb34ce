#!/bin/bash
# The round's closing 1-GPU call: -m gpu suite, smoke, every workload's bench line, launch list + ncu --set full of the
# default workload's tile kernel, the reference arm.  Everything lands in gpurun_out/ (copy what is judged to profiles/).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
TAG=${1:-r02z}
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/${TAG}_gputests.log 2>&1
echo "gpu tests rc=$?" | tee gpurun_out/${TAG}_summary.txt
tail -4 gpurun_out/${TAG}_gputests.log | tee -a gpurun_out/${TAG}_summary.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_summary.txt
for W in cfg4_5Mb_200x cfg4_5Mb_200x_simple cfg2_30kb_2000x cfg3_30kb_5000x cfg5_64x100kb_500x cfg4_5Mb_200x_30pct_complex; do
  extra="--no-cpu"; [ "$W" = "cfg4_5Mb_200x" ] && extra=""
  timeout 600 python bench.py --steps 20 --warmup 5 $extra --workload $W > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
  tail -c 300 gpurun_out/${TAG}_bench_$W.err
done
timeout 300 python bench.py --impl reference --steps 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for W in cfg4_5Mb_200x cfg3_30kb_5000x; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 48 --csv --log-file gpurun_out/${TAG}_launches_$W.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu --workload $W > gpurun_out/${TAG}_launches_$W.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pileup_tile -s 4 -c 1 -f -o gpurun_out/${TAG}_k1_cfg4 \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_cfg4.log 2>&1
ls -la gpurun_out | grep ${TAG} | wc -l
