#!/bin/bash
# Copy what tools/gpu_final.sh <TAG> left in gpurun_out/ into profiles/ under the round's names, and derive the text
# summaries from the ncu report.  profiles/ is what is judged; gpurun_out/ is scratch.
set -eu
cd "$(dirname "$0")/.."
TAG=${1:-r02z}
G=gpurun_out
P=profiles
for W in cfg4_5Mb_200x cfg4_5Mb_200x_simple cfg2_30kb_2000x cfg3_30kb_5000x cfg5_64x100kb_500x cfg4_5Mb_200x_30pct_complex; do
  [ -s $G/${TAG}_bench_$W.json ] && cp $G/${TAG}_bench_$W.json $P/r02_bench_n1_$W.json
done
[ -s $G/${TAG}_bench_reference.json ] && cp $G/${TAG}_bench_reference.json $P/r02_bench_reference_arm.json
for W in cfg4_5Mb_200x cfg3_30kb_5000x; do
  [ -s $G/${TAG}_launches_$W.csv ] && grep -v '^==' $G/${TAG}_launches_$W.csv > $P/r02_launches_$W.csv
done
cp $G/${TAG}_summary.txt $P/r02_gpu_tests_summary.txt
if [ -s $G/${TAG}_k1_cfg4.ncu-rep ]; then
  python $P/summarize_ncu.py $G/${TAG}_k1_cfg4.ncu-rep > $P/r02_k1_cfg4_ncu_summary.txt
  python tools/ncu_lines.py $G/${TAG}_k1_cfg4.ncu-rep 60 > $P/r02_k1_cfg4_ncu_lines.txt
fi
python tools/results_table.py > /tmp/results_table.md
echo "copied; results table in /tmp/results_table.md"
