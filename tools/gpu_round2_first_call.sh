#!/bin/bash
# First GPU call of the next round (ONE GPU, ~12 minutes): validate the emulator-checked candidates of round 1 on
# the device, time them against the default kernel, and capture ncu reports of the two best for tools/ncu_regions.py.
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first_call.sh'
#
# Everything lands in gpurun_out/ (copy what is to be judged into profiles/).  Nothing here is a test: the
# driver's own `pytest -m gpu` covers the default path.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."

# 1. parity of the experimental paths through the C ABI (opt-in tests)
KDL_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q \
    -k "lean or ws2 or wide or derives_seq_off" > gpurun_out/r02_experimental_tests.log 2>&1
echo "experimental tests rc=$?" | tee -a gpurun_out/r02_summary.txt

# 2. K0 + tile-owner kernel: time and bit-for-bit parity of every variant, at cfg 4's depth and at twice that
timeout 300 python tools/k1f_sweep.py > gpurun_out/r02_sweep_200x.log 2>&1
SWEEP_DEPTH=400 timeout 400 python tools/k1f_sweep.py > gpurun_out/r02_sweep_400x.log 2>&1
cat gpurun_out/r02_sweep_200x.log gpurun_out/r02_sweep_400x.log | tee -a gpurun_out/r02_summary.txt

# 3. whole bench line per candidate (device-resident value + e2e), and the compact wire format
for v in tiled lean ws2; do
    KDL_K1F=$v timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu \
        > gpurun_out/r02_bench_$v.json 2> gpurun_out/r02_bench_$v.err
done
KDL_K1F=tiled timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu --derive-seq-off \
    > gpurun_out/r02_bench_tiled_derive_seq_off.json 2> gpurun_out/r02_bench_tiled_derive_seq_off.err

# 4. ncu --set full of the candidates' pileup kernel (one launch each, after warm-up), with source
for v in lean ws2; do
    KDL_K1F=$v timeout 400 ncu --set full --clock-control none --import-source on -k regex:pileup_ -s 3 -c 1 \
        -f -o gpurun_out/r02_prof_$v python bench.py --steps 2 --warmup 3 --no-cpu \
        > gpurun_out/r02_ncu_$v.log 2>&1
done
ls -la gpurun_out | tail -20

# 5. shapes round 1 never timed (K1g on cfg 3, deep piles, multi-contig)
timeout 400 python tools/r2_baselines.py > gpurun_out/r02_baselines.log 2>&1
cat gpurun_out/r02_baselines.log | tee -a gpurun_out/r02_summary.txt
