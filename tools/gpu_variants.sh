#!/bin/bash
# A/B of prebuilt library variants (kindel_b200/_lib/variants/*.so, built here with different -D flags): each one is
# copied over the in-tree library and benched on the same workloads.  Lands in gpurun_out/<TAG>_variants.txt.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
TAG=${1:-r02v}
shift
WORKLOADS=${WORKLOADS:-"cfg4_5Mb_200x_simple cfg4_5Mb_200x"}
OUT=gpurun_out/${TAG}_variants.txt
: > $OUT
cp kindel_b200/_lib/libkindel_b200.so /tmp/keep.so
for V in "$@"; do
  cp kindel_b200/_lib/variants/$V.so kindel_b200/_lib/libkindel_b200.so
  for W in $WORKLOADS; do
    timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --workload $W > gpurun_out/${TAG}_${V}_${W}.json 2> gpurun_out/${TAG}_${V}_${W}.err
    python - "$V" "$W" gpurun_out/${TAG}_${V}_${W}.json >> $OUT <<'PY'
import json, sys
v, w, p = sys.argv[1:4]
try:
    d = json.loads(open(p).read().strip().splitlines()[-1])
    k = d.get("kernels_ms", {})
    print(f"{v:14s} {w:30s} step {d['ms_per_step']:.4f} ms  k0k1 {k.get('k0_k1_pileup', 0):.4f}  med {d['step_ms']['median']:.4f}  parity {d.get('parity')}  frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print(f"{v:14s} {w:30s} FAILED {e!r}")
PY
  done
done
cp /tmp/keep.so kindel_b200/_lib/libkindel_b200.so
cat $OUT
