#!/bin/bash
# Round 2, GPU call 9: symbolic phase hands its bitmaps to the numeric rank kernel (on / off).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "not cfg5 and not triple" 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-per-config"
run() {  # tag env...
  tag=$1; shift
  env "$@" $B --workload $W > gpurun_out/c9_${W}_$tag.json 2> gpurun_out/c9_${W}_$tag.err
  echo "== $W $tag rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/c9_${W}_$tag.json')); print(round(d['ms_per_step'],3), 'ms', round(d['roofline']['kernel_ms'],3), 'kernel ms', d['phases_ms'], 'frac', round(d['roofline']['frac'],3))" 2>&1 | tail -1)"
}
for W in cfg4r cfg1 cfg4; do
  run keep X=1
  run nokeep SPGEMM_B200_BITMAP_KEEP_MB=0
done
W=cfg4
run keep16g SPGEMM_B200_BITMAP_KEEP_MB=16384
