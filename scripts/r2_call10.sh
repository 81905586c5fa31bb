#!/bin/bash
set -u
mkdir -p gpurun_out
SPGEMM_B200_TRIPLE_3BLOCKS=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py -m gpu -q -x -k "triple or cfg3 or cfg5s or golden" 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-per-config"
run() {  # tag env...
  tag=$1; shift
  env "$@" $B --workload $W > gpurun_out/c10_${W}_$tag.json 2> gpurun_out/c10_${W}_$tag.err
  echo "== $W $tag rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/c10_${W}_$tag.json')); print(round(d['ms_per_step'],3), 'ms', round(d['roofline']['kernel_ms'],3), 'kernel ms', d['phases_ms'], 'frac', round(d['roofline']['frac'],3))" 2>&1 | tail -1)"
}
W=cfg5
run base X=1
run np5 SPGEMM_B200_TRIPLE_PANELS=5
for np in 6 7 8; do run three_np$np SPGEMM_B200_TRIPLE_3BLOCKS=1 SPGEMM_B200_TRIPLE_PANELS=$np; done
W=cfg3
run base X=1
run three SPGEMM_B200_TRIPLE_3BLOCKS=1
