#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py -m gpu -q -x -k "triple or cfg3 or cfg5s or golden or seeded" 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-per-config"
run() {  # tag env...
  tag=$1; shift
  env "$@" $B --workload $W > gpurun_out/c11_${W}_$tag.json 2> gpurun_out/c11_${W}_$tag.err
  echo "== $W $tag rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/c11_${W}_$tag.json')); print(round(d['ms_per_step'],3), 'ms', round(d['roofline']['kernel_ms'],3), 'kernel ms', 'frac', round(d['roofline']['frac'],3))" 2>&1 | tail -1)"
}
W=cfg5
for pct in 0 30 45 55 65 75; do run red$pct SPGEMM_B200_TRIPLE_RED_PCT=$pct; done
for np in 2 3 5; do run red55_np$np SPGEMM_B200_TRIPLE_RED_PCT=55 SPGEMM_B200_TRIPLE_PANELS=$np; done
run generic55 SPGEMM_B200_TRIPLE_GENERIC=1
run generic0 SPGEMM_B200_TRIPLE_GENERIC=1 SPGEMM_B200_TRIPLE_RED_PCT=0
W=cfg3
for pct in 0 40 55 70; do run red$pct SPGEMM_B200_TRIPLE_RED_PCT=$pct; done
