#!/bin/bash
# Round 2, GPU call 4: lean runs kernel for banded Q vs the generic paneled kernel; panel counts.
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py -m gpu -q -x -k "triple or cfg or golden or seeded or multi" 2>&1 | tail -3
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-per-config"
run() {  # tag env...
  tag=$1; shift
  env "$@" $B --workload $W > gpurun_out/c7_${W}_$tag.json 2> gpurun_out/c7_${W}_$tag.err
  echo "== $W $tag rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/c7_${W}_$tag.json')); print(round(d['ms_per_step'],3), 'ms', round(d['roofline']['kernel_ms'],3), 'kernel ms', d['phases_ms'])" 2>&1 | tail -1)"
}
W=cfg5
run auto X=1
for np in 3 4 5 6; do run np$np SPGEMM_B200_TRIPLE_PANELS=$np; done
run generic SPGEMM_B200_TRIPLE_GENERIC=1
W=cfg3
run auto X=1
run np2 SPGEMM_B200_TRIPLE_PANELS=2
run generic SPGEMM_B200_TRIPLE_GENERIC=1
P="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-per-config --workload cfg5"
$P > gpurun_out/c7_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_triple_runs -s 2 -c 1 -f -o gpurun_out/c7_prof_triple_cfg5 $P > gpurun_out/c7_ncu_full.log 2>&1
echo "full capture triple rc=$?"
