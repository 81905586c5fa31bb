#!/bin/bash
# Round 2, GPU call 3: paneled triple kernel (panel counts), hash-bin variants, launch list of a cfg5 step.
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/c3_pytest.log
tail -5 gpurun_out/c3_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-per-config"
run() {  # tag env...
  tag=$1; shift
  env "$@" $B --workload $W > gpurun_out/c3_${W}_$tag.json 2> gpurun_out/c3_${W}_$tag.err
  echo "== $W $tag rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/c3_${W}_$tag.json')); print(round(d['ms_per_step'],3), 'ms', round(d['roofline']['kernel_ms'],3), 'kernel ms', d['phases_ms'])" 2>&1 | tail -1)"
}
W=cfg5
run auto X=1
for np in 2 3 4 6 8; do run np$np SPGEMM_B200_TRIPLE_PANELS=$np; done
run red SPGEMM_B200_TRIPLE_MODE=2
W=cfg3
run auto X=1
run np2 SPGEMM_B200_TRIPLE_PANELS=2
run red SPGEMM_B200_TRIPLE_MODE=2
W=cfg4r
for hb in 0 1 3; do run hb$hb SPGEMM_B200_HASH_BINS=$hb; done
W=cfg4
for hb in 0 1; do run hb$hb SPGEMM_B200_HASH_BINS=$hb; done
P="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-per-config --workload cfg5"
$P > gpurun_out/c3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_triple_panels -s 2 -c 1 -f -o gpurun_out/c3_prof_triple_cfg5 $P > gpurun_out/c3_ncu_full.log 2>&1
echo "full capture triple rc=$?"
$P > gpurun_out/c3_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c3_launches_cfg5.csv $P > gpurun_out/c3_ncu_launches.log 2>&1
echo "launch list rc=$?"
