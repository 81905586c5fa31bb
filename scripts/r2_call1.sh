#!/bin/bash
# Round 2, GPU call 1: full GPU test suite, A/B of the triple-product kernels, launch list + full capture at cfg5.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
nproc
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-per-config"
run() {  # tag env... -- args
  tag=$1; shift
  env "$@" $B --workload $W > gpurun_out/c1_${W}_$tag.json 2> gpurun_out/c1_${W}_$tag.err
  echo "== $W $tag rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/c1_${W}_$tag.json')); print(round(d['ms_per_step'],3), 'ms', round(d['roofline']['kernel_ms'],3), 'kernel ms', d['phases_ms'])" 2>&1 | tail -1)"
}
for W in cfg5 cfg3; do
  run window X=1
  run red SPGEMM_B200_TRIPLE_MODE=2
  run g8 SPGEMM_B200_TRIPLE_GROUP=8
  run g32 SPGEMM_B200_TRIPLE_GROUP=32
done
W=cfg5
run win13k SPGEMM_B200_TRIPLE_WIN=13000
run win20k SPGEMM_B200_TRIPLE_WIN=20000
# launch list and one full capture of the window kernel (plain run first)
P="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-per-config --workload cfg5"
$P > gpurun_out/c1_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c1_launches_cfg5.csv $P > gpurun_out/c1_ncu_launches.log 2>&1
echo "launch list rc=$?"
$P > gpurun_out/c1_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_triple_window -s 2 -c 1 -f -o gpurun_out/c1_prof_triple_cfg5 $P > gpurun_out/c1_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -30
