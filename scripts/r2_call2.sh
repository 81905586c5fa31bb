#!/bin/bash
# Round 2, GPU call 2: flat-stream triple kernel (sync / ticket), block-hash numeric bins on/off.
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/c2_pytest.log
tail -5 gpurun_out/c2_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-per-config"
run() {  # tag env...
  tag=$1; shift
  env "$@" $B --workload $W > gpurun_out/c2_${W}_$tag.json 2> gpurun_out/c2_${W}_$tag.err
  echo "== $W $tag rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/c2_${W}_$tag.json')); print(round(d['ms_per_step'],3), 'ms', round(d['roofline']['kernel_ms'],3), 'kernel ms', d['phases_ms'])" 2>&1 | tail -1)"
}
for W in cfg5 cfg3; do
  run sync SPGEMM_B200_TRIPLE_SYNC=1
  run ticket SPGEMM_B200_TRIPLE_SYNC=0
done
for W in cfg4r cfg1 cfg4; do
  run hash X=1
  run nohash SPGEMM_B200_NO_HASH_BINS=1
done
P="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-per-config --workload cfg5"
$P > gpurun_out/c2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_triple_window -s 2 -c 1 -f -o gpurun_out/c2_prof_triple_cfg5 $P > gpurun_out/c2_ncu_full.log 2>&1
echo "full capture triple rc=$?"
P="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-per-config --workload cfg4r"
$P > gpurun_out/c2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_numeric_hash -s 3 -c 2 -f -o gpurun_out/c2_prof_hash_cfg4r $P > gpurun_out/c2_ncu_full2.log 2>&1
echo "full capture hash rc=$?"
$P > gpurun_out/c2_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c2_launches_cfg4r.csv $P > gpurun_out/c2_ncu_launches.log 2>&1
echo "launch list rc=$?"
