#!/bin/bash
# One profiling pass for profiles/rNN (run under gpurun, 1 GPU).  Every ncu command is preceded by the same
# command without ncu, in the same call, with && in between (B200_PROFILING.md).
set -u
OUT=gpurun_out
mkdir -p $OUT
# 1. the default bench command, as the driver runs it: launch list with per-launch device time
D="python bench.py --steps 2 --warmup 3"
$D > $OUT/bench_default_plain.json 2> $OUT/bench_default_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_default.csv \
    $D > $OUT/bench_default_under_ncu.json 2> $OUT/bench_default_under_ncu.err
echo "default launch list rc=$?"
# 2. launch lists of the other workloads (device-resident leg only)
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
for w in cfg1 cfg3 cfg4r; do
  $B --workload $w > $OUT/plain_$w.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$w.csv \
      $B --workload $w > $OUT/ncu_launches_$w.log 2>&1
  echo "launch list $w rc=$?"
done
# 3. one --set full capture per dominant kernel
cap() {  # workload kernel-regex skip name
  $B --workload $1 > $OUT/plain2_$4.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $OUT/prof_$4 \
      $B --workload $1 > $OUT/ncu_full_$4.log 2>&1
  echo "full capture $4 rc=$?"
}
cap cfg2 k_dense_rows_red 3 dense_cfg2
cap cfg3 k_triple_rows_red 3 triple_cfg3
cap cfg4r k_numeric_rank 3 numrank_cfg4r
cap cfg4r k_symbolic_bitmap 3 symbitmap_cfg4r
cap cfg1 "k_numeric_warp<256" 3 numwarp256_cfg1
