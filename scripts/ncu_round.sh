#!/bin/bash
# Profiling pass for one round (run under gpurun, 1 GPU): launch lists + one --set full capture per hot kernel.
# Every ncu command is preceded by the same command without ncu (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
for w in cfg2 cfg1 cfg3 cfg4r; do
  $B --workload $w > gpurun_out/plain_$w.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$w.csv \
      $B --workload $w > gpurun_out/ncu_launches_$w.log 2>&1
  echo "launch list $w rc=$?"
done
cap() {  # workload kernel-regex skip name
  $B --workload $1 > gpurun_out/plain2_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$4 \
      $B --workload $1 > gpurun_out/ncu_full_$4.log 2>&1
  echo "full capture $4 rc=$?"
}
cap cfg2 k_dense_rows_red 3 dense_cfg2
cap cfg3 k_triple_rows_red 3 triple_cfg3
cap cfg4r k_numeric_rank 3 numrank_cfg4r
cap cfg4r k_symbolic_bitmap 3 symbitmap_cfg4r
cap cfg1 "k_numeric_warp<256" 3 numwarp256_cfg1
ls -la gpurun_out
