#!/bin/bash
# Round 2 profiling pass (1 GPU): full GPU test suite, smoke, launch list of the default bench command, --set full captures
# of the dominant kernels.  Every ncu command is preceded by the same command without ncu (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/p_pytest.log
tail -4 gpurun_out/p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-per-config"
$B > gpurun_out/p_plain_default.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/p_launches_default.csv $B > gpurun_out/p_ncu_launches.log 2>&1
echo "launch list default rc=$?"
cap() {  # workload kernel-regex skip name [env]
  env ${5:-X=1} $B --workload $1 > gpurun_out/p_plain_$4.log 2>&1 &&
  env ${5:-X=1} ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/p_prof_$4 \
      $B --workload $1 > gpurun_out/p_ncu_full_$4.log 2>&1
  echo "full capture $4 rc=$?"
}
cap cfg5 k_triple_runs 2 triple_runs_cfg5
cap cfg3 k_triple_runs 2 triple_runs_cfg3
cap cfg5 k_triple_panels 2 triple_generic_cfg5 SPGEMM_B200_TRIPLE_GENERIC=1
cap cfg4r k_numeric_rank 2 numrank_cfg4r
cap cfg2 k_dense_rows_red 2 dense_cfg2
ls -la gpurun_out | grep p_prof
