#!/bin/bash
set -u
for D in 0 1 2 3; do
  SPGEMM_B200_TRIPLE_DIAG=$D timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg5 --no-per-config --no-cpu --no-e2e > gpurun_out/c21_diag$D.json 2> gpurun_out/c21_diag$D.err
  echo "== cfg5 diag=$D rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c21_diag$D.json')); print(round(d['ms_per_step'],3), 'ms/step', d['phases_ms'])" 2>&1 | tail -1)"
done
