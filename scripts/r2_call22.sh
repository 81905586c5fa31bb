#!/bin/bash
# Round 2, call 22 (1 GPU): banded-Q kernel with MLP independent load slots per warp (SPGEMM_B200_TRIPLE_MLP = 1 | 2 | 4).
set -u
mkdir -p gpurun_out
for M in 2 4; do
  SPGEMM_B200_TRIPLE_MLP=$M timeout 900 python -m pytest tests/test_gpu_device_api.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -m gpu -q -x -k "triple or cfg3 or cfg5" 2>&1 | tail -2
done
for M in 1 2 4; do
  for W in cfg5 cfg3; do
    SPGEMM_B200_TRIPLE_MLP=$M timeout 600 python bench.py --steps 10 --warmup 3 --workload $W --no-per-config --no-cpu --no-e2e > gpurun_out/c22_${W}_mlp$M.json 2> gpurun_out/c22_${W}_mlp$M.err
    echo "== $W mlp=$M rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c22_${W}_mlp$M.json')); print(round(d['ms_per_step'],3), 'ms/step', d['phases_ms'], 'frac', round(d['roofline']['frac'],4))" 2>&1 | tail -1)"
  done
done
