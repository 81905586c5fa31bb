#!/bin/bash
# Round 2, call 15 (1 GPU): banded-Q kernel with the stream pipelined across ranges (SPGEMM_B200_TRIPLE_PIPE=2) against
# the per-range pipeline (=1): parity tests of both, then cfg5 / cfg3 bench lines of both.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_device_api.py -m gpu -q -x -k "triple or partition" 2>&1 | tail -4
for P in 1 2; do
  SPGEMM_B200_TRIPLE_PIPE=$P timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_reference_suite.py -m gpu -q -x -k "triple or cfg3 or cfg5" 2>&1 | tail -2
  for W in cfg5 cfg3; do
    SPGEMM_B200_TRIPLE_PIPE=$P timeout 600 python bench.py --steps 10 --warmup 3 --workload $W --no-per-config --no-cpu --no-e2e \
        > gpurun_out/c15_${W}_pipe$P.json 2> gpurun_out/c15_${W}_pipe$P.err
    echo "== $W pipe=$P rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c15_${W}_pipe$P.json')); print(round(d['ms_per_step'],3), 'ms/step', d['phases_ms'], 'frac', round(d['roofline']['frac'],4))" 2>&1 | tail -1)"
  done
done
