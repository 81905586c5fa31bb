#!/bin/bash
# Round 2, call 23 (1 GPU): L2 prefetch of the rows of Q at the start of every (panel, row) item (SPGEMM_B200_TRIPLE_QPF).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_device_api.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -m gpu -q -x -k "triple or cfg3 or cfg5" 2>&1 | tail -2
for V in "0 1" "1 1" "1 2"; do
  set -- $V
  for W in cfg5 cfg3; do
    SPGEMM_B200_TRIPLE_QPF=$1 SPGEMM_B200_TRIPLE_MLP=$2 timeout 600 python bench.py --steps 10 --warmup 3 --workload $W --no-per-config --no-cpu --no-e2e > gpurun_out/c23_${W}_qpf$1_mlp$2.json 2> gpurun_out/c23_${W}_qpf$1_mlp$2.err
    echo "== $W qpf=$1 mlp=$2 rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c23_${W}_qpf$1_mlp$2.json')); print(round(d['ms_per_step'],3), 'ms/step', d['phases_ms'], 'frac', round(d['roofline']['frac'],4))" 2>&1 | tail -1)"
  done
done
