#!/bin/bash
# Round 2, call 20 (1 GPU): banded-Q kernel with the per-entry metadata in one int4 carried in registers.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_device_api.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py tests/test_reference_suite.py -m gpu -q -x -k "triple or cfg3 or cfg5 or multi" 2>&1 | tail -3
for W in cfg5 cfg3; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $W --no-per-config --no-cpu --no-e2e > gpurun_out/c20_$W.json 2> gpurun_out/c20_$W.err
  echo "== $W rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c20_$W.json')); print(round(d['ms_per_step'],3), 'ms/step', d['phases_ms'], 'frac', round(d['roofline']['frac'],4))" 2>&1 | tail -1)"
done
