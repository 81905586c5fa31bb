#!/bin/bash
# Round 2, call 19 (1 GPU): paneled transpose entries packed into 32 bits (12 bytes per stream entry instead of 16):
# parity (triple tests + fuzz), then cfg5 with the automatic panel count and forced 3 / 4 / 5, cfg3.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_device_api.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py tests/test_reference_suite.py -m gpu -q -x -k "triple or cfg3 or cfg5 or multi" 2>&1 | tail -4
for P in 0 3 4 5; do
  SPGEMM_B200_TRIPLE_PANELS=$P timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg5 --no-per-config --no-cpu --no-e2e \
      > gpurun_out/c19_cfg5_np$P.json 2> gpurun_out/c19_cfg5_np$P.err
  echo "== cfg5 panels=$P rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c19_cfg5_np$P.json')); print(round(d['ms_per_step'],3), 'ms/step', d['phases_ms'], 'frac', round(d['roofline']['frac'],4), 'cached', round(d['cached_transpose']['ms_per_step'],3))" 2>&1 | tail -1)"
done
timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg3 --no-per-config --no-cpu --no-e2e > gpurun_out/c19_cfg3.json 2> gpurun_out/c19_cfg3.err
echo "== cfg3 rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c19_cfg3.json')); print(round(d['ms_per_step'],3), 'ms/step', d['phases_ms'])" 2>&1 | tail -1)"
SPGEMM_B200_TRIPLE_GENERIC=1 timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg5 --no-per-config --no-cpu --no-e2e > gpurun_out/c19_cfg5_generic.json 2> gpurun_out/c19_cfg5_generic.err
echo "== cfg5 generic rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c19_cfg5_generic.json')); print(round(d['ms_per_step'],3), 'ms/step', d['phases_ms'])" 2>&1 | tail -1)"
