#!/bin/bash
# Round 2, call 25 (1 GPU): per-entry metadata prepared once per call (SPGEMM_B200_TRIPLE_ENTRY_META) and the next item's
# ticket drawn ahead (SPGEMM_B200_TRIPLE_TICKET_AHEAD): parity with both on, then the four combinations on cfg5.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_device_api.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -m gpu -q -x -k "triple or cfg3 or cfg5" 2>&1 | tail -2
for V in "0 0" "1 0" "0 1" "1 1"; do
  set -- $V
  SPGEMM_B200_TRIPLE_ENTRY_META=$1 SPGEMM_B200_TRIPLE_TICKET_AHEAD=$2 timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg5 --no-per-config --no-cpu --no-e2e > gpurun_out/c25_cfg5_m$1_t$2.json 2> gpurun_out/c25_cfg5_m$1_t$2.err
  echo "== cfg5 meta=$1 ticket=$2 rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c25_cfg5_m$1_t$2.json')); print(round(d['ms_per_step'],3), 'ms/step', d['phases_ms'], 'frac', round(d['roofline']['frac'],4))" 2>&1 | tail -1)"
done
timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg3 --no-per-config --no-cpu --no-e2e > gpurun_out/c25_cfg3.json 2> gpurun_out/c25_cfg3.err
echo "== cfg3 rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c25_cfg3.json')); print(round(d['ms_per_step'],3), 'ms/step', d['phases_ms'])" 2>&1 | tail -1)"
