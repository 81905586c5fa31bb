#!/bin/bash
# Round 2, GPU call 8: the default bench line (with per_config and cpu_baseline), the reference arm, host zero-fill threads.
set -u
mkdir -p gpurun_out
nproc; free -g | head -2
( time python bench.py > gpurun_out/c8_default.json 2> gpurun_out/c8_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/c8_default.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],1), 'first', round(d['e2e']['first_call_ms']), 'cpu', d.get('cpu_baseline'))
for k,v in d.get('per_config',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, 'ms', round(v['ms_per_step'],3), 'GF', round(v['value'],1), 'frac', round(v['roofline']['frac'],3), 'e2e', v['e2e'].get('ms_per_step'), 'setup', v['setup_s'], v['phases_ms'])
PY
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c8_reference.json 2> gpurun_out/c8_reference.err ) 2>&1 | grep real
python -c "import json; d=json.load(open('gpurun_out/c8_reference.json')); print('reference', d['value'], d['ms_per_step'], d['cpu_baseline'])"
for zt in 0 4 8 16 32; do
  SPGEMM_B200_ZERO_THREADS=$zt python bench.py --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/c8_zt$zt.json 2> gpurun_out/c8_zt$zt.err
  python -c "import json; d=json.load(open('gpurun_out/c8_zt$zt.json')); print('zero threads $zt: e2e', round(d['e2e']['ms_per_step'],1), d['e2e']['device_ms'])"
done
