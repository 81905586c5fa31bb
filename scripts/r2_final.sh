#!/bin/bash
# Round 2 final pass (1 GPU): what the driver runs at round end -- GPU tests, smoke, the default bench line, the reference arm.
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py > gpurun_out/final_default.json 2> gpurun_out/final_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_default.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],1), 'first', round(d['e2e']['first_call_ms']), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('cpu', d.get('cpu_baseline'))
for k,v in d.get('per_config',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, 'ms', round(v['ms_per_step'],3), 'GF', round(v['value'],1), 'frac', round(v['roofline']['frac'],3), 'e2e', v['e2e'].get('ms_per_step'), 'setup', v['setup_s'], v['phases_ms'])
PY
( time python bench.py --impl reference > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err ) 2>&1 | grep real
python -c "import json; d=json.load(open('gpurun_out/final_reference.json')); print('reference', d['value'], d['ms_per_step'], d['cpu_baseline'], d['config'].keys())"
