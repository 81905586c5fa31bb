#!/bin/bash
# Round 2, second profiling pass (1 GPU) after the 12-byte panel entries: launch list of the default bench command and a
# --set full capture of k_triple_runs on cfg5 (each ncu command preceded by the same command without ncu), plus the
# cfg4r end-to-end line with the LRU pinned cache.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg4r --no-cpu --no-per-config > gpurun_out/p2_cfg4r.json 2> gpurun_out/p2_cfg4r.err
echo "== cfg4r rc=$? $(python -c "import json; d=json.load(open('gpurun_out/p2_cfg4r.json')); print(round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['ms_per_step'],1), d['e2e']['device_ms'])" 2>&1 | tail -1)"
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-per-config"
$B > gpurun_out/p2_plain_default.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/p2_launches_default.csv $B > gpurun_out/p2_ncu_launches.log 2>&1
echo "launch list default rc=$?"
$B --workload cfg5 > gpurun_out/p2_plain_cfg5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_triple_runs -s 2 -c 1 -f -o gpurun_out/p2_prof_triple_runs_cfg5 \
    $B --workload cfg5 > gpurun_out/p2_ncu_full_cfg5.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | grep p2_prof
