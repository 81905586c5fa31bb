#!/bin/bash
# tests + bench lines for the main workloads (run under gpurun, 1 GPU); outputs in gpurun_out/
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in "$@"; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "== $w rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    r=d["roofline"]; e=d.get("e2e") or {}; c=d.get("cpu_baseline") or {}
    print(f"  value={d['value']:.2f} GF/s ms/step={d['ms_per_step']:.3f} kernel_ms={r['kernel_ms']:.3f} frac={r['frac']:.3f} achieved={r['achieved']:.0f}GB/s e2e_ms={e.get('ms_per_step')} launches={d['gpu_launches']} cpu_s={c.get('seconds')}")
except Exception as ex:
    print("  parse failed", ex)
PY
  tail -3 gpurun_out/bench_$w.err
done
