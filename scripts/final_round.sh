#!/bin/bash
# Final measurement sweep of a round (run under gpurun, 1 GPU): tests, smoke, every workload's bench line, the
# reference arm of the default workload.  Results land in gpurun_out/final_*.json
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for w in cfg2 cfg1 cfg3 cfg4r cfg4 cfg5; do
  extra=""
  [ "$w" = "cfg4" ] && extra="--no-cpu"
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 $extra > gpurun_out/final_$w.json 2> gpurun_out/final_$w.err
  echo "== $w rc=$?"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference_cfg2.json 2> gpurun_out/final_reference_cfg2.err
echo "== reference rc=$?"
