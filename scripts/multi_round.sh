#!/bin/bash
# usage: multi_round.sh N  -- bench lines of cfg2 and cfg5 at N GPUs (run under gpurun --gpus N)
N=$1
for w in cfg2 cfg5; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 5 --warmup 3 --workload $w > gpurun_out/bench${N}_$w.json 2> gpurun_out/bench${N}_$w.err
  echo "== $w N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench${N}_$w.json") if l.startswith("{")][-1])
    print("  value=%.1f ms/step=%.3f e2e_ms=%s gather=%s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"].get("gather")))
except Exception as ex:
    print("  parse failed", ex)
PY
done
