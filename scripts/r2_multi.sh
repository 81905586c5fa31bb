#!/bin/bash
# Round 2, multi-GPU call: N-GPU tests through the public API and the torchrun path, cfg5 scaling lines.
# usage: r2_multi.sh "2 4 8"   (GPU counts to bench; the box must have at least the largest)
set -u
mkdir -p gpurun_out
NS=${1:-"2 4"}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m pytest tests/test_gpu_device_api.py tests/test_multigpu.py -m gpu -q -x -k "multi or sharded or partition" 2>&1 | tail -5
for N in $NS; do
  for W in cfg5; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
        bench.py --gpus $N --steps 5 --warmup 3 --workload $W > gpurun_out/m_${W}_n$N.json 2> gpurun_out/m_${W}_n$N.err
    echo "== $W N=$N rc=$? $(python -c "import json; d=json.load(open('gpurun_out/m_${W}_n$N.json')); print(round(d['ms_per_step'],3), 'ms/step', 'e2e', round(d['e2e']['ms_per_step'],1), 'ms', 'parity', d['parity']['ok'], d['residual']['rank_ms'], d['residual']['bounds'])" 2>&1 | tail -1)"
    tail -3 gpurun_out/m_${W}_n$N.err | cut -c1-300
  done
done
timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg5 --no-per-config --no-cpu > gpurun_out/m_cfg5_n1.json 2> gpurun_out/m_cfg5_n1.err
echo "== cfg5 N=1 rc=$? $(python -c "import json; d=json.load(open('gpurun_out/m_cfg5_n1.json')); print(round(d['ms_per_step'],3), 'ms/step', 'e2e', round(d['e2e']['ms_per_step'],1), 'first', round(d['e2e']['first_call_ms'],1))" 2>&1 | tail -1)"
