#!/bin/bash
# Round 2, call 24 (2 GPUs): the multi-GPU paths with the final kernels -- driver tests, fuzz through the driver, cfg5 at N=2.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_device_api.py tests/test_multigpu.py tests/test_gpu_fuzz.py -m gpu -q -x -k "multi or sharded or partition" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --steps 5 --warmup 3 --workload cfg5 > gpurun_out/c24_cfg5_n2.json 2> gpurun_out/c24_cfg5_n2.err
echo "== cfg5 N=2 rc=$? $(python -c "import json; d=json.load(open('gpurun_out/c24_cfg5_n2.json')); print(round(d['ms_per_step'],3), 'ms/step', 'e2e', round(d['e2e']['ms_per_step'],1), 'ms', 'parity', d['parity']['ok'], d['residual']['rank_ms'], d['residual']['bounds'], 'cached', round(d['cached_transpose']['ms_per_step'],3))" 2>&1 | tail -1)"
