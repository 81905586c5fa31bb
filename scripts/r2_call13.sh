#!/bin/bash
# Round 2, GPU call 13: pass 2 of the rank kernel in column sub-windows for heavy rows (cfg4 / cfg4r), sizes and off.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py tests/test_gpu_fullsize.py -m gpu -q -x -k "not triple and not cfg5 and not cfg3 and not cfg2" 2>&1 | tail -3
SPGEMM_B200_RANK_SUBWIN=512 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py -m gpu -q -x -k "sparse or bins or wide or heavy or cfg4r or cfg1 or golden or seeded" 2>&1 | tail -3
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-per-config"
run() {  # tag env...
  tag=$1; shift
  env "$@" $B --workload $W > gpurun_out/c13_${W}_$tag.json 2> gpurun_out/c13_${W}_$tag.err
  echo "== $W $tag rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/c13_${W}_$tag.json')); print(round(d['ms_per_step'],3), 'ms', round(d['roofline']['kernel_ms'],3), 'kernel ms', d['phases_ms'], 'frac', round(d['roofline']['frac'],3))" 2>&1 | tail -1)"
}
W=cfg4
run off SPGEMM_B200_RANK_SUBWIN=0
for k in 16384 32768 65536 131072; do run sub$k SPGEMM_B200_RANK_SUBWIN=$k; done
W=cfg4r
run off SPGEMM_B200_RANK_SUBWIN=0
run sub8k SPGEMM_B200_RANK_SUBWIN=8192
run sub32k X=1
