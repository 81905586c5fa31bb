#!/bin/bash
# usage: sweep.sh <workload> VAR=val ... ; prints kernel_ms / ms_per_step for one env setting
w=$1; shift
env "$@" python bench.py --workload $w --steps 5 --no-cpu --no-e2e | python -c "
import json,sys; d=json.load(sys.stdin); print('$*', 'kernel_ms=%.3f ms_per_step=%.3f value=%.1f' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['value']))"
