#!/bin/bash
# TMA bulk-store dense writer: parity, then a tuning sweep of the statistics line passes and both writers at B and C
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
export SCL_DENSIFY=1
TAILN=8 run t_norm_tma 300 python -m pytest tests/test_ops_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu -k "normalize or sclens_matches"
unset SCL_DENSIFY
TAILN=20 run tune_B 200 python scripts/tune_norm.py B
TAILN=20 run tune_C 200 python scripts/tune_norm.py C
