#!/bin/bash
# strip passes (shared-memory parameter gathers): parity, then timing against the line passes at B and C
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
export SCL_STAT_VARIANT=8
TAILN=8 run t_norm_strips 300 python -m pytest tests/test_ops_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu -k "normalize or sclens_matches"
unset SCL_STAT_VARIANT
TAILN=10 run tune_B 200 python scripts/tune_norm.py B
TAILN=10 run tune_C 200 python scripts/tune_norm.py C
