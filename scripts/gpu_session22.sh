#!/bin/bash
# deep-pipelined line passes + fast reciprocal, TMA writer with cp.async-staged patches: parity, then tuning sweep
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
export SCL_DENSIFY=3 SCL_STAT_VARIANT=5
TAILN=8 run t_norm_v5w3 300 python -m pytest tests/test_ops_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu -k "normalize or sclens_matches"
export SCL_DENSIFY=2 SCL_STAT_VARIANT=4
TAILN=8 run t_norm_v4w2 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "normalize"
unset SCL_DENSIFY SCL_STAT_VARIANT
TAILN=24 run tune_B 200 python scripts/tune_norm.py B
TAILN=24 run tune_C 200 python scripts/tune_norm.py C
