#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=30 run t_gpu 900 python -m pytest tests -q -m gpu --durations=5
TAILN=6 run kb_norm_B 300 python scripts/kbench.py norm B
SCL_TRACE=1 TAILN=6 run trace_B 400 python scripts/trace_run.py B 2
