#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=30 run t_gpu 900 python -m pytest tests -q -m gpu --durations=5
TAILN=12 run diag_B 500 python scripts/diag_eig_error.py 10000 20000 64
TAILN=8 run kb_gram_B 300 python scripts/kbench.py gram B 0,32
TAILN=6 run kb_gram_C 300 python scripts/kbench.py gram C 0
SCL_TRACE=1 TAILN=4 run trace_B 400 python scripts/trace_run.py B 2
