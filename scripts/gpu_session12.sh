#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=30 run t_gpu 900 python -m pytest tests -q -m gpu --durations=5
TAILN=6 run kb_norm_B 300 python scripts/kbench.py norm B
SCL_TRACE=1 TAILN=60 run trace_B 400 python scripts/trace_run.py B 3
run ncu_stats 300 ncu --set full --clock-control none --import-source on -k regex:"k_row_sum|k_gene_stats|k_cell_l2|k_gene_center|k_densify" -s 6 -c 6 -o gpurun_out/prof_norm_B_r1d -f python scripts/kbench.py norm B
