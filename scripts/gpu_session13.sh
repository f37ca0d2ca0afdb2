#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=30 run t_gpu 900 python -m pytest tests -q -m gpu --durations=5
TAILN=6 run kb_norm_B 300 python scripts/kbench.py norm B
TAILN=6 run trace_B 400 python scripts/trace_run.py B 3
TAILN=3 run bench_B 900 python bench.py
run ncu_own 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 20000 --csv --log-file gpurun_out/launches_own_kernels_B.csv python scripts/own_kernels.py B
run ncu_norm 300 ncu --set full --clock-control none --import-source on -k regex:"k_row_sum|k_gene_stats|k_cell_l2|k_gene_center|k_densify" -s 6 -c 6 -o gpurun_out/prof_norm_B_r1e -f python scripts/kbench.py norm B
