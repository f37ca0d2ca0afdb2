#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=12 run t_gpu 900 python -m pytest tests -q -m gpu --durations=5
run smoke 300 python __graft_entry__.py --smoke
TAILN=3 run bench_B 900 python bench.py
TAILN=8 run kb_all_B 300 python scripts/kbench.py all B 0
run ncu_own 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 20000 --csv --log-file gpurun_out/launches_own_kernels_B.csv python scripts/own_kernels.py B
run ncu_norm 300 ncu --set full --clock-control none --import-source on -k regex:"k_row_sum|k_gene_stats|k_cell_l2|k_gene_center|k_cell_finish|k_densify|k_strip" -s 9 -c 9 -o gpurun_out/prof_norm_B_r1f -f python scripts/kbench.py norm B
run ncu_merge 300 ncu --set full --clock-control none --import-source on -k regex:"k_merge_lines|k_bucket_adds|k_count_adds" -c 8 -o gpurun_out/prof_merge_B_r1 -f python scripts/own_kernels.py B
