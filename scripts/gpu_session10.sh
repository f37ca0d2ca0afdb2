#!/bin/bash
# New strip writer, bitmap merge, bias calibration, Float64 refinement: full suite, numerics study, benches, ncu captures.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=40 run t_gpu 1500 python -m pytest tests -q -m gpu --durations=10
TAILN=14 run diag_B 600 python scripts/diag_eig_error.py 10000 20000 64
TAILN=8 run kb_norm_B 300 python scripts/kbench.py norm B
TAILN=3 run bench_B 900 python bench.py
run ncu_dens 400 ncu --set full --clock-control none --import-source on -k regex:"k_densify|k_strip" -s 2 -c 4 -o gpurun_out/prof_densify_B_r1b -f python scripts/kbench.py norm B
run ncu_stats 400 ncu --set full --clock-control none -k regex:"k_row_sum|k_gene_stats|k_cell_l2|k_gene_center|k_reduce|k_inv_s" -s 8 -c 8 -o gpurun_out/prof_stats_B_r1 -f python scripts/kbench.py norm B
run ncu_list_small 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 30000 --csv --log-file gpurun_out/launches_bench_small.csv python bench.py --workload small --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline
