#!/bin/bash
# round-1 final state: full -m gpu suite, smoke, default bench line, own-kernel ncu launch list, ncu --set full of the normalisation kernels
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=12 run t_gpu 600 python -m pytest tests -q -m gpu --durations=8
run smoke 200 python __graft_entry__.py --smoke
TAILN=3 run bench_B 600 python bench.py
run ncu_own 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 20000 --csv --log-file gpurun_out/launches_own_kernels_B.csv python scripts/own_kernels.py B
run ncu_norm 200 ncu --set full --clock-control none --import-source on -k regex:"k_lines|k_row_sum|k_reduce_vec|k_inv_s|k_strip_offsets|k_densify" -s 11 -c 22 -o gpurun_out/prof_norm_B_final -f python scripts/prof_norm.py B
ls -la gpurun_out/*.ncu-rep
