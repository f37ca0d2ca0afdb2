#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=25 run t_ops 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "normalize or merge or null"
TAILN=12 run outlier 500 python scripts/diag_solver_outlier.py
TAILN=30 run own_B 500 python scripts/own_kernels.py B
run ncu_own 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 20000 --csv --log-file gpurun_out/launches_own_kernels_B.csv python scripts/own_kernels.py B
run ncu_dens 300 ncu --set full --clock-control none --import-source on -k regex:"k_densify" -s 1 -c 2 -o gpurun_out/prof_densify_B_r1c -f python scripts/kbench.py norm B
