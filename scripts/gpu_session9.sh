#!/bin/bash
# Numerics study of the tensor-core accumulation (chunk length, diagonal), full GPU suite, ncu launch list (own kernels) + densify capture.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=60 run diag_B 900 python scripts/diag_eig_error.py 10000 20000 64,32,16,8
TAILN=40 run t_gpu 1500 python -m pytest tests -q -m gpu --durations=15 --deselect tests/test_large_gpu.py::test_signal_stage_full_size
TAILN=8 run kb_norm_B 400 python scripts/kbench.py norm B
run ncu_dens 600 ncu --set full --clock-control none --import-source on -k regex:"k_densify" -s 1 -c 2 -o gpurun_out/prof_densify_B_r1 -f python scripts/kbench.py norm B
run ncu_list 700 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 6000 --csv --log-file gpurun_out/launches_bench_B.csv python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline
