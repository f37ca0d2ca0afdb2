#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=30 run t_ops 600 python -m pytest tests/test_ops_gpu.py -q -m gpu --durations=25
run ncu_gram 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_umma" -s 1 -c 1 -o gpurun_out/prof_gram_B_r1b -f python scripts/kbench.py gram B 0
