#!/bin/bash
# new line-pass statistics kernels: parity first, then kernel timings at both benchmark shapes and an ncu --set full capture
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=12 run t_ops 400 python -m pytest tests/test_ops_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu --durations=5
TAILN=5 run kb_norm_B 200 python scripts/kbench.py norm B
TAILN=5 run kb_norm_C 200 python scripts/kbench.py norm C
run ncu_norm_C 300 ncu --set full --clock-control none --import-source on -k regex:"k_lines|k_row_sum|k_reduce_vec|k_densify" -s 10 -c 10 -o gpurun_out/prof_norm_C_r1 -f python scripts/kbench.py norm C
ls -la gpurun_out/*.ncu-rep
