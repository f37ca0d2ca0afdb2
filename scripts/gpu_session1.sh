#!/bin/bash
# First GPU bring-up: each group in its own process with a hard timeout so a hung kernel cannot eat the session.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -5 gpurun_out/$name.log; }
run t_sparse 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "normalize or perturb_merge or permute_null or device_null or syevd" 
run t_gemm_cg1 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "(gemm or gram) and (1-shape or syrk and 1- or accurate and 1)"
run t_gemm_cg2 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "(gemm or gram) and not (1-shape or syrk and 1- or accurate and 1)"
run t_misc 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "corr or topk or scores"
run t_pipeline 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu
run probe 600 python scripts/gpu_probe.py 4096 10000 20000
