#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
run t_gemm 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "gemm or gram or corr or topk or scores"
TAILN=12 run kb_gram_B 300 python scripts/kbench.py gram B
TAILN=12 run kb_gram_C 300 python scripts/kbench.py gram C
run t_pipeline 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu -x
