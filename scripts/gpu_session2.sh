#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log; }
run t_ops 600 python -m pytest tests/test_ops_gpu.py -q -m gpu
run t_pipeline 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu
