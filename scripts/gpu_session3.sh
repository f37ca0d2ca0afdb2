#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log; }
run t_pipeline 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu
run t_large 900 python -m pytest tests/test_large_gpu.py -q -m gpu
run smoke 300 python __graft_entry__.py --smoke
run bench_small 600 python bench.py --workload small --steps 1 --warmup 1
run bench_B 1500 python bench.py --workload B --steps 1 --warmup 1
