#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 140 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_B_short.log 2>&1; echo "rc=$?" >> gpurun_out/bench_B_short.log; tail -3 gpurun_out/bench_B_short.log
