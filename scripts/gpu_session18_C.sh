#!/bin/bash
# BASELINE.json configs[2] shape on one GPU: 68k cells x 20k genes.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=8 run kb_all_C 200 python scripts/kbench.py all C 0
TAILN=5 run bench_C 600 python bench.py --workload C --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline
