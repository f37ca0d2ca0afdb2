#!/bin/bash
# 2 GPUs: multi-rank parity test and the bench line including the multi-rank end-to-end leg
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=6 run t_multigpu 300 python -m pytest tests/test_multigpu.py -q -m gpu
TAILN=6 run bench_B_2gpu 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 1 --warmup 1
