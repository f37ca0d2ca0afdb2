#!/bin/bash
G=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29611 scripts/multi_gpu_check.py > gpurun_out/r2_multi_check_${G}gpu.log 2>&1
tail -3 gpurun_out/r2_multi_check_${G}gpu.log
timeout 500 $TR --master-port 29612 bench.py --gpus $G --steps 2 --warmup 2 > gpurun_out/r2_bench_C_${G}gpu.json 2> gpurun_out/r2_bench_C_${G}gpu.err
grep "warmup 1\|Error\|error" gpurun_out/r2_bench_C_${G}gpu.err | tail -3 | cut -c1-200; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_C_${G}gpu.json
