#!/bin/bash
# final single-GPU validation of the round: the driver's own sequence (GPU tests, smoke, bench) + an ncu look at k_sytrd
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; tail -2 gpurun_out/r2_smoke_final.log
timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_C_final.json 2> gpurun_out/r2_bench_C_final.err; grep warmup gpurun_out/r2_bench_C_final.err | tail -2; head -c 500 gpurun_out/r2_bench_C_final.json; echo
timeout 400 ncu --clock-control none --set full --import-source on -k regex:k_sytrd -c 1 -o gpurun_out/r2_sytrd_full -f python scripts/sytrd_trace.py 12000 > gpurun_out/r2_prof_sytrd_full.log 2>&1; tail -2 gpurun_out/r2_prof_sytrd_full.log
