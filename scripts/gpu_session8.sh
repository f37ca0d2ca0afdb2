#!/bin/bash
# Re-entry session: full GPU suite, smoke, default bench, kernel benches, ncu launch list + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
TAILN=25 run t_gpu 1200 python -m pytest tests -q -m gpu -x --durations=20
run smoke 300 python __graft_entry__.py --smoke
TAILN=3 run bench_B 1500 python bench.py
TAILN=12 run kb_all_B 400 python scripts/kbench.py all B 0,32
TAILN=6 run kb_gram_C 300 python scripts/kbench.py gram C 0,32
run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_bench_B.csv python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline
run ncu_gram 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_umma" -s 1 -c 2 -o gpurun_out/prof_gram_B_r1 -f python scripts/kbench.py gram B 0
run ncu_dens 600 ncu --set full --clock-control none --import-source on -k regex:"k_densify" -s 1 -c 2 -o gpurun_out/prof_densify_B_r1 -f python scripts/kbench.py norm B
