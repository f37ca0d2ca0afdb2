#!/bin/bash
# Re-entry session: full GPU test-suite, smoke, bench (small + B), ncu launch list + full capture of the top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -6 gpurun_out/$name.log; }
run t_ops 600 python -m pytest tests/test_ops_gpu.py -q -m gpu
run t_pipeline 900 python -m pytest tests/test_pipeline_gpu.py -q -m gpu
run t_large 900 python -m pytest tests/test_large_gpu.py -q -m gpu
run smoke 300 python __graft_entry__.py --smoke
run sig_B 300 python scripts/profile_signal.py B 3
run sig_C 600 python scripts/profile_signal.py C 2
run bench_B 1500 python bench.py --workload B --steps 1 --warmup 3
run ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_signal_B.csv python scripts/profile_signal.py B 1
run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_umma|k_densify" -c 4 -o gpurun_out/prof_gram_B -f python scripts/profile_signal.py B 1
