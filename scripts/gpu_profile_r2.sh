#!/bin/bash
# round-2 ncu captures at the headline shape (68k x 20k); run under gpurun from the repo root
set -x
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -k regex:k_ -c 400 --csv --log-file gpurun_out/r2_launches_own_C.csv python scripts/prof_own.py C > gpurun_out/r2_prof_own_launches.log 2>&1
$NCU --set full --import-source on -k regex:"k_merge_lines|k_count_adds|k_bucket_adds|k_densify_tma2|k_gemm_umma|k_zc_insert|k_zero_cand_flags" -c 16 -o gpurun_out/r2_own_C_full -f python scripts/prof_own.py C > gpurun_out/r2_prof_own_full.log 2>&1
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active -k regex:k_sytrd -c 1 --csv --log-file gpurun_out/r2_sytrd_20000_ncu.csv python scripts/sytrd_trace.py 20000 > gpurun_out/r2_prof_sytrd.log 2>&1
tail -3 gpurun_out/r2_prof_own_launches.log gpurun_out/r2_prof_own_full.log gpurun_out/r2_prof_sytrd.log
ls -la gpurun_out/r2_own_C_full.ncu-rep gpurun_out/r2_launches_own_C.csv gpurun_out/r2_sytrd_20000_ncu.csv
