#!/bin/bash
# source-level ncu capture of the line passes and both dense writers at the 68k x 20k shape
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout -s KILL $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -${TAILN:-6} gpurun_out/$name.log; }
run ncu_src_C 400 ncu --set full --clock-control none --import-source on -k regex:"k_lines|k_densify" -s 5 -c 15 -o gpurun_out/prof_norm_C_src -f python scripts/prof_norm.py C
ls -la gpurun_out/*.ncu-rep
