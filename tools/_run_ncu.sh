#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/asm_sweep.py --n 2048 --reps 1 --configs auto > gpurun_out/ncu_plain.log 2>&1; tail -1 gpurun_out/ncu_plain.log | cut -c1-150
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_assemble_p1_ell --launch-skip 1 --launch-count 1 -o gpurun_out/r2_p1ell_v2 -f python tools/asm_sweep.py --n 2048 --reps 1 --configs auto > gpurun_out/ncu_run.log 2>&1; tail -2 gpurun_out/ncu_run.log | cut -c1-150
ls -la gpurun_out/*.ncu-rep
