#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/asm_sweep.py --n 4096 --configs auto auto+TFELCU_P1_VARIANT=1 auto+TFELCU_P1_VARIANT=2 auto+TFELCU_P1_VARIANT=4 auto+TFELCU_P1_VARIANT=5 2>&1 | tee gpurun_out/r02_asm_sweep_prologue2.jsonl | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_assemble_p1_ell --launch-skip 1 --launch-count 1 -o gpurun_out/r2_p1ell_v3 -f python tools/asm_sweep.py --n 2048 --reps 1 --configs auto > gpurun_out/ncu_run.log 2>&1; tail -2 gpurun_out/ncu_run.log | cut -c1-150
