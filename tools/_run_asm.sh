#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q 2>&1 | tail -3
timeout 600 python tools/asm_sweep.py --n 4096 --configs auto auto+TFELCU_P1_PREFETCH=740 auto+TFELCU_P1_PREFETCH=1480 auto+TFELCU_P1_PREFETCH=2960 auto+TFELCU_P1_PREFETCH=6000 auto+TFELCU_NO_VTAB=1+TFELCU_P1_PREFETCH=1480 auto+TFELCU_P1STREAM=1 2>&1 | tee gpurun_out/r02_asm_sweep_prefetch.jsonl | cut -c1-220
