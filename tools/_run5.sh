cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2_tile_tests.log
timeout 300 python tools/asm_sweep.py --n 4096 --configs rows auto 2>&1 | tail -5 | tee gpurun_out/r2_asm_sweep_tile.jsonl
TFELCU_NO_ISO=1 timeout 300 python tools/asm_sweep.py --n 4096 --configs auto 2>&1 | tail -1 | sed 's/auto/auto-no-iso/' | tee -a gpurun_out/r2_asm_sweep_tile.jsonl
