cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q 2>&1 | tail -25 | tee gpurun_out/r2_tile_tests.log
timeout 300 python tools/asm_sweep.py --n 4096 --configs rows tile 2>&1 | tail -5 | tee gpurun_out/r2_asm_sweep_tile.jsonl
timeout 300 python tools/asm_sweep.py --n 2048 --reps 2 --configs tile > gpurun_out/plain_tile.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiles -s 1 -c 2 -o gpurun_out/r2_tile_prof python tools/asm_sweep.py --n 2048 --reps 2 --configs tile > gpurun_out/ncu_tile.log 2>&1; tail -3 gpurun_out/ncu_tile.log
