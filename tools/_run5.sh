cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q 2>&1 | tail -25 | tee gpurun_out/r2_tile_tests.log
timeout 300 python tools/asm_sweep.py --n 4096 --configs rows tile auto 2>&1 | tail -5 | tee gpurun_out/r2_asm_sweep_tile.jsonl
TFELCU_NO_COMPACT=1 timeout 300 python tools/asm_sweep.py --n 4096 --configs auto 2>&1 | tail -1 | sed 's/auto/auto-16B-records/' | tee -a gpurun_out/r2_asm_sweep_tile.jsonl
timeout 300 python tools/asm_sweep.py --n 2048 --reps 2 --configs auto > gpurun_out/plain_tile.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_assemble_p1_ell -s 1 -c 2 -o gpurun_out/r2_p1ell_prof python tools/asm_sweep.py --n 2048 --reps 2 --configs auto > gpurun_out/ncu_tile.log 2>&1; tail -3 gpurun_out/ncu_tile.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_gpu_tests_b.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1_b.json').read().strip().splitlines()[-1]); print(d['value'], d['assembly'], d['solve'])"
