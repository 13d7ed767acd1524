cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_gpu_tests_c.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1_c.json').read().strip().splitlines()[-1]); print(d['value'], d['assembly']['ms'], d['solve'], d['roofline']['frac'], d['roofline']['moved_frac'], d.get('fp64_fma_peak_tflops'))"
