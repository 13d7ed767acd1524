cd $GRAFT_REPO_ROOT
( time timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err ) 2>&1 | tail -3
tail -c 600 gpurun_out/r2_bench_n1_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_a.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','assembly','solve','e2e','like_for_like','other_configs'):
    print(k, json.dumps(d.get(k))[:1500])
PY
timeout 600 python -m pytest tests/test_cpp_surface.py -m gpu -x -q 2>&1 | tail -5
