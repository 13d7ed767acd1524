cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_dist.py -m gpu -v 2>&1 | tail -15 | tee gpurun_out/r2_dist_tests_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err
tail -c 400 gpurun_out/r2_bench_n2_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2_a.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','assembly','solve','parity','e2e'):
    print(k, json.dumps(d.get(k))[:900])
PY
