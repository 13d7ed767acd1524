cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_dist.py -m gpu -v 2>&1 | tail -9 | tee gpurun_out/r2_dist_tests_n2.log
TFELCU_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err
grep "tfelcu trace\] rank" gpurun_out/r2_bench_n2_b.err | tail -2
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n2_b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['solve'], d['parity'], d['e2e'])"
