cd $GRAFT_REPO_ROOT
N=$1
for mode in 1 0; do
TFELCU_TRACE=1 TFELCU_FUSED_DIST=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$mode bench.py --gpus $N --steps 3 --warmup 3 --e2e-steps 0 > gpurun_out/r2_bench_n${N}_fused$mode.json 2> gpurun_out/r2_bench_n${N}_fused$mode.err
grep "tfelcu trace\] rank" gpurun_out/r2_bench_n${N}_fused$mode.err | sort | awk 'NR%2==0' | head -8
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n${N}_fused$mode.json').read().strip().splitlines()[-1])
print('N=$N fused=$mode', d['value'], d['ms_per_step'], json.dumps(d['solve']), json.dumps(d['assembly'])[:200], (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('true_residual_rel'))
PY
done
