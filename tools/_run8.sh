cd $GRAFT_REPO_ROOT
N=$1
for ov in 0 1; do
if [ $ov = 1 ]; then export TFELCU_OVERLAP=1; else unset TFELCU_OVERLAP; fi
TFELCU_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$ov bench.py --gpus $N --steps 3 --warmup 3 --e2e-steps $((1-ov)) > gpurun_out/r2_bench_n${N}_ov$ov.json 2> gpurun_out/r2_bench_n${N}_ov$ov.err
grep "tfelcu trace\] rank" gpurun_out/r2_bench_n${N}_ov$ov.err | sort | awk 'NR%3==0' | head -8
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n${N}_ov$ov.json').read().strip().splitlines()[-1])
print('N=$N overlap=$ov', d['value'], d['ms_per_step'], json.dumps(d['solve']), d['assembly']['ms'], d['roofline']['frac'], (d.get('parity') or {}).get('ok'), (d.get('e2e') or {}).get('value'))
PY
done
