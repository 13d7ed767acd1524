#!/bin/bash
# one point of the strong-scaling curve of configs[1] (run under gpurun --gpus N): bash tools/run_scaling_point.sh N
mkdir -p gpurun_out
N=${1:-8}
TFELCU_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29730 bench.py --gpus $N --steps 5 --warmup 3 --no-extras > gpurun_out/d${N}_bench.json 2> gpurun_out/d${N}_bench.err
grep "tfelcu trace\] rank" gpurun_out/d${N}_bench.err | sort | awk 'NR%6==0' | head -8
python - <<PY
import json
d=json.loads(open('gpurun_out/d${N}_bench.json').read().strip().splitlines()[-1])
print('N=$N', d['value'], d['ms_per_step'], json.dumps(d['solve']), d['assembly']['ms'], d['roofline']['frac'], (d.get('parity') or {}).get('ok'), (d.get('e2e') or {}).get('value'))
PY
