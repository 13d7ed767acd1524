#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['assembly'], d['solve'], d['roofline']['frac'], d['roofline']['moved_frac'], d['e2e']['value'], d['clocks'])
print(json.dumps(d.get('like_for_like'))[:600])
print({k: (v.get('assembly_ms'), v.get('its'), v.get('solve_s')) if isinstance(v, dict) else v for k, v in (d.get('other_configs') or {}).items()})
PY
tail -4 gpurun_out/f_bench_n1.err
