#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/v_pytest.log
tail -6 gpurun_out/v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/v_smoke.log 2>&1; tail -2 gpurun_out/v_smoke.log
( time timeout 900 python bench.py --steps 5 --warmup 3 --no-extras ) > gpurun_out/v_bench.log 2> gpurun_out/v_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_bench.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['assembly'], d['solve'], d['roofline']['frac'], d['parity']['ok'] if d.get('parity') else None, d['e2e']['value'])
PY
tail -3 gpurun_out/v_bench.err
