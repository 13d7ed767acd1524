#!/bin/bash
# One-GPU validation of a code state (run under gpurun): the -m gpu suite, smoke(), the default bench line, and the ncu
# launch list of tools/prof_step.py; everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/v_pytest.log
tail -4 gpurun_out/v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/v_smoke.log 2>&1; tail -2 gpurun_out/v_smoke.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['assembly']['ms'], d['assembly']['frac_of_peak'], d['solve'], d['roofline']['frac'], d['roofline']['moved_frac'], d['e2e']['value'], d['clocks'])
oc=d.get('other_configs') or {}
print(json.dumps(oc.get('c4', {}).get('solve', {}))[:400])
print(json.dumps(oc.get('c5', {}))[:700])
PY
tail -3 gpurun_out/v_bench.err
timeout 300 python tools/prof_step.py --n 4096 --its 12 --scatter auto --reps 1 > gpurun_out/p_plain.log 2>&1; tail -1 gpurun_out/p_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python tools/prof_step.py --n 4096 --its 12 --scatter auto --reps 1 > gpurun_out/p_ncu.log 2>&1; tail -1 gpurun_out/p_ncu.log
