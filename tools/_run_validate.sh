#!/bin/bash
# one-GPU validation: gpu test suite, smoke, default bench line (with extras), reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/v_pytest.log
tail -5 gpurun_out/v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/v_smoke.log 2>&1; tail -2 gpurun_out/v_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/v_bench.log 2> gpurun_out/v_bench.err; tail -c 3000 gpurun_out/v_bench.log; tail -4 gpurun_out/v_bench.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/v_ref.log 2> gpurun_out/v_ref.err; tail -c 1500 gpurun_out/v_ref.log; tail -4 gpurun_out/v_ref.err
