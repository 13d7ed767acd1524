set -x
cd $GRAFT_REPO_ROOT
export TFELCU_TRACE=1
timeout 900 python -m pytest tests/test_gpu_iluk.py -x -q 2>&1 | tail -40 > gpurun_out/r2_iluk_tests.log
cat gpurun_out/r2_iluk_tests.log | tail -30
timeout 600 python tools/stokes_probe.py 64 128 256 --ilufill=2 > gpurun_out/r2_stokes_probe_a.log 2>&1; tail -20 gpurun_out/r2_stokes_probe_a.log
timeout 600 python tools/stokes_probe.py 128 --ilufill=0 --restart=1000 >> gpurun_out/r2_stokes_probe_a.log 2>&1; tail -5 gpurun_out/r2_stokes_probe_a.log
unset TFELCU_TRACE
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gpu_tests_a.log; cat gpurun_out/r2_gpu_tests_a.log
