cd $GRAFT_REPO_ROOT
export TFELCU_TRACE=1
L=gpurun_out/r2_tri_sweep2.log; : > $L
run() { echo "== $*" >> $L; env "$@" timeout 300 python tools/stokes_probe.py 256 --ilufill=2 --maxits=40 2>&1 | grep "stokes n\|trace" | sed 's/.*gmres/gmres/' >> $L; }
run TFELCU_TRI_GATE_NS=-1
run TFELCU_TRI_GATE_NS=0
run TFELCU_TRI_GATE_NS=100
run TFELCU_TRI_GATE_NS=400
run TFELCU_TRI_GATE_NS=100 TFELCU_TRI_CTAS=4
run TFELCU_TRI_GATE_NS=100 TFELCU_TRI_CTAS=2
run TFELCU_TRI_GATE_NS=100 TFELCU_TRI_LANES=16
cat $L
timeout 600 python -m pytest tests/test_gpu_iluk.py -x -q 2>&1 | tail -5
timeout 900 python tools/stokes_probe.py 256 512 --ilufill=2 2>&1 | grep "stokes n\|trace" | tee gpurun_out/r2_stokes_probe_b.log
