cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
for mode in 0 1; do
TFELCU_TRACE=1 TFELCU_FUSED_DIST=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2970$mode bench.py --gpus 2 --steps 1 --warmup 3 --e2e-steps 0 > gpurun_out/r2_bench_n2_fused$mode.json 2> gpurun_out/r2_bench_n2_fused$mode.err &
sleep 45; nvidia-smi --query-gpu=index,clocks.sm,power.draw --format=csv,noheader; wait
grep "tfelcu trace\] rank" gpurun_out/r2_bench_n2_fused$mode.err | tail -2
done
