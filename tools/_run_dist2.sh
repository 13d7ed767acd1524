#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -m gpu -v > gpurun_out/d2_tests.log 2>&1; tail -8 gpurun_out/d2_tests.log
TFELCU_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > gpurun_out/d2_bench.json 2> gpurun_out/d2_bench.err
grep "tfelcu trace\] rank" gpurun_out/d2_bench.err | sort | tail -4
tail -c 2500 gpurun_out/d2_bench.json
