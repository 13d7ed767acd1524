#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/t_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/t_pytest.log
tail -15 gpurun_out/t_pytest.log
( cd /tmp && timeout 600 /root/repo/examples/bin/poisson_device 4096 ) > gpurun_out/t_poisson_device_4096.log 2>&1; tail -12 gpurun_out/t_poisson_device_4096.log
