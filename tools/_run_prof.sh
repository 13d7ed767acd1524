#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_step.py --n 4096 --its 12 --scatter auto --reps 1 > gpurun_out/p_plain.log 2>&1; tail -1 gpurun_out/p_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_final.csv python tools/prof_step.py --n 4096 --its 12 --scatter auto --reps 1 > gpurun_out/p_ncu.log 2>&1; tail -1 gpurun_out/p_ncu.log
( cd /tmp && time timeout 1500 /root/repo/examples/bin/navier_stokes_2d_p2_p1 256 3 ) > gpurun_out/r02_ns_n256_3steps.log 2>&1; grep -v "^  " gpurun_out/r02_ns_n256_3steps.log | tail -12
