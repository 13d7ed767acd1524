#!/bin/bash
# DRAM traffic of the SpMV at the row counts ONE rank holds when configs[1] (n = 4096) is split over 2 / 4 / 8 GPUs, measured on
# one GPU (ncu profiles one process): n = 2896 / 2048 / 1448 give 8.39 M / 4.20 M / 2.10 M rows. Run under gpurun.
mkdir -p gpurun_out
for n in 2896 2048 1448; do
  timeout 300 python tools/prof_step.py --n $n --its 12 --scatter auto --reps 1 > gpurun_out/tp_plain_$n.log 2>&1 || exit 1
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --kernel-name regex:k_spmv_stream --launch-skip 4 --launch-count 4 --csv --log-file gpurun_out/r02_spmv_traffic_n$n.csv python tools/prof_step.py --n $n --its 12 --scatter auto --reps 1 > gpurun_out/tp_ncu_$n.log 2>&1
  tail -n 4 gpurun_out/r02_spmv_traffic_n$n.csv | cut -c1-60,200-400
done
