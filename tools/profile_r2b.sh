#!/bin/bash
# second part of the round-2 captures: training launch list and the dX chain kernel
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 100 --csv --log-file gpurun_out/train_launches_r2.csv python tools/train_bench.py --steps 4 --warmup 3 --impl fused > gpurun_out/ncu_train_list_r2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_chain_kernel python tools/train_bench.py --steps 2 --warmup 2 --impl fused > gpurun_out/ncu_full_r2_chain_kernel.log 2>&1
