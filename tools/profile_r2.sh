#!/bin/bash
# ncu evidence for the shipped round-2 build (run under gpurun, one GPU): launch lists + `--set full` captures of the dominant
# kernels.  Outputs land in gpurun_out/; tools/summarize_profiles_r2.py turns them into the tracked profiles/r2_*.
set -x
mkdir -p gpurun_out
B="python bench.py --no-extras --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2.csv $B --steps 2 --warmup 1 > gpurun_out/ncu_list_r2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render3_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_render3 $B --steps 1 --warmup 1 > gpurun_out/ncu_full_r2.log 2>&1
NFB_KERNEL=v6 timeout 400 ncu --set full --clock-control none --import-source on -k regex:render2_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_render2 $B --steps 1 --warmup 1 > gpurun_out/ncu_full_r2_v6.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_exact $B --steps 1 --warmup 1 --precision exact > gpurun_out/ncu_full_r2_exact.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render3_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2_cfg4 $B --steps 1 --warmup 1 --height 1024 --width 1024 --num-coarse 128 --num-fine 256 > gpurun_out/ncu_full_r2_cfg4.log 2>&1
# training step (fused trainer): launch list, then full captures of the three heavy kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/train_launches_r2.csv python tools/train_bench.py --steps 4 --warmup 3 --impl fused > gpurun_out/ncu_train_list_r2.log 2>&1
for k in dw_kernel chain_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_r2_$k python tools/train_bench.py --steps 2 --warmup 2 --impl fused > gpurun_out/ncu_full_r2_$k.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:render_kernel" -s 3 -c 1 -f -o gpurun_out/prof_r2_fwd_save python tools/train_bench.py --steps 2 --warmup 2 --impl fused > gpurun_out/ncu_full_r2_fwd_save.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python tools/train_bench.py --steps 20 --warmup 3 --impl fused > gpurun_out/r2_train_bench_n1.json 2> gpurun_out/r2_train_bench_n1.err
ls -la gpurun_out | tail -30
