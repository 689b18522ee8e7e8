#!/bin/bash
# Final round-2 pass on the shipped build (one GPU, under gpurun): the whole -m gpu suite, the default bench line, then the
# evidence for the training kernels as they ship (record-saver warps in the forward / chain, 8-group one-launch weight-gradient
# kernel): launch list + one `ncu --set full` capture each.  tools/summarize_profiles_r2.py --r2c turns them into profiles/r2c_*.
set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x --timeout 100 > gpurun_out/r2c_gputests.log 2>&1
tail -3 gpurun_out/r2c_gputests.log
timeout 330 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
cut -c1-300 gpurun_out/r2c_bench_n1.json
timeout 100 python tools/train_bench.py --steps 30 --warmup 5 --impl fused > gpurun_out/r2c_train_bench_n1.json 2> gpurun_out/r2c_train_bench_n1.err
T="python tools/train_bench.py --steps 2 --warmup 2 --impl fused"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 100 --csv --log-file gpurun_out/train_launches_r2c.csv python tools/train_bench.py --steps 4 --warmup 3 --impl fused > gpurun_out/ncu_train_list_r2c.log 2>&1
for k in dw_kernel chain_kernel render_kernel; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_r2c_$k $T > gpurun_out/ncu_full_r2c_$k.log 2>&1
done
ls -la gpurun_out | grep r2c
