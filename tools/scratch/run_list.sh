mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 100 --csv --log-file gpurun_out/train_launches_r2c.csv python tools/train_bench.py --steps 4 --warmup 3 --impl fused > gpurun_out/ncu_train_list_r2c.log 2>&1
tail -3 gpurun_out/ncu_train_list_r2c.log | cut -c1-300
