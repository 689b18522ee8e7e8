set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 220 2>&1 | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --extras rows,train > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
tail -c 1500 gpurun_out/r2c_bench_n2.json
tail -3 gpurun_out/r2c_bench_n2.err
