set -x
timeout 300 python -m pytest tests/test_backward_gpu.py tests/test_fused_train_gpu.py tests/test_render_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 120 python tools/train_bench.py --steps 30 --warmup 5 --impl fused 2>&1 | tail -2 | cut -c1-400
