python -m pytest tests/test_gpu_encode.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python scripts/prof_encode.py --chunks 16384 --iters 3 2>&1 | tail -1
