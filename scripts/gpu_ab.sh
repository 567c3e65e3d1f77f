timeout 300 python scripts/prof_decode.py --chunks 16384 --iters 4 2>&1 | grep -E "iter 3|parity|rror"
python -m pytest tests/test_gpu_decode.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -1
