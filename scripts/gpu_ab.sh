timeout 300 python scripts/prof_encode.py --chunks 16384 --iters 3 2>&1 | tail -1
python -m pytest tests/test_gpu_encode.py -m gpu -x -q 2>&1 | tail -1
