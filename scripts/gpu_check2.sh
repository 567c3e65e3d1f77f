timeout 900 python -m pytest tests/test_gpu_encode.py -x -q 2>&1 | tail -2
timeout 300 python scripts/prof_encode.py --chunks 16384 --iters 2 2>&1 | tail -1 | cut -c1-140
timeout 600 python scripts/bench_mixed.py --mib-per-class 512 2>&1 | grep "rep encode stages" | cut -c1-200
