# quick check after an encoder change: encode tests, headline stage times, per-class stage times of the mixed corpus, 8 x 16 MiB
timeout 900 python -m pytest tests/test_gpu_encode.py tests/test_gpu_configs.py -x -q 2>&1 | tail -3
timeout 300 python scripts/prof_encode.py --chunks 16384 --iters 2 2>&1 | tail -1 | cut -c1-520
timeout 600 python scripts/bench_mixed.py --mib-per-class 512 2>&1 | grep "encode stages" | cut -c1-300
timeout 300 python scripts/try_long.py 2>&1 | grep -v "^ok " | tail -4 | cut -c1-400
