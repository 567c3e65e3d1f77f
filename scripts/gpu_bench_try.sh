set -x
timeout 600 python -m pytest tests/test_gpu_encode.py tests/test_gpu_configs.py -x -q 2>&1 | tail -3
timeout 300 python scripts/prof_encode.py --chunks 16384 --iters 2 2>&1 | tail -1
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --total-gib 2 --wave-gib 1 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo rc=$?; tail -3 gpurun_out/bench_c5.err
