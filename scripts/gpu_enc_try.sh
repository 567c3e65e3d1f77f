# Encoder iteration pass: encode tests (bounded), then stage timings on the 64 KiB text workload.
set -x
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q 2>&1 | tail -15
timeout 300 python scripts/prof_encode.py --chunks 16384 --iters 3 2>&1 | tail -5
