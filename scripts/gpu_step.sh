mkdir -p gpurun_out
timeout 300 python scripts/try_long.py --no-time > gpurun_out/try_long.log 2>&1; grep -v "^ok " gpurun_out/try_long.log | tail -3 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
