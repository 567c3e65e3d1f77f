# GPU tests, then decode timing of the 1 GiB text workload and of the 16 MiB streams with both expansion kernels.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/prof_decode.py --chunks 16384 --iters 4 2>&1 | tail -3
LZB_EXPAND=warp timeout 300 python scripts/prof_large.py 2>&1 | tail -4
LZB_EXPAND=cta timeout 300 python scripts/prof_large.py 2>&1 | tail -3
