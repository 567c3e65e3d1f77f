timeout 300 python scripts/prof_decode.py --chunks 16384 --iters 4 2>&1 | grep -E "iter 3|parity|rror"
