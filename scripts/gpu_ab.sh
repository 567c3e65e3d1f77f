python -m pytest tests -m gpu -x -q 2>&1 | grep -E "mismatch|passed|failed|^FAILED|Error" | head -20
timeout 300 python scripts/prof_decode.py --chunks 16384 --iters 4 2>&1 | grep -E "iter 3|parity|rror"
timeout 600 python scripts/prof_mixed.py --mib 512 --kinds small 2>&1 | tail -1
