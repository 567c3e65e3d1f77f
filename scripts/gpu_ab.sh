python -m pytest tests -m gpu -x -q 2>&1 | grep -E "mismatch|passed|failed|^FAILED|Error" | head -20
timeout 600 python scripts/prof_mixed.py --mib 512 --kinds small 2>&1 | tail -1
