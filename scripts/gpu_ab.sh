python -m pytest tests -m gpu -x -q 2>&1 | grep -E "mismatch|passed|failed|^FAILED|Error" | head -20
