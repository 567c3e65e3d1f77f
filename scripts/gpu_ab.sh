python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k many_small 2>&1 | grep -E "^E|assert|Error" | head -20
