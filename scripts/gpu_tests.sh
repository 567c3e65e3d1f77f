set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
