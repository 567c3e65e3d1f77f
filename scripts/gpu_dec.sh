python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('decode', d['value'], 'GB/s', d['ms_per_step'], 'ms', d['stage_ms']); print('e2e', d['e2e'])"
