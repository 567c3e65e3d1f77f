for kb in 0 56 75 110 200; do
echo "smem_kb=$kb"; LZB_EXPAND_SMEM_KB=$kb python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('decode', d['value'], 'GB/s', d['stage_ms']['expand'])"
done
