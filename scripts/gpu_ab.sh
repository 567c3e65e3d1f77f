# Host-path slice tuning: e2e of the bench workload for a few slice layouts.
for sl in "4,12,28,52,76,100" "2,8,20,44,72,100" "3,10,25,50,75,100" "5,15,35,65,100" "8,24,48,74,100" "6,18,40,70,100"; do
  echo "== $sl"
  LZB_SLICES=$sl timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
done
