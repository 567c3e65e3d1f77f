set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo rc=$?
tail -5 gpurun_out/bench_r1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1.json'))
print('decode', d['value'], d['ms_per_step'], d['stage_ms']); print('e2e', d['e2e']['value']); print('encode', d['encode']); print('clocks', d['clocks'])"
