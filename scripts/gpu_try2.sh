set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/try_a.json 2> gpurun_out/try_a.err; echo rc=$?
python -c "import json; d=json.load(open('gpurun_out/try_a.json')); print(d['value'], d['ms_per_step'], d['stage_ms'])"
