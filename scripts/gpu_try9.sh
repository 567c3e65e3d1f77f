set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/try_a.json 2> gpurun_out/try_a.err; echo rc=$?; tail -3 gpurun_out/try_a.err
python -c "import json; d=json.load(open('gpurun_out/try_a.json')); print(d['value'], d['ms_per_step'], d['stage_ms'], d['encode']['value'])"
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/try_c4.json 2> gpurun_out/try_c4.err; echo rc=$?
python -c "import json; d=json.load(open('gpurun_out/try_c4.json')); print(d['value'], d['ms_per_step'], d['stage_ms'])"
