# Quick A/B pass: decode tests, then bench config 2 with and without the change under test.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_hostile.py tests/test_gpu_configs.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/try_a.json 2> gpurun_out/try_a.err; echo rc=$?
python -c "import json; d=json.load(open('gpurun_out/try_a.json')); print(d['value'], d['ms_per_step'], d['stage_ms'])"
LZB_LIT4=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/try_b.json 2> gpurun_out/try_b.err; echo rc=$?
python -c "import json; d=json.load(open('gpurun_out/try_b.json')); print(d['value'], d['ms_per_step'], d['stage_ms'])"
