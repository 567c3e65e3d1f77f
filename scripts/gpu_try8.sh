set -x
mkdir -p gpurun_out
timeout 900 python scripts/enc_determinism.py --iters 3 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_encode.py tests/test_gpu_configs.py tests/test_gpu_api.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/try_a.json 2> gpurun_out/try_a.err; echo rc=$?; tail -3 gpurun_out/try_a.err
python -c "import json; d=json.load(open('gpurun_out/try_a.json')); print(d['value'], d['encode']['value'], d['encode']['stage_ms'])"
