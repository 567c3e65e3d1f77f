set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py -x -q 2>&1 | tail -15
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --total-gib 4 --wave-gib 2 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo c5_rc=$?; tail -3 gpurun_out/bench_c5.err
python -c "import json; d=json.load(open('gpurun_out/bench_c5.json')); print(d['value'], d['ms_per_step'], d['stage_ms'], d['encode']['value'])"
