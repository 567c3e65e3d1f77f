mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo c4_rc=$?
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo bench_rc=$?
for f in ours c4; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$f.json"))
    print("$f", d["value"], d["ms_per_step"], d.get("stage_ms"), "e2e", d["e2e"].get("value"), "enc", d.get("encode", {}).get("value"), d.get("encode", {}).get("stage_ms"), d["cpu_baseline"])
except Exception as e:
    print("$f", "ERR", e)
PY
done
