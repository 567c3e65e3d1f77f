# Round-2 closing pass (e): full GPU suite, smoke, bench lines (reference, config 2, 4, 5), per-class stage times of the mixed corpus.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo bench_rc=$?
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo c4_rc=$?
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --total-gib 4 --wave-gib 2 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo c5_rc=$?
timeout 600 python scripts/bench_mixed.py --mib-per-class 512 2>&1 | grep "encode stages" | cut -c1-300
for f in ours reference c4 c5; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$f.json"))
    print("$f", d["value"], d["ms_per_step"], d.get("stage_ms"), "e2e", d["e2e"].get("value"), "enc", d.get("encode", {}).get("value"), d.get("encode", {}).get("stage_ms"))
except Exception as e:
    print("$f", "ERR", e)
PY
done
