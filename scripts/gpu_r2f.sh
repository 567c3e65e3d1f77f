# Round-2 final measurement pass (f): the state the round ends with: smoke, bench lines (reference, config 2, 4, 5), ncu launch lists, full captures of the
# decode kernels, the 64 KiB encode kernels and the long-stream encode kernels.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo bench_rc=$?
tail -3 gpurun_out/bench_ours.err
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo c4_rc=$?
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --total-gib 4 --wave-gib 2 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo c5_rc=$?; tail -3 gpurun_out/bench_c5.err
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_fse_literals|k_fse_lmds|^k_expand$' -s 11 -c 3 -o gpurun_out/decode_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 300 python scripts/prof_encode.py --chunks 1184 --iters 2 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_enc_find|k_long_replay|k_long_blocks|k_enc_fse_blocks' -s 4 -c 4 -o gpurun_out/enc_fast -f python scripts/prof_encode.py --chunks 1184 --iters 2 > gpurun_out/ncu_enc_fast.log 2>&1
tail -2 gpurun_out/ncu_enc_fast.log
timeout 300 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv python bench.py --config 4 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_c4.log 2>&1
timeout 300 python scripts/prof_large.py --iters 1 > gpurun_out/plain_large.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_long_' -c 12 -o gpurun_out/enc_long -f python scripts/prof_large.py --iters 1 > gpurun_out/ncu_enc_long.log 2>&1
tail -2 gpurun_out/ncu_enc_long.log
for f in ours reference c4 c5; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$f.json"))
    print("$f", d["value"], d["ms_per_step"], d.get("stage_ms"), "e2e", d["e2e"].get("value"), "enc", d.get("encode", {}).get("value"), d.get("encode", {}).get("stage_ms"))
except Exception as e:
    print("$f", "ERR", e)
PY
done
