# Round-2 measurement pass: tests, smoke, bench (reference + ours), ncu launch list.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo bench_rc=$?
tail -3 gpurun_out/bench_ours.err
cat gpurun_out/bench_ours.json
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo c4_rc=$?
cat gpurun_out/bench_c4.json
