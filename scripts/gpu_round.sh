# Round measurement pass: tests, smoke, bench (ours + reference), ncu launch list, full capture of the decode kernels.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo bench_rc=$?
tail -3 gpurun_out/bench_ours.err
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_fse_literals|k_fse_lmds|^k_expand$' -s 11 -c 3 -o gpurun_out/decode_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
cat gpurun_out/bench_ours.json
# mixed corpus (BASELINE configs[4]) on one GPU, oracle port timed on a 64 MiB sample per class
timeout 1200 python scripts/bench_mixed.py --mib-per-class 1024 --waves 1 --cpu-mib 64 > gpurun_out/bench_mixed_1gpu.json 2> gpurun_out/bench_mixed_1gpu.err; echo mixed_rc=$?
cut -c1-300 gpurun_out/bench_mixed_1gpu.json
# BASELINE configs[3]: 8 streams of 16 MiB
timeout 300 python scripts/prof_large.py 2>&1 | tail -5 > gpurun_out/prof_large.log; cat gpurun_out/prof_large.log
