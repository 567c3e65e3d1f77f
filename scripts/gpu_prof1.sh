set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/prof_decode.py --chunks 4096 --iters 4 > gpurun_out/prof_decode_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_launches.csv python scripts/prof_decode.py --chunks 4096 --iters 2 > gpurun_out/prof_decode_ncu.log 2>&1
cat gpurun_out/prof_decode_plain.log
grep -E 'k_' gpurun_out/decode_launches.csv | awk -F'","' '{print $5, $NF}' | tail -7
