# Full ncu capture (source counters + warp-state sampling) of the three decode kernels of one 1 GiB step.
set -x
python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_fse_literals|k_fse_lmds|^k_expand$' -s 11 -c 3 -o gpurun_out/decode_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-400
