set -x
python scripts/prof_decode.py --chunks 1024 --iters 2 > gpurun_out/prof_decode_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_fse_lmds|k_fse_literals|k_expand' -s 12 -c 3 -o gpurun_out/prof_decode_r1 -f python scripts/prof_decode.py --chunks 1024 --iters 1 > gpurun_out/prof_decode_ncu2.log 2>&1
tail -3 gpurun_out/prof_decode_ncu2.log
ls -la gpurun_out/
