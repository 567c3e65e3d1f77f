set -x
python scripts/prof_encode.py --chunks 2048 --iters 2 > gpurun_out/prof_encode_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_enc_parse|k_enc_fse_blocks' -s 0 -c 2 -o gpurun_out/prof_encode_r1 -f python scripts/prof_encode.py --chunks 2048 --iters 1 > gpurun_out/prof_encode_ncu.log 2>&1
cat gpurun_out/prof_encode_plain.log; tail -2 gpurun_out/prof_encode_ncu.log
