# Full ncu capture of the fast-parse kernels on a small batch (8 streams per SM).
set -x
timeout 300 python scripts/prof_encode.py --chunks 1184 --iters 2 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_enc_find|k_enc_replay|k_enc_fse_blocks' -s 3 -c 3 -o gpurun_out/enc_fast -f python scripts/prof_encode.py --chunks 1184 --iters 2 > gpurun_out/ncu_enc_fast.log 2>&1
tail -2 gpurun_out/ncu_enc_fast.log
