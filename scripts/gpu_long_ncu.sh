set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_long_find|k_long_chain|k_long_info|k_long_replay|k_long_blocks' -c 5 -o gpurun_out/enc_long -f python scripts/prof_large.py --iters 1 > gpurun_out/ncu_enc_long.log 2>&1
tail -3 gpurun_out/ncu_enc_long.log
