# Full ncu capture of the encode parse kernel on 4096 text chunks (256 MiB).
set -x
timeout 900 ncu --set full --clock-control none --import-source on --warp-sampling-interval 2 -k regex:k_enc_parse --launch-skip 0 -c 1 -o gpurun_out/parse_full -f \
  python scripts/prof_encode.py --chunks 4096 --iters 1 > gpurun_out/parse_full.log 2>&1
tail -3 gpurun_out/parse_full.log
