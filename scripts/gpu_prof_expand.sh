# Full ncu capture of the expansion kernel on a 256 MiB slice of the text workload (skips the 4 single-stream launches
# that build the word pool).  LZB_SO selects a tuning variant.
set -x
timeout 600 ncu --set full --clock-control none --import-source on --warp-sampling-interval 1 -k regex:k_expand --launch-skip 4 -c 1 -o gpurun_out/expand_full -f \
  python scripts/prof_decode.py --chunks 4096 --iters 1 > gpurun_out/expand_full.log 2>&1
tail -3 gpurun_out/expand_full.log
