# Scratch A/B script for `gpurun -- 'bash scripts/gpu_ab.sh'`: tuning builds under scratch/variants/ (see build_variants.sh)
# against the 1 GiB text decode.  Edit freely; the measurements that count are in profiles/.
for so in scratch/variants/*.so; do
  echo "== $so"
  LZB_SO=$PWD/$so timeout 300 python scripts/prof_decode.py --chunks 16384 --iters 4 2>&1 | grep -E "iter 3|parity|rror"
done
