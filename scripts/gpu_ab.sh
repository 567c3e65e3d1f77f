for so in scratch/variants/pf256.so scratch/variants/pf384.so; do
  echo "== $so"; LZB_SO=$PWD/$so timeout 300 python scripts/prof_decode.py --chunks 16384 --iters 4 2>&1 | grep -E "iter 3|rror"
done
