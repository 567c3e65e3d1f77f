# Times the 1 GiB text decode with every tuning variant under scratch/variants/.
for so in scratch/variants/*.so; do
  echo "== $so"
  LZB_SO=$PWD/$so timeout 300 python scripts/prof_decode.py --chunks 16384 --iters 3 2>&1 | grep -E "iter 2|parity|rror"
done
