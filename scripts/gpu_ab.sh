for so in scratch/variants/vn_c2.so scratch/variants/vn_c3.so; do
  echo "== $so"; LZB_SO=$PWD/$so timeout 600 python scripts/prof_mixed.py --mib 512 --kinds small 2>&1 | tail -1 | grep -o "decode.*"
done
