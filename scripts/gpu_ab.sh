for pad in 0 56; do
  echo "== pad_kb $pad"
  LZB_EXPAND_PAD_KB=$pad timeout 300 python scripts/prof_decode.py --chunks 16384 --iters 4 2>&1 | grep -E "iter 3|rror"
done
