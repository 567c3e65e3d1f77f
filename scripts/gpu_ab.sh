for so in scratch/variants/*.so; do
  echo "== $so"
  LZB_SO=$PWD/$so timeout 120 python scripts/prof_encode.py --chunks 16384 --iters 2 2>&1 | tail -1 | cut -c1-250
done
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q 2>&1 | tail -2
