# Times the 1 GiB text encode with every tuning variant under scratch/variants/; parity tests with the default build.
for so in scratch/variants/*.so; do
  echo "== $so"
  LZB_SO=$PWD/$so timeout 300 python scripts/prof_encode.py --chunks 16384 --iters 3 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_encode.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -2
