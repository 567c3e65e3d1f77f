mkdir -p gpurun_out
for v in lzfse_rust_b200/liblzfse_b200.so gpurun_tmp_r8192.so gpurun_tmp_r4096.so; do
echo "== $v"
LZB_SO=$PWD/$v timeout 300 python scripts/try_long.py > gpurun_out/try_long.log 2>&1; grep -v "^ok " gpurun_out/try_long.log | tail -5 | cut -c1-400
LZB_SO=$PWD/$v timeout 300 python scripts/prof_encode.py --chunks 16384 --iters 3 2>&1 | tail -1 | cut -c1-600
done
timeout 900 python -m pytest tests/test_gpu_encode.py tests/test_gpu_configs.py -x -q 2>&1 | tail -5
