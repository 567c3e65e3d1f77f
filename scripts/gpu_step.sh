mkdir -p gpurun_out
for v in lzfse_rust_b200/liblzfse_b200.so gpurun_tmp_m6.so gpurun_tmp_m8.so; do
echo "== $v"
LZB_SO=$PWD/$v timeout 300 python scripts/try_long.py > gpurun_out/try_long.log 2>&1; grep -v "^ok " gpurun_out/try_long.log | tail -4 | cut -c1-400
done
