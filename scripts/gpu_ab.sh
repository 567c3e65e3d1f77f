# ncu capture of the single-LZVN-block expansion kernel on the small-chunk class.
timeout 600 python scripts/prof_mixed.py --mib 256 --kinds small 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand_vn -s 1 -c 1 -o gpurun_out/vn_full -f python scripts/prof_mixed.py --mib 256 --kinds small > gpurun_out/vn_full.log 2>&1
tail -2 gpurun_out/vn_full.log
