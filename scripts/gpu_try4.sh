set -x
LZB_SO=$PWD/scratch/liblzfse_drain.so timeout 900 python scripts/enc_words_diff.py --iters 5 2>&1 | grep -E "^run" | tail -20
timeout 600 python - <<'PY'
import ctypes as C, numpy as np, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import lzfse_rust_b200 as L, testkit as tk
enc, dec = L.LzfseEncoder(0), L.LzfseDecoder(0)
chunks = [tk.synth_text(0x5EED0000 + i, 65536) for i in range(1024)]
frames, st = enc.encode_batch(chunks)
outs, st2 = dec.decode_batch(frames)
lib = dec._lib
w = (C.c_uint32 * 32)()
lib.lzfse_b200_debug_decoder_counters.argtypes = [C.c_void_p, C.c_void_p]
print("counters", lib.lzfse_b200_debug_decoder_counters(dec._h, w), list(w))
PY
timeout 300 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv python bench.py --config 4 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_c4.log 2>&1
grep -E "k_expand_long|k_fse_lit|k_fse_lmds|k_scan" gpurun_out/launches_c4.csv | awk -F'","' '{print $5, $(NF)}' | tail -24
