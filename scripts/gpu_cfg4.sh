# BASELINE configs[3] (8 streams of 16 MiB): encode + decode timing with the kernel the library picks, and with a warp per stream.
set -x
timeout 300 python scripts/prof_large.py 2>&1 | tail -5
LZB_EXPAND=warp timeout 300 python scripts/prof_large.py --iters 2 2>&1 | tail -2
