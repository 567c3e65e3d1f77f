"""Encode timing helper: python scripts/prof_encode.py [--chunks N] [--iters K]"""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lzfse_rust_b200 as L
from bench_support import workload as W
ap = argparse.ArgumentParser(); ap.add_argument("--chunks", type=int, default=2048); ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dec, enc = L.LzfseDecoder(0), L.LzfseEncoder(0)
pool, woff = W.word_pool(dec)
n, cl = a.chunks, 65536
raw = torch.from_numpy(W.text_chunks(pool, woff, n, cl)).cuda()
i64 = lambda x: torch.tensor(np.asarray(x, dtype=np.int64), device="cuda")
bound = enc.encode_bound(cl)
comp = torch.empty(n * bound, dtype=torch.uint8, device="cuda")
enc.set_timing(True)
for it in range(a.iters):
    c_len, st = enc.encode_batch_device(raw, i64(np.arange(n) * cl), i64(np.full(n, cl)), comp, i64(np.arange(n) * bound), i64(np.full(n, bound)))
    print(it, enc.last_stage_ms(), "ratio %.3f" % (n * cl / int(c_len.sum())))
assert int((st != 0).sum()) == 0
