"""Interim timing of BASELINE.json configs[3]: a few 16 MiB streams (multi-block frames, cross-block matches).
Usage: python scripts/prof_large.py [--streams N] [--mib M] [--iters K]   (LZB_EXPAND=warp|cta selects the expansion kernel)"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lzfse_rust_b200 as L  # noqa: E402
from bench_support import workload as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=8)
ap.add_argument("--mib", type=int, default=16)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()

enc, dec = L.LzfseEncoder(0), L.LzfseDecoder(0)
dec.set_timing(True)
enc.set_timing(True)
pool, woff = W.word_pool(dec)
n, cl = a.streams, a.mib << 20
streams = [W.text_chunks(pool, woff, 1, cl, seed0=0x16000000 + i).tobytes() for i in range(n)]
frames, st = enc.encode_batch(streams)
assert not st.any()
print("encode stages", {k: round(v, 2) for k, v in enc.last_stage_ms().items()}, "ratio %.3f" % (n * cl / sum(len(f) for f in frames)))
c_len = np.array([len(f) for f in frames], np.int64)
offs = np.concatenate([[0], np.cumsum(c_len)[:-1]]).astype(np.int64)
dev = torch.device("cuda:0")
d_src = torch.from_numpy(np.frombuffer(b"".join(frames), np.uint8).copy()).to(dev)
d_soff = torch.from_numpy(offs).to(dev); d_slen = torch.from_numpy(c_len).to(dev)
d_dst = torch.zeros(n * cl, dtype=torch.uint8, device=dev)
d_doff = torch.from_numpy((np.arange(n) * cl).astype(np.int64)).to(dev); d_dcap = torch.full((n,), cl, dtype=torch.int64, device=dev)
for it in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    out_len, status = dec.decode_batch_device(d_src, d_soff, d_slen, d_dst, d_doff, d_dcap)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("iter %d: %.3f ms  %.2f GB/s uncompressed  stages %s" % (it, ms, n * cl / ms / 1e6, {k: round(v, 3) for k, v in dec.last_stage_ms().items()}))
assert int((status != 0).sum()) == 0
assert bytes(d_dst.cpu().numpy()) == b"".join(streams)
print("parity ok")
