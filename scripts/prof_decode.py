"""Interim decode timing: synthetic text chunks -> oracle-encoded frames -> GPU batched decode.
Usage: python scripts/prof_decode.py [--chunks N] [--iters K]"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lzfse_rust_b200 as L  # noqa: E402
import oracle_binding as ob  # noqa: E402
from bench_support import workload as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=4096)
ap.add_argument("--chunk-len", type=int, default=65536)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()

dec = L.LzfseDecoder(0)
dec.set_timing(True)
pool, woff = W.word_pool(dec)
n, cl = a.chunks, a.chunk_len
raw = W.text_chunks(pool, woff, n, cl)
u64 = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
src_off = u64(np.arange(n) * cl); src_len = u64(np.full(n, cl))
bound = ob.lib().orc_encode_bound(cl)
comp = np.empty(n * bound, dtype=np.uint8)
c_off = u64(np.arange(n) * bound); c_cap = u64(np.full(n, bound)); c_len = np.zeros(n, np.uint64); st = np.zeros(n, np.int32)
p = lambda x: x.ctypes.data_as(C.c_void_p)
t = time.time()
ob.lib().orc_encode_batch(p(raw), src_off.ctypes.data_as(C.POINTER(C.c_uint64)), src_len.ctypes.data_as(C.POINTER(C.c_uint64)), p(comp),
                          c_off.ctypes.data_as(C.POINTER(C.c_uint64)), c_cap.ctypes.data_as(C.POINTER(C.c_uint64)),
                          c_len.ctypes.data_as(C.POINTER(C.c_uint64)), st.ctypes.data_as(C.POINTER(C.c_int32)), n, os.cpu_count())
print("oracle encode: %.2fs, ratio %.3f" % (time.time() - t, n * cl / c_len.sum()))
# pack frames tightly
offs = np.concatenate([[0], np.cumsum(c_len)[:-1]]).astype(np.int64)
packed = np.empty(int(c_len.sum()), np.uint8)
for i in range(n):
    packed[offs[i]:offs[i] + int(c_len[i])] = comp[i * bound:i * bound + int(c_len[i])]
dev = torch.device("cuda:0")
d_src = torch.from_numpy(packed).to(dev)
d_soff = torch.from_numpy(offs).to(dev); d_slen = torch.from_numpy(c_len.astype(np.int64)).to(dev)
d_dst = torch.zeros(n * cl, dtype=torch.uint8, device=dev)
d_doff = torch.from_numpy((np.arange(n) * cl).astype(np.int64)).to(dev); d_dcap = torch.full((n,), cl, dtype=torch.int64, device=dev)
for it in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    out_len, status = dec.decode_batch_device(d_src, d_soff, d_slen, d_dst, d_doff, d_dcap)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("iter %d: %.3f ms  %.1f GB/s uncompressed, %.1f GB/s (U+C)  stages %s" % (it, ms, n * cl / ms / 1e6, (n * cl + int(c_len.sum())) / ms / 1e6, {k: round(v, 3) for k, v in dec.last_stage_ms().items()}))
assert int((status != 0).sum()) == 0
assert bytes(d_dst.cpu().numpy()) == raw.tobytes()
print("parity ok; launches", dec.last_launches)
