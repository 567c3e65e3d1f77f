"""Scratch: where does the time of encoding 262 144 LZVN-sized inputs go?  (bench_mixed's "small" class)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lzfse_rust_b200 as L
from bench_support import workload as W
import bench_mixed as BM
enc, dec = L.LzfseEncoder(0), L.LzfseDecoder(0)
enc.set_timing(True)
pool, woff = W.word_pool(dec)
raw, offs, lens = BM.make("small", 512 << 20, 1, pool, woff)
dev = torch.device("cuda:0")
t = lambda x: torch.from_numpy(np.asarray(x, np.int64)).to(dev)
d_raw = torch.from_numpy(raw).to(dev)
ulens, inv = np.unique(lens, return_inverse=True)
caps = np.array([enc.encode_bound(int(l)) for l in ulens], np.int64)[inv]
coff = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
d_comp = torch.empty(int(caps.sum()), dtype=torch.uint8, device=dev)
d_offs, d_lens, d_coff, d_caps = t(offs), t(lens), t(coff), t(caps)
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); w0 = time.time(); e0.record()
    c_len, st = enc.encode_batch_device(d_raw, d_offs, d_lens, d_comp, d_coff, d_caps)
    e1.record(); torch.cuda.synchronize(); w1 = time.time()
    print("iter %d: events %.2f ms, wall %.2f ms, stages %s" % (it, e0.elapsed_time(e1), (w1 - w0) * 1e3, {k: round(v, 2) for k, v in enc.last_stage_ms().items() if v > 0.01}))
