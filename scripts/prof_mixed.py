"""Per-class timing of the BASELINE.json configs[4] mix (incompressible / LZVN-sized / repetitive / text), one class per batch.
Usage: python scripts/prof_mixed.py [--mib M]   (M = MiB of input per class)"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lzfse_rust_b200 as L  # noqa: E402
from bench_support import workload as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=128)
ap.add_argument("--kinds", default="text,noise,small,rep")
a = ap.parse_args()
enc, dec = L.LzfseEncoder(0), L.LzfseDecoder(0)
enc.set_timing(True); dec.set_timing(True)
pool, woff = W.word_pool(dec)
total = a.mib << 20
rng = np.random.default_rng(7)


def lcg_bytes(seed, n):
    x = (np.arange(n // 4 + 1, dtype=np.uint64) * 2654435761 + seed * 40503) & 0xFFFFFFFF
    x = (x * 1103515245 + 12345) & 0xFFFFFFFF
    x ^= x >> 13
    return (x.astype(np.uint32)).view(np.uint8)[:n]


def make(kind):
    if kind == "text":
        n = total // 65536
        raw = W.text_chunks(pool, woff, n, 65536)
        lens = np.full(n, 65536, np.int64)
    elif kind == "noise":
        n = total // 65536
        raw = rng.integers(0, 256, total, dtype=np.uint8)
        lens = np.full(n, 65536, np.int64)
    elif kind == "small":
        lens = (21 + rng.integers(0, 4076, total // 2048)).astype(np.int64)
        src = W.text_chunks(pool, woff, 1, int(lens.sum()) + 8, seed0=0x1234)
        raw = src[: int(lens.sum())]
    else:  # repetitive: period-p repeats
        n = total // 65536
        raw = np.empty(total, np.uint8)
        for i in range(n):
            p = [1, 2, 3, 4, 5, 7, 8, 13, 16, 32, 64][i % 11]
            raw[i * 65536:(i + 1) * 65536] = np.resize(rng.integers(0, 256, p, dtype=np.uint8), 65536)
        lens = np.full(n, 65536, np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    return np.ascontiguousarray(raw), offs, lens


dev = torch.device("cuda:0")
for kind in a.kinds.split(","):
    raw, offs, lens = make(kind)
    n = len(lens)
    d_raw = torch.from_numpy(raw).to(dev)
    bounds = np.array([enc.encode_bound(int(l)) for l in np.unique(lens)])
    bmap = dict(zip(np.unique(lens).tolist(), bounds.tolist()))
    caps = np.array([bmap[int(l)] for l in lens], np.int64)
    coff = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
    d_comp = torch.empty(int(caps.sum()), dtype=torch.uint8, device=dev)
    t = lambda x: torch.from_numpy(np.asarray(x, np.int64)).to(dev)
    for it in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        c_len, st = enc.encode_batch_device(d_raw, t(offs), t(lens), d_comp, t(coff), t(caps))
        e1.record(); torch.cuda.synchronize()
        enc_ms = e0.elapsed_time(e1)
    assert int((st != 0).sum()) == 0
    d_out = torch.zeros(len(raw), dtype=torch.uint8, device=dev)
    for it in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        out_len, dst = dec.decode_batch_device(d_comp, t(coff), c_len.to(torch.int64), d_out, t(offs), t(lens))
        e1.record(); torch.cuda.synchronize()
        dec_ms = e0.elapsed_time(e1)
    assert int((dst != 0).sum()) == 0 and torch.equal(d_out, d_raw)
    U = len(raw)
    print("%-6s n=%6d ratio %.3f  encode %.2f GB/s (%.1f ms %s)  decode %.2f GB/s (%.2f ms %s)" % (
        kind, n, U / int(c_len.sum()), U / enc_ms / 1e6, enc_ms, {k: round(v, 1) for k, v in enc.last_stage_ms().items()},
        U / dec_ms / 1e6, dec_ms, {k: round(v, 2) for k, v in dec.last_stage_ms().items()}))
