"""Encodes the config-2 batch several times on the GPU and reports every stream whose frame changes between runs or
differs from the oracle encoder's (all streams, not a sample): python scripts/enc_determinism.py [--chunks N] [--iters K]"""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lzfse_rust_b200 as L
from bench_support import workload as W
import bench
ap = argparse.ArgumentParser(); ap.add_argument("--chunks", type=int, default=16384); ap.add_argument("--iters", type=int, default=6)
a = ap.parse_args()
dec, enc = L.LzfseDecoder(0), L.LzfseEncoder(0)
pool, woff = W.word_pool(dec)
n, cl = a.chunks, 65536
raw_np = W.text_chunks(pool, woff, n, cl, seed0=0x5EED0000)
raw = torch.from_numpy(raw_np).cuda()
i64 = lambda x: torch.tensor(np.asarray(x, dtype=np.int64), device="cuda")
bound = enc.encode_bound(cl)
offs, lens = np.arange(n, dtype=np.int64) * cl, np.full(n, cl, np.int64)
_, _, _, comp_o, c_off_o, c_len_o = bench.cpu_codec_bench(raw_np, offs, lens, os.cpu_count() or 1)
runs = []
for it in range(a.iters):
    comp = torch.empty(n * bound, dtype=torch.uint8, device="cuda")
    c_len, st = enc.encode_batch_device(raw, i64(offs), i64(lens), comp, i64(np.arange(n) * bound), i64(np.full(n, bound)))
    assert int((st != 0).sum()) == 0
    runs.append((comp.cpu().numpy(), c_len.cpu().numpy().astype(np.int64)))
    bad = []
    for i in range(n):
        o = comp_o[int(c_off_o[i]):int(c_off_o[i]) + int(c_len_o[i])]
        g = runs[-1][0][i * bound:i * bound + int(runs[-1][1][i])]
        if len(o) != len(g) or not np.array_equal(o, g):
            k = int(np.argmax(o[:min(len(o), len(g))] != g[:min(len(o), len(g))])) if len(o) and len(g) else -1
            bad.append((i, len(o), len(g), k))
    print("run", it, "frames differing from the oracle:", len(bad), bad[:8], flush=True)
