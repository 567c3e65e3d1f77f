"""Measures PCIe copy rates and the host decode entry point: python scripts/prof_e2e.py"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lzfse_rust_b200 as L
from bench_support import workload as W
dec, enc = L.LzfseDecoder(0), L.LzfseEncoder(0)
if os.environ.get("LZB_TIMING"): dec.set_timing(True)
pool, woff = W.word_pool(dec)
n, cl = 16384, 65536
raw_h = torch.empty(n * cl, dtype=torch.uint8).pin_memory()
W.text_chunks(pool, woff, n, cl, out=raw_h.numpy())
d = raw_h.cuda()
def t(f, reps=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
half = raw_h[: n * cl // 2]
dh = torch.empty_like(half, device="cuda")
print("H2D 512 MiB pinned: %.2f ms (%.1f GB/s)" % (1e3 * t(lambda: dh.copy_(half, non_blocking=True)), half.numel() / t(lambda: dh.copy_(half, non_blocking=True)) / 1e9))
back = torch.empty(n * cl, dtype=torch.uint8).pin_memory()
print("D2H 1 GiB pinned: %.2f ms (%.1f GB/s)" % (1e3 * t(lambda: back.copy_(d, non_blocking=True)), d.numel() / t(lambda: back.copy_(d, non_blocking=True)) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): dh.copy_(half, non_blocking=True)
    with torch.cuda.stream(s2): back.copy_(d, non_blocking=True)
print("H2D 512 MiB + D2H 1 GiB concurrently: %.2f ms" % (1e3 * t(both)))
# frames
bound = enc.encode_bound(cl)
comp = np.empty(n * bound, np.uint8)
c_len, st = enc.encode_batch_into(raw_h.numpy(), np.arange(n) * cl, np.full(n, cl), comp, np.arange(n) * bound, np.full(n, bound))
offs = np.concatenate([[0], np.cumsum(c_len)[:-1]]).astype(np.uint64)
packed = torch.empty(int(c_len.sum()), dtype=torch.uint8).pin_memory()
pk = packed.numpy()
for i in range(n): pk[int(offs[i]):int(offs[i]) + int(c_len[i])] = comp[i * bound:i * bound + int(c_len[i])]
out = torch.empty(n * cl, dtype=torch.uint8).pin_memory()
args = (pk, offs, c_len, out.numpy(), np.arange(n) * cl, np.full(n, cl))
for it in range(4):
    t0 = time.perf_counter(); ol, s = dec.decode_batch_into(*args); dt = time.perf_counter() - t0
    print("decode_batch_host: %.2f ms  launches %d" % (dt * 1e3, dec.last_launches))
assert np.array_equal(out.numpy(), raw_h.numpy())
