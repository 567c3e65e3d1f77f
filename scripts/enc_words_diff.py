"""Runs the GPU encoder twice on the config-2 batch and compares the per-position words of k_enc_find between the runs."""
import argparse, ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lzfse_rust_b200 as L
from bench_support import workload as W
ap = argparse.ArgumentParser(); ap.add_argument("--chunks", type=int, default=16384); ap.add_argument("--iters", type=int, default=4)
a = ap.parse_args()
dec, enc = L.LzfseDecoder(0), L.LzfseEncoder(0)
pool, woff = W.word_pool(dec)
n, cl = a.chunks, 65536
raw_np = W.text_chunks(pool, woff, n, cl, seed0=0x5EED0000)
raw = torch.from_numpy(raw_np).cuda()
i64 = lambda x: torch.tensor(np.asarray(x, dtype=np.int64), device="cuda")
bound = enc.encode_bound(cl)
offs, lens = np.arange(n, dtype=np.int64) * cl, np.full(n, cl, np.int64)
lib = enc._lib
lib.lzfse_b200_debug_encoder_words.restype = C.c_size_t
lib.lzfse_b200_debug_encoder_words.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
lib.lzfse_b200_debug_encoder_packs.restype = C.c_size_t
lib.lzfse_b200_debug_encoder_packs.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
ref = None
refp = None
reff = None
for it in range(a.iters):
    comp = torch.empty(n * bound, dtype=torch.uint8, device="cuda")
    c_len, st = enc.encode_batch_device(raw, i64(offs), i64(lens), comp, i64(np.arange(n) * bound), i64(np.full(n, bound)))
    w = np.empty(n * cl, np.uint32)
    got = lib.lzfse_b200_debug_encoder_words(enc._h, w.ctypes.data, w.size)
    assert got >= n * cl, got
    pk = np.zeros(n * 24000, np.uint64)
    gotp = lib.lzfse_b200_debug_encoder_packs(enc._h, pk.ctypes.data, pk.size)
    fr = comp.cpu().numpy(); cl_h = c_len.cpu().numpy()
    if ref is None:
        ref = w; refp = pk; reff = (fr, cl_h); continue
    dp = np.nonzero(refp[:gotp] != pk[:gotp])[0]
    print("run", it, "pack words differing from run 0:", len(dp), dp[:10], "of", gotp, flush=True)
    for x in dp[:6]:
        print("    pack", int(x), hex(int(refp[x])), hex(int(pk[x])))
    dl = np.nonzero(reff[1] != cl_h)[0]
    dfr = [i for i in range(n) if not np.array_equal(reff[0][i * bound:i * bound + int(reff[1][i])], fr[i * bound:i * bound + int(cl_h[i])])]
    print("run", it, "frames differing from run 0:", len(dfr), dfr[:10], "length diffs", len(dl), flush=True)
    d = np.nonzero(ref != w)[0]
    print("run", it, "words differing from run 0:", len(d), flush=True)
    for x in d[:12]:
        s_, p_ = divmod(int(x), cl)
        a_, b_ = int(ref[x]), int(w[x])
        f = lambda v: "D=%d len=%d bw=%d" % (v & 0x3FFFF, (v >> 18) & 0x3FF, v >> 28)
        print("   stream %d pos %d: run0 %s | now %s" % (s_, p_, f(a_), f(b_)))
        # what does the text look like
        src = raw_np[s_ * cl:(s_ + 1) * cl]
        for v in (a_, b_):
            D = v & 0x3FFFF
            if D:
                ln = 0
                while p_ + ln < cl and src[p_ + ln] == src[p_ - D + ln]: ln += 1
                print("      D=%d true forward length %d" % (D, ln))
