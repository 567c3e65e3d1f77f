"""Scratch check of the long-stream encoder (encode_long.cuh): frames of assorted streams > 64 KiB against the oracle's,
then the timing of 8 x 16 MiB text.  python scripts/try_long.py [--no-time]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lzfse_rust_b200 as L  # noqa: E402
import oracle_binding as ob  # noqa: E402
import testkit as tk  # noqa: E402
from bench_support import workload as W  # noqa: E402


def assorted():
    rng = np.random.default_rng(1)
    t = tk.synth_text(9, 1 << 20)
    out = bytearray()
    while len(out) < (1 << 20):
        k = rng.integers(0, 5); n = int(rng.integers(1, 3000))
        if k == 0: out += t[int(rng.integers(0, len(t) - n * 10)):][:n * 10]
        elif k == 1: out += bytes([int(rng.integers(0, 256))]) * n
        elif k == 2: out += rng.integers(0, 256, n * 10, dtype=np.uint8).tobytes()
        elif k == 3: out += (rng.integers(0, 4, n * 10, dtype=np.uint8) * 16).tobytes()
        else:
            d = int(rng.integers(1, min(len(out), 300000) + 1)) if out else 1
            for _ in range(n): out.append(out[-d] if len(out) >= d else 0)
    r = bytearray(rng.integers(0, 256, 1 << 20, dtype=np.uint8).tobytes())
    for i in range(2000):
        a = int(rng.integers(100000, len(r) - 100)); d = int(rng.integers(8, 90000)); n = int(rng.integers(4, 30))
        r[a:a + n] = r[a - d:a - d + n]
    return {"text200k": t[:200000], "text1m": t, "text65537": t[:65537], "text70001": t[5:70006], "zeros300k": bytes(300000), "rand300k": tk.rng_gen_vec(3, 300000),
            "period1000": (tk.rng_gen_vec(5, 1000) * 400)[:333333], "period3": (b"abc" * 50000)[:140001], "patch": bytes(out), "sparse": bytes(r),
            "far": (tk.rng_gen_vec(5, 300000) * 4)[:1 << 20], "lits": tk.seq_bytes(1, 500000, 0x0F0F0F0F)}


enc = L.LzfseEncoder(0)
oenc = ob.Encoder()
enc.set_timing(True)
bad = 0
data = assorted()
names = list(data)
frames, st = enc.encode_batch([data[k] for k in names])
print("stages", {k: round(v, 3) for k, v in enc.last_stage_ms().items()})
for k, f, s in zip(names, frames, st):
    ref = oenc.encode(data[k])[1]
    ok = s == 0 and f == ref
    if not ok:
        bad += 1
        i = next((i for i in range(min(len(f), len(ref))) if f[i] != ref[i]), min(len(f), len(ref)))
        print("MISMATCH %s: status %d, %d vs %d bytes, first difference at %d" % (k, s, len(f), len(ref), i))
    else:
        print("ok %s (%d -> %d)" % (k, len(data[k]), len(f)))
if "--no-time" not in sys.argv:
    dec = L.LzfseDecoder(0)
    pool, woff = W.word_pool(dec)
    streams = [W.text_chunks(pool, woff, 1, 16 << 20, seed0=0x16000000 + i).tobytes() for i in range(8)]
    for it in range(3):
        t0 = time.time()
        frames, st = enc.encode_batch(streams)
        print("8 x 16 MiB: wall %.1f ms, stages" % ((time.time() - t0) * 1e3), {k: round(v, 2) for k, v in enc.last_stage_ms().items()})
    ref = oenc.encode(streams[0])[1]
    print("16 MiB frame 0", "ok" if frames[0] == ref else "MISMATCH", len(frames[0]), len(ref))
    bad += frames[0] != ref
    outs, dst = dec.decode_batch(frames)
    print("round trip", "ok" if outs == streams and not dst.any() else "FAILED")
    bad += outs != streams
print("BAD" if bad else "ALL OK")
