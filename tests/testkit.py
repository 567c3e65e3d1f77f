"""Deterministic generators mirroring the reference's test_kit crate
(/root/reference/test_kit/src/rng.rs:15-58, seq.rs:25-35) plus the synthetic-text generator that
SURVEY.md §8(d) specifies for the benchmark workload."""
import glob
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def lcg_words(seed, n):
    """n successive states of x <- x*1103515245 + 12345 (mod 2^32), starting with `seed` itself
    (Rng::gen_vec emits the current state before stepping)."""
    out = np.empty(n, dtype=np.uint32)
    x = seed & 0xFFFFFFFF
    # vectorised jump-ahead: x_k = a^k x + c (a^k - 1)/(a - 1); do it in chunks with python ints
    a, c, m = 1103515245, 12345, 1 << 32
    for i in range(n):
        out[i] = x
        x = (x * a + c) % m
    return out


def rng_gen_vec(seed, length):
    """Rng::new(seed).gen_vec(length) (rng.rs:41-58)."""
    nw = length // 4
    words = lcg_words_fast(seed, nw + 1)
    b = words[:nw].astype("<u4").tobytes()
    tail = int(words[nw])
    b += bytes((tail >> (8 * i)) & 0xFF for i in range(length % 4))
    return b


def lcg_words_fast(seed, n):
    """Same sequence as lcg_words, O(n) numpy via doubling."""
    a, c = np.uint64(1103515245), np.uint64(12345)
    mask = np.uint64(0xFFFFFFFF)
    out = np.empty(max(n, 1), dtype=np.uint64)
    out[0] = seed & 0xFFFFFFFF
    filled = 1
    ak, ck = a, c  # affine map for a jump of `filled`
    while filled < n:
        take = min(filled, n - filled)
        out[filled : filled + take] = (out[:take] * ak + ck) & mask
        # compose jump: f(f(x)) => a' = a*a, c' = a*c + c
        ck = (ak * ck + ck) & mask
        ak = (ak * ak) & mask
        filled += take
    return out[:n].astype(np.uint32)


def seq_bytes(seed, length, mask=0xFFFFFFFF):
    """Seq::masked(Rng::new(seed), mask).take(length) (seq.rs:25-35): each gen() steps the LCG first."""
    nw = (length + 3) // 4
    words = lcg_words_fast(seed, nw + 1)[1:] & np.uint32(mask)
    return words.astype("<u4").tobytes()[:length]


_POOL = None


def word_pool():
    """Whitespace-split tokens of the decoded text fixtures (alice29, asyoulik, lcet10, plrabn12)."""
    global _POOL
    if _POOL is None:
        import oracle_binding as ob

        toks = []
        for name in ("alice29.txt", "asyoulik.txt", "lcet10.txt", "plrabn12.txt"):
            frame = open(os.path.join(GOLDEN, "data", "snappy", name + ".lzfse"), "rb").read()
            st, raw = ob.decode(frame)
            assert st == 0
            toks.extend(raw.split())
        _POOL = toks
    return _POOL


def synth_text(seed, length):
    """SURVEY.md §8(d) config 2 generator: pool[(x>>8) % len] + ' ', newline once a line is >= 72
    chars, truncated to `length`; x from the LCG seeded with `seed`."""
    pool = word_pool()
    n = len(pool)
    est = length // 4 + 64
    words = lcg_words_fast(seed, est + 1)[1:]
    out = bytearray()
    line = 0
    i = 0
    while len(out) < length:
        if i >= len(words):
            words = np.concatenate([words, lcg_words_fast(int(words[-1]), est + 1)[1:]])
        w = pool[(int(words[i]) >> 8) % n]
        i += 1
        out += w
        line += len(w) + 1
        if line >= 72:
            out += b"\n"
            line = 0
        else:
            out += b" "
    return bytes(out[:length])


def golden_frames(*dirs):
    """[(name, frame_bytes, sha256_or_None)] for the committed reference fixtures."""
    res = []
    for d in dirs or ("snappy", "mutate", "special"):
        for f in sorted(glob.glob(os.path.join(GOLDEN, "data", d, "*.lzfse"))):
            h = f[:-6] + ".hash"
            res.append((d + "/" + os.path.basename(f)[:-6], open(f, "rb").read(), open(h, "rb").read() if os.path.exists(h) else None))
    return res


def sha256(b):
    return hashlib.sha256(b).digest()
