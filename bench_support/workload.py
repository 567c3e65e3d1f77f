"""Synthetic workloads for bench.py and the full-size tests.

The word pool is the whitespace-split text of four public-domain fixtures (alice29, asyoulik, lcet10,
plrabn12) stored compressed under tests/golden; it is decoded with the GPU codec itself."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "libtextgen.so")
TEXT_FIXTURES = ("alice29.txt", "asyoulik.txt", "lcet10.txt", "plrabn12.txt")


def build(force=False):
    src = os.path.join(_HERE, "textgen.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-pthread", "-o", _SO, src])
    return _SO


def word_pool(decoder):
    """(pool_bytes uint8[], word_off uint32[n+1]) built from the text fixtures, decoded by `decoder`."""
    toks = []
    for name in TEXT_FIXTURES:
        frame = open(os.path.join(_ROOT, "tests", "golden", "data", "snappy", name + ".lzfse"), "rb").read()
        out = bytearray()
        decoder.decode_bytes(frame, out)
        toks.extend(bytes(out).split())
    lens = np.array([len(t) for t in toks], dtype=np.uint32)
    off = np.zeros(len(toks) + 1, dtype=np.uint32)
    np.cumsum(lens, out=off[1:])
    return np.frombuffer(b"".join(toks), dtype=np.uint8).copy(), off


def text_chunks(pool, word_off, n_chunks, chunk_len, seed0=0x5EED0000, out=None, threads=None):
    """n_chunks x chunk_len bytes of synthetic text, chunk i seeded seed0 + i.  Returns a uint8 array."""
    lib = C.CDLL(build())
    lib.textgen_chunks.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int]
    lib.textgen_chunks.restype = None
    if out is None:
        out = np.empty(n_chunks * chunk_len, dtype=np.uint8)
    threads = threads or os.cpu_count() or 1
    lib.textgen_chunks(pool.ctypes.data, word_off.ctypes.data, len(word_off) - 1, out.ctypes.data, chunk_len, n_chunks, seed0 & 0xFFFFFFFF, threads)
    return out
