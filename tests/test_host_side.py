"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, status codes agree between the product header and the oracle, the package fails loudly
without a GPU, and the multi-GPU sharding logic (one process per GPU, gloo here) partitions streams
without loss.  No compute kernels run in this file."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from lzfse_rust_b200 import _ffi, build

    so = build.build()
    lib = ctypes.CDLL(so)
    header = open(os.path.join(ROOT, "include", "lzfse_b200.h")).read()
    declared = set(re.findall(r"\b(lzfse_b200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)
    for sym in declared:
        getattr(lib, sym)
    lib.lzfse_b200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.lzfse_b200_version()
    lib.lzfse_b200_encode_bound.restype = ctypes.c_size_t
    lib.lzfse_b200_encode_bound.argtypes = [ctypes.c_size_t]
    import oracle_binding as ob

    for n in (0, 1, 20, 4096, 65536, 1 << 24):
        assert lib.lzfse_b200_encode_bound(n) == ob.lib().orc_encode_bound(n)


def test_sass_is_sm100a_only():
    so = os.path.join(ROOT, "lzfse_rust_b200", "liblzfse_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_status_codes_match_oracle_header():
    h1 = open(os.path.join(ROOT, "include", "lzfse_b200.h")).read()
    h2 = open(os.path.join(ROOT, "oracle", "lzfse_oracle.h")).read()
    a = {k: int(v) for k, v in re.findall(r"LZFSE_B200_([A-Z_]+) = (\d+)", h1)}
    b = {k: int(v) for k, v in re.findall(r"ORC_([A-Z_]+) = (\d+)", h2)}
    assert len(b) > 15
    for k, v in b.items():
        assert a[k] == v, k


def test_no_cpu_fallback():
    import torch

    import lzfse_rust_b200 as L

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.LzfseError) as ei:
        L.LzfseDecoder(0)
    assert ei.value.status == 65  # NoDevice
    with pytest.raises(L.LzfseError):
        L.encode_bytes(b"test", bytearray())
    # the product never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "lzfse_rust_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(root, f), errors="ignore").read().lower(), f


def test_shard_ranges_cover_and_balance():
    from lzfse_rust_b200 import shard_ranges

    rng = np.random.default_rng(3)
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 100, 16384):
            w = rng.integers(1, 200000, n)
            r = shard_ranges(w, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            if n >= 100 * world:
                loads = [w[a:b].sum() for a, b in r]
                assert max(loads) < 1.2 * (w.sum() / world)


_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from lzfse_rust_b200 import shard_ranges
import oracle_binding as ob, testkit as tk
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# every rank sees the same stream list; it owns one contiguous range (host-side scatter, no data-path collective)
chunks = [tk.synth_text(1000 + i, 2000 + 97 * i) for i in range(24)]
enc = ob.Encoder()
frames = [enc.encode(c)[1] for c in chunks]
lo, hi = shard_ranges([len(c) + len(f) for c, f in zip(chunks, frames)], world)[rank]
# the CPU oracle stands in for the device here: this test is about the partition, not the kernels
mine = [ob.decode(f)[1] for f in frames[lo:hi]]
assert mine == chunks[lo:hi]
n = torch.tensor([hi - lo, sum(len(m) for m in mine)])
dist.all_reduce(n)  # bookkeeping only
assert int(n[0]) == len(chunks) and int(n[1]) == sum(len(c) for c in chunks)
t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)  # bench.py's max-over-ranks timing
assert float(t) == world
dist.destroy_process_group()
print("ok", rank, lo, hi)
"""


def test_two_rank_sharding_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_rust_sys_crate_matches_header():
    """rust/lzfse_b200_sys declares exactly the functions include/lzfse_b200.h declares, with the same number of
    parameters (the crates are source only: no rustc in this image), and its build.rs compiles every CUDA source the
    Python build does."""
    from lzfse_rust_b200 import build

    header = open(os.path.join(ROOT, "include", "lzfse_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    c_decl = {m.group(1): m.group(2) for m in re.finditer(r"\b(lzfse_b200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", header)}
    rs = open(os.path.join(ROOT, "rust", "lzfse_b200_sys", "src", "lib.rs")).read()
    rs_decl = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (lzfse_b200_[a-z0-9_]+)\s*\(([^)]*)\)", rs)}
    assert set(c_decl) == set(rs_decl), set(c_decl) ^ set(rs_decl)
    count = lambda a: 0 if a.strip() in ("", "void") else a.count(",") + 1
    for name in c_decl:
        assert count(c_decl[name]) == count(rs_decl[name]), name
    build_rs = open(os.path.join(ROOT, "rust", "lzfse_b200_sys", "build.rs")).read()
    for src in build.SOURCES:
        assert '"%s"' % src in build_rs, src
    assert "compute_100a" in build_rs
