"""Pins the CPU oracle's encoder: byte-exact KATs (src/encode/frontend_bytes.rs:455-531,
src/encode/mod.rs:50-54), the block-type policy tests (frontend_bytes.rs:534-546), DummyBackend LMD
expectations (frontend_bytes.rs:549-625), weight-normalisation invariants (fse/weights.rs:366-501),
and round-trips through the fixture-proven decoder (test/src/pattern_*.rs, random_*.rs, len.rs)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
import testkit as tk


def test_encoder_kats(golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "encoder_kat.json")))
    assert len(kat) == 7
    for name, v in kat.items():
        data = bytes(v["input_zero_len"]) if "input_zero_len" in v else v["input_ascii"].encode()
        st, out = ob.encode(data)
        assert st == 0 and out.hex() == v["frame_hex"], name


def test_block_type_policy():
    noise = tk.rng_gen_vec(0, 4096)
    assert ob.encode(noise)[1][:4] == b"bvx-"          # rand_vn_cutoff
    noise = tk.rng_gen_vec(0, 4097)
    assert ob.encode(noise)[1][:4] == b"bvx2"          # rand_vn_cutoff_add_1
    assert ob.encode(bytes(21))[1][:4] == b"bvxn"
    assert ob.encode(bytes(20))[1][:4] == b"bvx-"


def test_frontend_lmds_dummy_backend():
    enc = ob.Encoder()
    # `Dummy` (encode/dummy.rs:17-64) has MATCH_UNIT 3 like Vn; literal-only pushes are (L, 0, 0) here.
    assert enc.frontend_lmds(bytes(4), vn=True)[1] == [(4, 0, 0)]     # match_short_zero_4
    # the reference runs the next two over every n in 5..0x1000 / 12..0x1000 (`#[ignore = "expensive"]`,
    # encode/frontend_bytes.rs:580-625); so does this port, with both hash flavours
    for n in range(5, 0x1000):                                        # match_short_zero_n
        assert enc.frontend_lmds(bytes(n), vn=True)[1] == [(1, n - 1, 1)], n
        assert enc.frontend_lmds(bytes(n), vn=False)[1] == [(1, n - 1, 1)], n
    for n in range(12, 0x1000):                                       # sandwich_n_short
        b = bytearray(n); b[0:4] = b"\1\2\3\4"; b[n - 4:n] = b"\1\2\3\4"
        assert enc.frontend_lmds(bytes(b), vn=True)[1] == [(5, n - 9, 1), (0, 4, n - 4)], n
        if n >= 13:  # (the bvx2 flavour's shortest match is 4 bytes: at n = 12 the run of zeros is only 3 long)
            assert enc.frontend_lmds(bytes(b), vn=False)[1] == [(5, n - 9, 1), (0, 4, n - 4)], n


def test_normalize_invariants():
    rng = np.random.default_rng(7)
    for n_sym, n_states in ((20, 64), (64, 256), (256, 1024)):
        for _ in range(300):
            w = rng.integers(0, 50, n_sym).astype(np.uint16) * (rng.random(n_sym) < rng.random()).astype(np.uint16)
            if rng.random() < 0.3:
                w[rng.integers(0, n_sym)] = rng.integers(1, 10000)
            total = int(w.sum())
            if total == 0:
                continue
            arr = (C.c_uint16 * n_sym)(*w.tolist())
            ob.lib().orc_normalize_m1(arr, n_sym, total, n_states)
            out = np.array(arr[:])
            assert out.sum() == n_states                      # weights.rs:366-430
            assert ((out > 0) == (w > 0)).all()               # non-zero stays non-zero


def _patterns():
    yield "zeros", bytes(70000)
    yield "noise", tk.rng_gen_vec(1, 70001)
    yield "seq_masked", tk.seq_bytes(0, 100000, 0x03030000)   # test/src/huge.rs:16
    yield "text", tk.synth_text(0x5EED0000, 65536)
    yield "period7", (b"abcdefg" * 20000)[:131072]
    yield "long_literals_then_match", tk.rng_gen_vec(3, 50000) + tk.rng_gen_vec(3, 50000)
    for n in (0, 1, 2, 3, 4, 5, 19, 20, 21, 22, 63, 64, 4095, 4096, 4097, 4098, 40000, 40001):
        yield "len%d" % n, tk.seq_bytes(n, n, 0x0F0F0F0F)     # test/src/len.rs


@pytest.mark.parametrize("name,data", list(_patterns()), ids=[p[0] for p in _patterns()])
def test_roundtrip(name, data):
    enc = ob.Encoder()
    st, comp = enc.encode(data)
    assert st == 0
    assert len(comp) <= ob.lib().orc_encode_bound(len(data))
    st, back = ob.decode(comp)
    assert st == 0 and back == data
    # encoder object is reusable (bench.rs:195-209): same bytes the second time
    assert enc.encode(data) == (0, comp)


def test_fixture_decode_encode_decode():
    # test/src/data.rs: decode -> encode -> decode identity; ratio sanity vs Apple's C encoder
    enc = ob.Encoder()
    for name, frame, digest in tk.golden_frames("snappy"):
        st, raw = ob.decode(frame)
        st2, comp = enc.encode(raw)
        st3, back = ob.decode(comp)
        assert (st, st2, st3) == (0, 0, 0) and back == raw
        assert len(comp) <= 1.01 * len(frame), name


def test_fse_backend_lmd_fuzz():
    """fse/test.rs:257-319: random LMD lists through the FSE backend decode back to the same bytes."""
    rng = np.random.default_rng(11)
    enc = ob.Encoder()
    for it in range(20):
        lmds, lits, size = [], bytearray(), 0
        for _ in range(int(rng.integers(1, 3000))):
            l = int(rng.integers(0, 400)) if rng.random() < 0.1 else int(rng.integers(0, 8))
            m = int(rng.integers(0, 3000)) if rng.random() < 0.05 else int(rng.integers(0, 40))
            if size + l == 0:
                l = 1
            d = int(rng.integers(1, min(size + l, 262139) + 1))
            lits += rng.integers(0, 256, l, dtype=np.uint8).tobytes()
            if l == 0 and m == 0:
                m = 4
            lmds.append((l, m, d if m else 0))
            size += l + m
        st, blocks = enc.fse_encode_lmds(bytes(lits), lmds)
        assert st == 0
        st, out, tr = ob.decode_trace(blocks + b"bvx$", size)
        assert st == 0 and len(out) == size
        # replay the LMD list directly
        exp = bytearray(); p = 0
        for l, m, d in lmds:
            exp += lits[p:p + l]; p += l
            for _ in range(m):
                exp.append(exp[-d])
        assert bytes(exp) == out


def test_batch_threads_match_single():
    chunks = [tk.synth_text(0x5EED0000 + i, 8192 + 37 * i) for i in range(24)]
    src = b"".join(chunks)
    n = len(chunks)
    u64 = C.c_uint64 * n
    off = np.cumsum([0] + [len(c) for c in chunks[:-1]]).tolist()
    bound = [ob.lib().orc_encode_bound(len(c)) for c in chunks]
    doff = np.cumsum([0] + bound[:-1]).tolist()
    dst = C.create_string_buffer(sum(bound))
    out_len, status = u64(), (C.c_int32 * n)()
    sbuf = C.create_string_buffer(src, len(src))
    ob.lib().orc_encode_batch(sbuf, u64(*off), u64(*[len(c) for c in chunks]), dst, u64(*doff), u64(*bound), out_len, status, n, 4)
    enc = ob.Encoder()
    for i, c in enumerate(chunks):
        assert status[i] == 0
        assert dst.raw[doff[i]:doff[i] + out_len[i]] == enc.encode(c)[1]
    # decode them back, 3 threads
    back = C.create_string_buffer(len(src))
    out2, st2 = u64(), (C.c_int32 * n)()
    ob.lib().orc_decode_batch(dst, u64(*doff), out_len, back, u64(*off), u64(*[len(c) for c in chunks]), out2, st2, n, 3)
    assert list(st2) == [0] * n and back.raw == src
