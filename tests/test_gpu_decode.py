"""GPU parity tests for the decode path, through the C-ABI (ctypes) -- the parity gate proper.

Mirrors the reference's integration tests: fixture SHA-256s (test/src/data.rs:84-97), generated
patterns round-tripped (test/src/pattern_*.rs, random_*.rs, len.rs), the mutate family
(test/src/mutate_0..7.rs: a broken frame yields an error, never a crash, and neighbours in the same
batch are untouched).  The CPU oracle (oracle/) is the checker: outputs must be byte-identical and
per-stream statuses equal."""
import numpy as np
import pytest

import oracle_binding as ob
import testkit as tk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["warp", "cta"])
def dec(request):
    """Every test runs once per expansion kernel (warp per stream / CTA per stream); left alone the library picks by
    batch size.  LZB_EXPAND is read when the decoder is created."""
    import os

    import lzfse_rust_b200 as L

    old = os.environ.get("LZB_EXPAND")
    os.environ["LZB_EXPAND"] = request.param
    try:
        d = L.LzfseDecoder(0)
    finally:
        if old is None:
            del os.environ["LZB_EXPAND"]
        else:
            os.environ["LZB_EXPAND"] = old
    yield d
    d.close()


@pytest.fixture(scope="module")
def oenc():
    return ob.Encoder()


def test_fixtures_one_by_one(dec):
    for name, frame, digest in tk.golden_frames():
        out = bytearray(b"keep")
        if name == "special/null.vx2":
            import lzfse_rust_b200 as L

            with pytest.raises(L.LzfseError) as ei:
                dec.decode_bytes(frame, out)
            assert ei.value.status == 30
            continue
        n = dec.decode_bytes(frame, out)
        assert out[:4] == b"keep" and n == len(out) - 4          # appended, like Vec<u8>
        assert tk.sha256(bytes(out[4:])) == digest, name
        assert ob.decode(frame) == (0, bytes(out[4:]))


def test_fixtures_one_batch(dec):
    gold = tk.golden_frames()
    frames = [g[1] for g in gold]
    outs, status = dec.decode_batch(frames)
    for (name, frame, digest), out, st in zip(gold, outs, status):
        ost, oout = ob.decode(frame)
        assert st == ost, name
        if ost == 0:
            assert out == oout and tk.sha256(out) == digest, name
    raw, nb, st = dec.probe_batch(np.frombuffer(b"".join(frames), np.uint8), np.cumsum([0] + [len(f) for f in frames[:-1]]), [len(f) for f in frames])
    for (name, frame, _), r, b, s in zip(gold, raw, nb, st):
        ost, oraw, onb = ob.probe(frame)
        assert s == ost, name
        if ost == 0:
            assert (r, b) == (oraw, onb), name


def test_synth_frames(dec, golden_dir):
    import os

    frames = [open(os.path.join(golden_dir, "data", "synth", n + ".lzfse"), "rb").read() for n in ("random", "word08", "repl01", "repl16", "reps08", "repsin")]
    outs, status = dec.decode_batch(frames)
    for f, o, s in zip(frames, outs, status):
        assert s == 0 and (0, o) == ob.decode(f)


def _patterns():
    yield "zeros", bytes(300000)                                        # M splits at 2359, D = 1 overlap
    yield "noise", tk.rng_gen_vec(1, 100001)                            # literal-only LMDs, 315-byte runs, 3 blocks
    yield "seq_masked", tk.seq_bytes(0, 200000, 0x03030000)             # test/src/huge.rs:16
    yield "text64k", tk.synth_text(0x5EED0000, 65536)
    yield "text1m", tk.synth_text(0x16000000, 1 << 20)                  # ~16 blocks, cross-block distances
    for p in (1, 2, 3, 5, 7, 8, 13, 16, 31, 64):
        yield "period%d" % p, (tk.rng_gen_vec(p, p) * (70000 // p + 1))[:70000]   # data/synth/repl*
    yield "lits_then_far_match", tk.rng_gen_vec(3, 250000) + tk.rng_gen_vec(3, 250000)
    for n in (0, 1, 2, 3, 4, 5, 19, 20, 21, 22, 63, 64, 300, 1000, 4095, 4096, 4097, 4098, 40000, 40001):
        yield "len%d" % n, tk.seq_bytes(n, n, 0x0F0F0F0F)               # test/src/len.rs


def test_patterns_roundtrip(dec, oenc):
    names, datas, frames = [], [], []
    for name, data in _patterns():
        st, comp = oenc.encode(data)
        assert st == 0
        names.append(name); datas.append(data); frames.append(comp)
    outs, status = dec.decode_batch(frames)
    for name, data, out, st in zip(names, datas, outs, status):
        assert st == 0, name
        assert out == data, name


def test_text_chunk_batch(dec, oenc):
    """The benchmark workload in miniature: independent 64 KiB text chunks."""
    chunks = [tk.synth_text(0x5EED0000 + i, 65536) for i in range(96)]
    frames = [oenc.encode(c)[1] for c in chunks]
    outs, status = dec.decode_batch(frames)
    assert (status == 0).all()
    assert outs == chunks
    assert dec.last_launches >= 6


def test_ragged_batch_device_api(dec, oenc):
    import torch

    rng = np.random.default_rng(5)
    chunks = [tk.synth_text(100 + i, int(rng.integers(0, 30000))) for i in range(70)] + [b"", b"x", bytes(5000)]
    frames = [oenc.encode(c)[1] for c in chunks]
    lens = np.array([len(f) for f in frames]); offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    caps = np.array([len(c) for c in chunks]); doff = np.concatenate([[0], np.cumsum(caps)[:-1]])
    dev = torch.device("cuda:0")
    src = torch.frombuffer(bytearray(b"".join(frames)), dtype=torch.uint8).to(dev)
    dst = torch.zeros(int(caps.sum()) + 1, dtype=torch.uint8, device=dev)
    t = lambda a: torch.tensor(a, dtype=torch.int64, device=dev)
    out_len, status = dec.decode_batch_device(src, t(offs), t(lens), dst, t(doff), t(caps))
    assert (status.cpu().numpy() == 0).all()
    assert (out_len.cpu().numpy() == caps).all()
    assert dst.cpu().numpy()[:-1].tobytes() == b"".join(chunks)


def _mutations(frame, seed, n):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        b = bytearray(frame)
        kind = rng.integers(0, 5)
        if kind == 0:
            i = rng.integers(0, len(b)); b[i] ^= 1 << rng.integers(0, 8)
        elif kind == 1:
            i = rng.integers(0, len(b)); b[i] = rng.integers(0, 256)
        elif kind == 2:
            i = rng.integers(0, max(1, len(b) - 4)); b[i:i + 4] = rng.integers(0, 256, 4, dtype=np.uint8).tobytes()
        elif kind == 3:
            b = b[: rng.integers(0, len(b))]
        else:
            b += rng.integers(0, 256, rng.integers(1, 9), dtype=np.uint8).tobytes()
        yield bytes(b)


@pytest.mark.parametrize("name", ["raw", "vx1", "vx2", "vxn", "compound"])
def test_mutated_frames_status_parity(dec, golden_dir, name):
    import os

    sub = "special" if name == "compound" else "mutate"
    frame = open(os.path.join(golden_dir, "data", sub, name + ".lzfse"), "rb").read()
    st0, good = ob.decode(frame)
    muts = list(_mutations(frame, 99, 300))
    # pristine copies interleaved: a failing neighbour must not disturb them
    frames = []
    for m in muts:
        frames += [m, frame]
    cap = 4 * len(good) + 64
    outs, status = dec.decode_batch(frames, caps=[cap] * len(frames))
    n_bad = 0
    for i, m in enumerate(muts):
        ost, oout = ob.decode(m, cap=cap)
        assert status[2 * i] == ost, (name, i, status[2 * i], ost)
        if ost == 0:
            assert outs[2 * i] == oout
        n_bad += ost != 0
        assert status[2 * i + 1] == 0 and outs[2 * i + 1] == good
    assert n_bad > 50


def test_frame_level_errors(dec):
    frames = [b"bvx$", b"bvx$\0", b"bvx", b"bvxq\0\0\0\0", b"", b"bvx-\x04\0\0\0abcdbvx$", b"bvx-\x05\0\0\0abcdbvx$"]
    outs, status = dec.decode_batch(frames, caps=[16] * len(frames))
    assert list(status) == [ob.decode(f, cap=16)[0] for f in frames] == [0, 6, 7, 1, 7, 0, 7]
    assert outs[0] == b"" and outs[5] == b"abcd"
    # capacity too small: the C-ABI's BufferOverflow
    st, comp = ob.encode(bytes(1000))
    outs, status = dec.decode_batch([comp], caps=[999])
    assert status[0] == 5
    assert dec.decode_batch([], caps=[])[0] == []


def test_large_host_batch_pipelined(dec):
    """>= 256 MiB crossing the bus: the host entry point slices the batch and overlaps copies with kernels.
    Size-independent property: decode(encode(x)) == x for every stream; one corrupted stream in the
    middle must fail alone."""
    import lzfse_rust_b200 as L
    from bench_support import workload as W

    enc = L.LzfseEncoder(0)
    pool, woff = W.word_pool(dec)
    n, cl = 3200, 65536
    raw = W.text_chunks(pool, woff, n, cl, seed0=0x0BAD5EED)
    bound = enc.encode_bound(cl)
    comp = np.empty(n * bound, np.uint8)
    c_len, st = enc.encode_batch_into(raw, np.arange(n) * cl, np.full(n, cl), comp, np.arange(n) * bound, np.full(n, bound))
    assert not st.any()
    offs = np.concatenate([[0], np.cumsum(c_len)[:-1]]).astype(np.uint64)
    packed = np.concatenate([comp[i * bound:i * bound + int(c_len[i])] for i in range(n)])
    bad = n // 2
    packed[int(offs[bad]) + 40] ^= 0xFF  # inside the weight table / payload of one frame
    out = np.zeros(n * cl, np.uint8)
    o_len, status = dec.decode_batch_into(packed, offs, c_len, out, np.arange(n) * cl, np.full(n, cl))
    frame = packed[int(offs[bad]):int(offs[bad]) + int(c_len[bad])].tobytes()
    assert status[bad] == ob.decode(frame, cap=cl)[0] != 0
    ok = np.ones(n, bool); ok[bad] = False
    assert not status[ok].any() and (o_len[ok] == cl).all()
    got, want = out.reshape(n, cl), raw.reshape(n, cl)
    assert np.array_equal(got[ok], want[ok])
    enc.close()
