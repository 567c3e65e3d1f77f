"""Parity on the other BASELINE.json workloads at reduced scale: large multi-block streams with
cross-block match distances (configs[3]) and the mixed corpus (configs[4]: incompressible, LZVN-sized
and raw-sized chunks, highly repetitive data, text).  Both directions, checked against the oracle."""
import numpy as np
import pytest

import oracle_binding as ob
import testkit as tk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    import lzfse_rust_b200 as L

    e, d = L.LzfseEncoder(0), L.LzfseDecoder(0)
    yield e, d
    e.close(); d.close()


def test_large_streams(codec):
    """One 16 MiB stream (~250 bvx2 blocks, matches reaching up to 262139 bytes back across blocks) plus 4 MiB ones."""
    from bench_support import workload as W

    enc, dec = codec
    pool, woff = W.word_pool(dec)
    big = W.text_chunks(pool, woff, 1, 16 << 20, seed0=0x16000000).tobytes()
    mids = [W.text_chunks(pool, woff, 1, 4 << 20, seed0=0x16000001 + i).tobytes() for i in range(2)]
    rep = (tk.rng_gen_vec(5, 300000) * 14)[: 4 << 20]  # period 300000 > max distance: long far matches fail, near ones hit
    streams = [big] + mids + [rep]
    frames, st = enc.encode_batch(streams)
    assert not st.any()
    oenc = ob.Encoder()
    for s, f in zip(streams, frames):
        assert f == oenc.encode(s)[1]                      # bit-exact with the reference encoder
    raw, nb, pst = dec.probe_batch(np.frombuffer(b"".join(frames), np.uint8), np.cumsum([0] + [len(f) for f in frames[:-1]]), [len(f) for f in frames])
    assert not pst.any() and list(raw) == [len(s) for s in streams] and nb[0] > 200
    outs, dst = dec.decode_batch(frames)
    assert not dst.any() and outs == streams


def _assorted_long_inputs():
    rng = np.random.default_rng(1)
    t = tk.synth_text(9, 1 << 20)
    patch = bytearray()
    while len(patch) < 700000:
        k = int(rng.integers(0, 5)); n = int(rng.integers(1, 3000))
        if k == 0: patch += t[int(rng.integers(0, len(t) - n * 10)):][:n * 10]
        elif k == 1: patch += bytes([int(rng.integers(0, 256))]) * n               # byte runs (M > 2359 splits)
        elif k == 2: patch += rng.integers(0, 256, n * 10, dtype=np.uint8).tobytes()  # noise (L > 315 splits)
        elif k == 3: patch += (rng.integers(0, 4, n * 10, dtype=np.uint8) * 16).tobytes()
        else:
            d = int(rng.integers(1, min(len(patch), 300000) + 1)) if patch else 1
            for _ in range(n): patch.append(patch[-d] if len(patch) >= d else 0)
    sparse = bytearray(rng.integers(0, 256, 1 << 20, dtype=np.uint8).tobytes())
    for _ in range(2000):  # short repeats in noise: matches stay pending across segment borders
        a = int(rng.integers(100000, len(sparse) - 100)); d = int(rng.integers(8, 90000)); n = int(rng.integers(4, 30))
        sparse[a:a + n] = sparse[a - d:a - d + n]
    return {"text200k": t[:200000], "text1m": t, "text65537": t[:65537], "text70001": t[5:70006], "zeros300k": bytes(300000),
            "noise300k": tk.rng_gen_vec(3, 300000), "period1000": (tk.rng_gen_vec(5, 1000) * 400)[:333333], "period3": (b"abc" * 50000)[:140001],
            "patch": bytes(patch), "sparse": bytes(sparse), "far": (tk.rng_gen_vec(5, 300000) * 4)[:1 << 20], "lits": tk.seq_bytes(1, 500000, 0x0F0F0F0F)}


def test_long_encoder_assorted(codec):
    """Streams longer than 64 KiB take the segmented front end (encode_long.cuh: hash chain per 64 Ki positions, find_match for
    every position, speculative replay per segment, stitch, packs).  Frames must equal the oracle's byte for byte on inputs that
    stress each piece: lengths around the segment sizes, one match covering the whole stream, no matches at all, long periods
    (several candidates longer than the per-position word can say), L / M splits and block closes in odd places, matches pending
    across many segments, distances beyond the window."""
    enc, dec = codec
    data = _assorted_long_inputs()
    names = list(data)
    frames, st = enc.encode_batch([data[k] for k in names])
    assert not st.any()
    oenc = ob.Encoder()
    for k, f in zip(names, frames):
        assert f == oenc.encode(data[k])[1], k
    outs, dst = dec.decode_batch(frames)
    assert not dst.any() and outs == [data[k] for k in names]
    # the same streams next to short ones (all three parse paths in one batch)
    mixed = [data["text200k"], tk.synth_text(1, 3000), data["zeros300k"], tk.synth_text(2, 65536), b"", data["text70001"], tk.synth_text(3, 20000)]
    frames, st = enc.encode_batch(mixed)
    assert not st.any()
    for s, f in zip(mixed, frames):
        assert f == oenc.encode(s)[1]


def test_mixed_corpus(codec):
    enc, dec = codec
    rng = np.random.default_rng(2024)
    chunks = []
    for i in range(48):                                                     # (i) incompressible 64 KiB
        chunks.append(tk.rng_gen_vec(i, 65536))
    x = 12345
    for i in range(400):                                                    # (ii) LZVN-sized + a sprinkle of raw-sized
        x = (x * 1103515245 + 12345) & 0xFFFFFFFF
        n = 21 + (x % 4076) if i % 10 else x % 21
        chunks.append(tk.synth_text(7000 + i, n) if i % 3 else tk.seq_bytes(i, n, 0x07070707))
    for p in list(range(1, 17)) + [32, 64]:                                  # (iii) period-p repeats + the huge-test stream
        chunks.append((tk.rng_gen_vec(p, p) * (65536 // p + 1))[:65536])
    chunks.append(tk.seq_bytes(0, 65536, 0x03030000))
    for i in range(48):                                                     # (iv) text
        chunks.append(tk.synth_text(0x5EED0000 + i, 65536))
    order = rng.permutation(len(chunks))
    chunks = [chunks[i] for i in order]
    frames, st = enc.encode_batch(chunks)
    assert not st.any()
    oenc = ob.Encoder()
    kinds = set()
    for c, f in zip(chunks, frames):
        assert f == oenc.encode(c)[1]
        kinds.add(f[:4])
    assert kinds >= {b"bvx-", b"bvxn", b"bvx2"}
    outs, dst = dec.decode_batch(frames)
    assert not dst.any() and outs == chunks


def test_many_small_streams(codec):
    """5 000 inputs of 0..4200 bytes in one batch: the LZVN window kernel, raw blocks and (above 4096 bytes) small bvx2
    frames side by side, with enough streams for the warp-per-stream expansion kernel.  Frames must equal the oracle
    encoder's (sampled) and decode back on both sides."""
    enc, dec = codec
    x, chunks = 99, []
    for i in range(5000):
        x = (x * 1103515245 + 12345) & 0xFFFFFFFF
        n = (x >> 8) % 4201
        kind = i % 4
        if kind == 0:
            c = tk.synth_text(90000 + i, n)
        elif kind == 1:
            c = tk.seq_bytes(i, n, 0x0F0F0F0F)
        elif kind == 2:
            c = (tk.rng_gen_vec(i, 1 + i % 37) * (n // (1 + i % 37) + 1))[:n]
        else:
            c = tk.rng_gen_vec(i, n)
        chunks.append(c)
    chunks[0], chunks[1] = b"", b"x" * 4097   # the empty frame and the smallest bvx2 input
    frames, st = enc.encode_batch(chunks)
    assert not st.any()
    oenc = ob.Encoder()
    for i in range(0, len(chunks), 41):
        assert frames[i] == oenc.encode(chunks[i])[1], i
        assert ob.decode(frames[i]) == (0, chunks[i]), i
    outs, dst = dec.decode_batch(frames)
    assert not dst.any() and outs == chunks
    assert {f[:4] for f in frames} >= {b"bvx-", b"bvxn", b"bvx2"}


def test_long_stream_compound_and_hostile(codec):
    """The two-pass expansion of long streams (expand_long.cu): bvx2 blocks with a raw and an LZVN block between them
    (blocks of independent frames concatenated: their matches still reach the same bytes), next to short streams that
    take the in-order kernels; then the same long stream with one bit flipped in many places and cut short -- statuses
    and bytes must equal the oracle's."""
    from bench_support import workload as W

    enc, dec = codec
    pool, woff = W.word_pool(dec)
    parts = [W.text_chunks(pool, woff, 1, 640 << 10, seed0=0x17000000).tobytes(), b"0123456789", tk.synth_text(77, 2000),
             W.text_chunks(pool, woff, 1, 512 << 10, seed0=0x17000001).tobytes(), tk.rng_gen_vec(9, 70000), bytes(400000)]
    pf, st = enc.encode_batch(parts)
    assert not st.any() and all(f.endswith(b"bvx$") for f in pf)
    compound = b"".join(f[:-4] for f in pf) + b"bvx$"
    want = b"".join(parts)
    assert ob.decode(compound) == (0, want)
    shorts = [tk.synth_text(500 + i, 65536) for i in range(3)]
    sf, st = enc.encode_batch(shorts)
    assert not st.any()
    outs, dst = dec.decode_batch([compound, sf[0], pf[0], sf[1], pf[3], sf[2]])
    assert not dst.any() and outs == [want, shorts[0], parts[0], shorts[1], parts[3], shorts[2]]
    # hostile: bit flips spread over the long frame (headers, weights, payloads of different blocks), truncations
    frames = []
    for k in range(96):
        b = bytearray(compound)
        pos = (k * 7919 * 131) % len(b)
        b[pos] ^= 1 << (k % 8)
        frames.append(bytes(b))
    for k in range(16):
        frames.append(compound[: (len(compound) * (k + 1)) // 17])
    frames.append(compound)
    caps = [len(want)] * len(frames)
    outs, dst = dec.decode_batch(frames, caps=caps)
    for f, o, s in zip(frames, outs, dst):
        es, eo = ob.decode(f, cap=len(want))
        assert int(s) == es, (int(s), es)
        if es == 0:
            assert o == eo
