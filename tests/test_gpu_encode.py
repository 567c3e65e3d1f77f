"""GPU parity tests for the encode path, through the C-ABI.

The GPU encoder is designed to reproduce lzfse_rust's encode_bytes bit for bit, so the gate is frame
equality with the CPU oracle (itself pinned by the reference's byte-exact KATs).  That subsumes the
north-star gates: FSE stage bit-exact for a given LMD stream and frequency table, every frame decodes
under the (fixture-proven) oracle decoder, compression ratio identical at equal chunking."""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
import testkit as tk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def enc():
    import lzfse_rust_b200 as L

    e = L.LzfseEncoder(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def dec():
    import lzfse_rust_b200 as L

    d = L.LzfseDecoder(0)
    yield d
    d.close()


def test_encoder_kats(enc, golden_dir):
    """src/encode/frontend_bytes.rs:455-531 and src/encode/mod.rs:50-54, byte for byte."""
    kat = json.load(open(os.path.join(golden_dir, "encoder_kat.json")))
    for name, v in kat.items():
        data = bytes(v["input_zero_len"]) if "input_zero_len" in v else v["input_ascii"].encode()
        out = bytearray(b"keep")
        n = enc.encode_bytes(data, out)
        assert out[:4] == b"keep" and n == len(out) - 4          # appended, like Vec<u8>
        assert bytes(out[4:]).hex() == v["frame_hex"], name


def _patterns():
    yield "zeros", bytes(300000)
    yield "noise", tk.rng_gen_vec(1, 100001)
    yield "seq_masked", tk.seq_bytes(0, 200000, 0x03030000)
    yield "text64k", tk.synth_text(0x5EED0000, 65536)
    yield "text1m", tk.synth_text(0x16000000, 1 << 20)
    for p in (1, 2, 3, 5, 7, 8, 13, 16, 31, 64):
        yield "period%d" % p, (tk.rng_gen_vec(p, p) * (70000 // p + 1))[:70000]
    yield "lits_then_far_match", tk.rng_gen_vec(3, 250000) + tk.rng_gen_vec(3, 250000)
    yield "noise_vn", tk.rng_gen_vec(9, 4096)                 # LZVN not smaller => raw fallback
    yield "noise_fse", tk.rng_gen_vec(9, 4097)                # never raw above the cutoff
    for n in (0, 1, 2, 3, 4, 5, 19, 20, 21, 22, 63, 64, 300, 1000, 4095, 4096, 4097, 4098, 40000, 40001):
        yield "len%d" % n, tk.seq_bytes(n, n, 0x0F0F0F0F)
    for n in (21, 100, 1000, 4096):
        yield "zeros%d" % n, bytes(n)
        yield "text%d" % n, tk.synth_text(n, n)


def test_frames_equal_oracle(enc, dec):
    names, datas = zip(*_patterns())
    frames, status = enc.encode_batch(list(datas))
    oenc = ob.Encoder()
    for name, data, frame, st in zip(names, datas, frames, status):
        ost, oframe = oenc.encode(data)
        assert st == 0 and ost == 0, name
        assert frame == oframe, (name, len(frame), len(oframe))
        assert ob.decode(frame) == (0, data), name
    outs, dstat = dec.decode_batch(list(frames))
    assert (dstat == 0).all() and list(outs) == list(datas)


def test_all_front_ends_agree(enc):
    """The encoder has three front ends for bvx2 streams -- per segment (default), per stream (`LZB_ENC_SEG=0`: k_enc_replay for
    streams <= 64 KiB) and one warp per stream (`LZB_ENC_LONG=0`: k_enc_parse for longer ones).  All must emit the same frames."""
    import lzfse_rust_b200 as L

    datas = [d for n, d in _patterns() if n in ("zeros", "seq_masked", "text64k", "period3", "period64", "lits_then_far_match", "len40001", "noise_fse")]
    datas += [tk.synth_text(0x77, 200000), tk.synth_text(0x78, 65536)]
    ref, st = enc.encode_batch(datas)
    assert not st.any()
    for var in ("LZB_ENC_SEG", "LZB_ENC_LONG"):
        os.environ[var] = "0"
        try:
            e2 = L.LzfseEncoder(0)
            frames, st = e2.encode_batch(datas)
            e2.close()
        finally:
            del os.environ[var]
        assert not st.any() and list(frames) == list(ref), var


def test_text_chunk_batch(enc):
    chunks = [tk.synth_text(0x5EED0000 + i, 65536) for i in range(64)]
    frames, status = enc.encode_batch(chunks)
    assert (status == 0).all()
    oenc = ob.Encoder()
    for c, f in zip(chunks, frames):
        assert f == oenc.encode(c)[1]
    assert enc.last_launches >= 4


def test_fixture_payloads(enc, dec):
    """decode -> encode -> decode identity on the reference's corpus (test/src/data.rs)."""
    gold = [g for g in tk.golden_frames("snappy")]
    raws, _ = dec.decode_batch([g[1] for g in gold])
    frames, status = enc.encode_batch(raws)
    oenc = ob.Encoder()
    for (name, _, digest), raw, f, st in zip(gold, raws, frames, status):
        assert st == 0 and f == oenc.encode(raw)[1], name
    back, st2 = dec.decode_batch(frames)
    assert (st2 == 0).all() and back == raws


def test_ragged_device_api(enc):
    import torch

    rng = np.random.default_rng(17)
    chunks = [tk.synth_text(500 + i, int(rng.integers(0, 20000))) for i in range(90)] + [b"", b"x", bytes(5000), tk.rng_gen_vec(4, 777)]
    lens = np.array([len(c) for c in chunks]); offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    caps = np.array([enc.encode_bound(int(l)) for l in lens]); doff = np.concatenate([[0], np.cumsum(caps)[:-1]])
    dev = torch.device("cuda:0")
    src = torch.frombuffer(bytearray(b"".join(chunks) + b"\0"), dtype=torch.uint8).to(dev)
    dst = torch.zeros(int(caps.sum()) + 1, dtype=torch.uint8, device=dev)
    t = lambda a: torch.tensor(a, dtype=torch.int64, device=dev)
    out_len, status = enc.encode_batch_device(src, t(offs), t(lens), dst, t(doff), t(caps))
    out_len, status, host = out_len.cpu().numpy(), status.cpu().numpy(), dst.cpu().numpy()
    oenc = ob.Encoder()
    for i, c in enumerate(chunks):
        assert status[i] == 0
        assert host[doff[i]:doff[i] + out_len[i]].tobytes() == oenc.encode(c)[1], i


def test_capacity_too_small(enc):
    data = tk.synth_text(1, 30000)
    n = len(ob.encode(data)[1])
    src = np.frombuffer(data, np.uint8)
    dst = np.zeros(n, np.uint8)
    out_len, status = enc.encode_batch_into(src, [0], [len(data)], dst, [0], [n - 1])
    assert status[0] == 5 and out_len[0] == 0
    out_len, status = enc.encode_batch_into(src, [0], [len(data)], dst, [0], [n])
    assert status[0] == 0 and out_len[0] == n and dst.tobytes() == ob.encode(data)[1]
    assert enc.encode_batch([])[0] == []


def _patchwork(seed, n):
    """Random patchwork in the spirit of test/src/patchwork_*.rs: literal runs of several kinds interleaved with copies of
    earlier output at random distances (short, medium, beyond 64 KiB, beyond the 262 139-byte window) and random lengths
    (below GOOD_MATCH_LEN, above it, above MAX_M_VALUE), plus byte runs."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    while len(out) < n:
        kind = rng.integers(0, 6)
        if kind == 0:
            out += rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8).tobytes()
        elif kind == 1:
            out += tk.synth_text(int(rng.integers(0, 1 << 20)), int(rng.integers(1, 600)))
        elif kind == 2:
            out += bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 300))
        elif len(out) > 8:
            d = int(rng.choice([rng.integers(1, 17), rng.integers(1, 1024), rng.integers(1, 70000), rng.integers(1, 300000)]))
            d = min(d, len(out))
            m = int(rng.choice([rng.integers(3, 12), rng.integers(12, 60), rng.integers(60, 5000)]))
            start = len(out) - d
            for k in range(m):  # LZ77 semantics, overlaps included
                out.append(out[start + k])
    return bytes(out[:n])


def test_patchwork_frames_equal(enc, dec):
    """120 random patchworks of 1 B .. 400 KB (LZVN, single- and multi-block bvx2, far and out-of-window copies)."""
    rng = np.random.default_rng(4242)
    sizes = [int(x) for x in np.concatenate([rng.integers(1, 4097, 40), rng.integers(4097, 70000, 50), rng.integers(70000, 400000, 30)])]
    datas = [_patchwork(1000 + i, n) for i, n in enumerate(sizes)]
    frames, status = enc.encode_batch(datas)
    assert (status == 0).all()
    oenc = ob.Encoder()
    for i, (d, f) in enumerate(zip(datas, frames)):
        assert f == oenc.encode(d)[1], (i, len(d))
    outs, dst = dec.decode_batch(list(frames))
    assert (dst == 0).all() and list(outs) == datas


def _fast_path_cases():
    """Inputs for the shared-memory parse (k_enc_find / k_enc_replay: bvx2 streams of 4097..65536 bytes): the length
    limits on both sides, runs longer than the per-lane extension cap (1024) and than one pack's M (2359), several
    candidates beyond the cap at once (periodic data: exact lengths decide), literal runs beyond 315 and beyond a
    block's 40 000, and a duplicate that sends the backward extension past what a word stores (15)."""
    rnd = tk.rng_gen_vec(7, 70000)
    txt = tk.synth_text(0xFA57, 65536)
    yield "len4097", txt[:4097]
    yield "len4100", rnd[:4100]
    yield "len65535", txt[:65535]
    yield "len65536_text", txt
    yield "len65536_noise", rnd[:65536]
    yield "len65537_slow_path", (txt + b"x")[:65537]
    yield "zeros65536", bytes(65536)
    for p in (1, 2, 3, 4, 5, 7, 16, 64, 333, 1000, 5000):
        yield "period%d" % p, (rnd[:p] * (65536 // p + 1))[:65536]
    yield "long_dup", txt[:20000] + txt[:20000] + txt[:20000]
    yield "dup_after_noise", rnd[:30000] + txt[1000:9000] + rnd[30000:40000] + txt[1000:9000]
    yield "two_periods", (rnd[:3] * 4000)[:9000] + (rnd[10:17] * 4000)[:30000] + (rnd[:3] * 4000)[:9000]
    yield "noise_then_zeros", rnd[:41000] + bytes(24000)
    yield "back_ext", rnd[:500] + txt[:3000] + rnd[500:520] + txt[:3000] + rnd[600:5000]


def test_fast_parse_cases(enc, dec):
    names, datas = zip(*_fast_path_cases())
    frames, status = enc.encode_batch(list(datas))
    oenc = ob.Encoder()
    for name, data, frame, st in zip(names, datas, frames, status):
        ost, oframe = oenc.encode(data)
        assert st == 0 and ost == 0, name
        assert frame == oframe, (name, len(frame), len(oframe))
    outs, dstat = dec.decode_batch(list(frames))
    assert (dstat == 0).all() and list(outs) == list(datas)


def test_fast_and_general_parse_agree(dec):
    """LZB_ENC_FAST=0 sends every stream through the general kernel (k_enc_parse): same frames."""
    import lzfse_rust_b200 as L

    datas = [d for _, d in _fast_path_cases()] + [tk.synth_text(0x5EED0000 + i, 65536) for i in range(40)]
    old = os.environ.get("LZB_ENC_FAST")
    try:
        os.environ["LZB_ENC_FAST"] = "0"
        e0 = L.LzfseEncoder(0)
        f0, s0 = e0.encode_batch(datas)
        e0.close()
        os.environ["LZB_ENC_FAST"] = "1"
        e1 = L.LzfseEncoder(0)
        f1, s1 = e1.encode_batch(datas)
        e1.close()
    finally:
        if old is None:
            os.environ.pop("LZB_ENC_FAST", None)
        else:
            os.environ["LZB_ENC_FAST"] = old
    assert (s0 == 0).all() and (s1 == 0).all() and list(f0) == list(f1)


def test_reference_frontend_families_all_lengths(enc, dec):
    """The reference's exhaustive front-end tests (`match_short_zero_n`, `sandwich_n_short`, encode/frontend_bytes.rs:580-625:
    every n up to 0x1000, ignored there as expensive) plus the same two families on the bvx2 side of the cutoff (4097..4300):
    one GPU batch, every frame byte-identical to the port's -- whose LMDs for these inputs are pinned to the reference's
    expected values in tests/test_oracle_encode.py -- and decoded back."""
    chunks = [bytes(n) for n in range(5, 0x1000)]
    for n in list(range(12, 0x1000)) + list(range(4097, 4300)):
        b = bytearray(n); b[0:4] = b"\1\2\3\4"; b[n - 4:n] = b"\1\2\3\4"
        chunks.append(bytes(b))
    chunks += [bytes(n) for n in range(4097, 4300)]
    frames, st = enc.encode_batch(chunks)
    assert not st.any()
    oenc = ob.Encoder()
    for c, f in zip(chunks, frames):
        assert f == oenc.encode(c)[1], len(c)
    outs, dst = dec.decode_batch(frames)
    assert not dst.any() and outs == chunks
