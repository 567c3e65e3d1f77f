"""GPU tests of the C-ABI's calling conventions: the asynchronous batch calls (two batches in flight on two streams
from one thread), the NULL = stream 0 rule, the documented difference from the reference at dst[0]
(lz/writer.rs:156-157), and hostile headers inside a batch sized from untrusted probes."""
import struct

import numpy as np
import pytest

import oracle_binding as ob
import testkit as tk

pytestmark = pytest.mark.gpu


def _batch(torch, enc, chunks):
    frames, st = enc.encode_batch(chunks)
    assert not st.any()
    lens = np.array([len(f) for f in frames], np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    raw_len = np.array([len(c) for c in chunks], np.int64)
    raw_off = np.concatenate([[0], np.cumsum(raw_len)[:-1]])
    dev = torch.device("cuda", 0)
    t = lambda a: torch.tensor(a, dtype=torch.int64, device=dev)
    src = torch.tensor(np.frombuffer(b"".join(frames), np.uint8).copy(), device=dev)
    dst = torch.zeros(int(raw_len.sum()), dtype=torch.uint8, device=dev)
    return src, t(offs), t(lens), dst, t(raw_off), t(raw_len)


def test_async_two_batches_overlap():
    """Two decoders, two streams, one host thread: both calls return before their kernels have finished, the second
    batch starts on the device before the first has ended, and both results are right after sync()."""
    import torch

    import lzfse_rust_b200 as L

    enc, d1, d2 = L.LzfseEncoder(0), L.LzfseDecoder(0), L.LzfseDecoder(0)
    chunks = [tk.synth_text(0x700000 + i, 65536) for i in range(2048)]
    want = np.frombuffer(b"".join(chunks), np.uint8)
    a, b = _batch(torch, enc, chunks), _batch(torch, enc, chunks)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for d, x, s in ((d1, a, s1), (d2, b, s2)):  # warm-up: scratch allocation
        d.decode_batch_device(*x, stream=s)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(s1)
    l1, st1 = d1.decode_batch_device(*a, stream=s1, wait=False)
    ev[1].record(s1)
    ev[2].record(s2)
    l2, st2 = d2.decode_batch_device(*b, stream=s2, wait=False)
    ev[3].record(s2)
    running = not ev[3].query()  # the call came back while its kernels were still queued or running
    d1.sync(); d2.sync()
    torch.cuda.synchronize()
    assert running
    # overlap on the device: batch 2 began before batch 1 ended
    assert ev[0].elapsed_time(ev[2]) < ev[0].elapsed_time(ev[1])
    for x, ln, st in ((a, l1, st1), (b, l2, st2)):
        assert int((st != 0).sum()) == 0 and bool(torch.equal(ln, x[5]))
        assert np.array_equal(x[3].cpu().numpy(), want)
    for h in (enc, d1, d2):
        h.close()


def test_null_stream_orders_with_torch_default_stream():
    """cuda_stream = NULL is stream 0: inputs produced by work queued on torch's default stream just before the call
    (non-blocking H2D copies, fills) are seen, and a fill queued right after the call does not race with it."""
    import torch

    import lzfse_rust_b200 as L

    enc, dec = L.LzfseEncoder(0), L.LzfseDecoder(0)
    chunks = [tk.synth_text(0x710000 + i, 30000 + 17 * i) for i in range(256)]
    src, so, sl, dst, do, dl = _batch(torch, enc, chunks)
    want = np.frombuffer(b"".join(chunks), np.uint8)
    assert torch.cuda.current_stream().cuda_stream == 0
    for _ in range(3):
        pinned = src.cpu().pin_memory()
        src2 = torch.empty_like(src)
        src2.copy_(pinned, non_blocking=True)      # still in flight on stream 0 when the call is made
        dst.fill_(0xAA)
        ln, st = dec.decode_batch_device(src2, so, sl, dst, do, dl, wait=False)
        out = dst.clone()                           # ordered behind the decode on stream 0
        dst.fill_(0)
        dec.sync()
        assert int((st != 0).sum()) == 0 and np.array_equal(out.cpu().numpy(), want)
    enc.close(); dec.close()


def _vn_frame_first_match_before_dst0():
    # bvxn: sml_d opcode with L = 3, M = 3, D = 10 (3 bytes exist when the match starts), then end-of-stream
    payload = bytes([0xC0, 10]) + b"abc" + bytes([6, 0, 0, 0, 0, 0, 0, 0])
    return b"bvxn" + struct.pack("<II", 6, len(payload)) + payload + b"bvx$"


def test_match_before_dst0_is_bad_d_value():
    """Documented difference (include/lzfse_b200.h): the reference appends to a Vec and lets a match
    reach into bytes that were already there (lz/writer.rs:156-157); the C-ABI's destination starts empty, so a match
    that reaches before dst[0] is BadDValue -- what the reference itself reports for an empty Vec."""
    import lzfse_rust_b200 as L

    dec = L.LzfseDecoder(0)
    frame = _vn_frame_first_match_before_dst0()
    assert ob.decode(frame, cap=64)[0] == 3
    outs, st = dec.decode_batch([frame, frame], caps=[64, 64])
    assert list(st) == [3, 3]
    out = bytearray(b"0123456789abcdef")   # bytes already in the caller's buffer do not become match sources
    with pytest.raises(L.LzfseError) as ei:
        dec.decode_bytes(frame, out)
    assert ei.value.status == 3 and out == b"0123456789abcdef"
    dec.close()


def test_hostile_announcements_do_not_size_the_batch():
    """A 24-byte bvxn frame announcing 0xF0000000 bytes and a bvx2 header announcing 23 MB sit between good frames:
    decode_batch sizes its buffers from the frames' own headers, capped at what a frame of that size can produce, so the
    call neither allocates gigabytes nor fails as a whole -- the liars fail alone, the neighbours decode."""
    import lzfse_rust_b200 as L

    enc, dec = L.LzfseEncoder(0), L.LzfseDecoder(0)
    good = [tk.synth_text(0x720000 + i, 50000) for i in range(4)]
    frames, st = enc.encode_batch(good)
    assert not st.any()
    payload = bytes([0xE3]) + b"abc" + bytes([6, 0, 0, 0, 0, 0, 0, 0])
    liar_vn = b"bvxn" + struct.pack("<II", 0xF0000000, len(payload)) + payload + b"bvx$"
    f0 = bytearray(frames[0])
    f0[4:8] = struct.pack("<I", 23 * 1000 * 1000)   # n_raw_bytes of the first block
    batch = [frames[0], liar_vn, frames[1], bytes(f0), frames[2], frames[3]]
    outs, st = dec.decode_batch(batch)
    assert [int(x) for x in st[[0, 2, 4, 5]]] == [0, 0, 0, 0]
    assert [outs[0], outs[2], outs[4], outs[5]] == good
    assert st[1] != 0 and st[3] != 0 and outs[1] is None and outs[3] is None
    # with honest capacities the statuses are the reference's
    assert int(st[1]) == ob.decode(liar_vn, cap=len(liar_vn) * dec.MAX_RATIO + dec.MAX_SLACK)[0]
    enc.close(); dec.close()


def test_streaming_front_doors_and_cli(tmp_path):
    """LzfseRingEncoder / LzfseRingDecoder, the reader / writer adaptors and the lzfoo command line produce and accept
    the frames of encode_bytes (what the reference's test/src/ops.rs checks across its engines)."""
    import io
    import subprocess
    import sys

    import lzfse_rust_b200 as L

    data = tk.synth_text(0x730000, 300000)
    want = ob.Encoder().encode(data)[1]
    renc, rdec = L.LzfseRingEncoder(0), L.LzfseRingDecoder(0)
    dst = io.BytesIO()
    assert renc.encode(io.BytesIO(data), dst) == (len(data), len(want)) and dst.getvalue() == want
    out = io.BytesIO()
    assert rdec.decode(io.BytesIO(want), out) == (len(data), len(want)) and out.getvalue() == data
    w = renc.writer(io.BytesIO())
    for k in range(0, len(data), 7777):
        w.write(data[k:k + 7777])
    assert w.finalize().getvalue() == want
    r = rdec.reader(io.BytesIO(want))
    got = bytearray()
    while True:
        c = r.read(10001)
        if not c:
            break
        got += c
    assert bytes(got) == data
    bad = bytearray(want); bad[40] ^= 0x55
    with pytest.raises(L.LzfseError) as ei:
        rdec.decode(io.BytesIO(bytes(bad)), io.BytesIO())
    assert ei.value.status == ob.decode(bytes(bad), cap=len(data))[0]
    renc.close(); rdec.close()
    # lzfoo: file -> file, stdin -> stdout, -v statistics on stderr, exit code 1 on a bad frame
    src, enc_f, dec_f = tmp_path / "in.txt", tmp_path / "out.lzfse", tmp_path / "back.txt"
    src.write_bytes(data)
    run = lambda args, stdin=None: subprocess.run([sys.executable, "-m", "lzfse_rust_b200"] + args, input=stdin, capture_output=True, timeout=300)
    p = run(["-encode", "-i", str(src), "-o", str(enc_f), "-v"])
    assert p.returncode == 0 and enc_f.read_bytes() == want and b"Compression ratio" in p.stderr
    p = run(["-decode", "-i", str(enc_f), "-o", str(dec_f)])
    assert p.returncode == 0 and dec_f.read_bytes() == data
    p = run(["-decode"], stdin=want)
    assert p.returncode == 0 and p.stdout == data
    p = run(["-decode"], stdin=bytes(bad))
    assert p.returncode == 1 and p.stderr.startswith(b"Error: ")


def test_bounded_decode_matches_prefixes():
    """lzfse_b200_decode_prefix_batch_host (the reference's decode_n as a batch call): for every fixture and a few
    generated frames, limits of 0, 1, mid-block, a block boundary, the exact size and beyond must return that prefix of
    the oracle's output with the right `more`; damage behind the decoded blocks is not seen, damage inside them is."""
    import lzfse_rust_b200 as L

    enc, dec = L.LzfseEncoder(0), L.LzfseDecoder(0)
    cases = [(f, ob.decode(f)[1]) for name, f, _ in tk.golden_frames() if name != "special/null.vx2"]
    gen = [tk.synth_text(0x740000, 300000), bytes(500000), tk.rng_gen_vec(3, 100000), tk.synth_text(0x740001, 3000), b"0123456789", b""]
    frames, st = enc.encode_batch(gen)
    assert not st.any()
    cases += list(zip(frames, gen))
    batch, limits, want = [], [], []
    for f, raw in cases:
        _, _, nb = ob.probe(f)
        for lim in sorted({0, 1, len(raw) // 3, 40000, 65536, max(len(raw) - 1, 0), len(raw), len(raw) + 1000}):
            batch.append(f); limits.append(lim); want.append(raw[:lim])
    outs, st, more = dec.decode_prefix_batch(batch, limits)
    assert not st.any()
    for k, (o, w, lim) in enumerate(zip(outs, want, limits)):
        assert o == w, (k, lim, len(o), len(w))
        full = len(ob.decode(batch[k])[1])
        assert int(more[k]) == (1 if lim < full else 0), (k, lim, full, int(more[k]))
    # a long frame damaged in its last block: a limit inside the first blocks does not see it, the full size does
    big, raw = cases[-6]
    bad = big[:-300]   # cut inside its last block
    es, _ = ob.decode(bytes(bad), cap=len(raw))
    assert es != 0
    outs, st, more = dec.decode_prefix_batch([bytes(bad), bytes(bad), big], [50000, len(raw), len(raw)])
    assert list(st) == [0, es, 0] and outs[0] == raw[:50000] and outs[1] is None and outs[2] == raw and list(more) == [1, 0, 0]
    # untrusted announcements: a 24-byte frame claiming 0xE0000000 bytes costs nothing and fails like in the reference
    import struct
    payload = bytes([0xE3]) + b"abc" + bytes([6, 0, 0, 0, 0, 0, 0, 0])
    liar = b"bvxn" + struct.pack("<II", 0xE0000000, len(payload)) + payload + b"bvx$"
    outs, st, more = dec.decode_prefix_batch([liar, big], [10, 10])
    assert int(st[0]) == 33 and int(st[1]) == 0 and outs[1] == raw[:10]   # VnBadPayload, as the oracle says with room to spare
    assert ob.decode(liar, cap=1 << 20)[0] == 33
    enc.close(); dec.close()
