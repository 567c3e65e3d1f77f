"""Hostile-input matrix for the batched GPU decoder: the reference's mutate family
(test/src/mutate_0..7.rs) run as whole batches through the C-ABI.

The reference only asks its decoder not to hang, crash or panic on these inputs.  A batched GPU decoder has more to
lose -- one bad stream next to thousands of good ones -- so the bar here is higher:

  * every mutated frame gets exactly the status the CPU oracle (= the reference's sequential decoder) gives it, and
    the same bytes when that status is Ok;
  * pristine frames interleaved in the same batch decode to the fixture's SHA-256, untouched by failing neighbours;
  * canary bytes between the output regions survive (no stream writes outside its region).

mutate_1 (every index x every XOR byte) is run in full on the LZVN and raw fixtures; on the two FSE fixtures all 255
values hit the first 128 and last 32 bytes (headers, the start of the weight table, the end of the LMD payload where
the bit reader starts) and a stride-7 subset hits the rest; the bit-level mutate_0 covers every position."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_binding as ob
import testkit as tk

pytestmark = pytest.mark.gpu

FIXTURES = ("raw", "vxn", "vx1", "vx2")
CANARY = 0xA5
GAP = 32  # canary bytes after every output region


@pytest.fixture(scope="module", params=["warp", "cta"])
def dec(request):
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


def _fixture(golden_dir, name):
    frame = open(os.path.join(golden_dir, "data", "mutate", name + ".lzfse"), "rb").read()
    digest = open(os.path.join(golden_dir, "data", "mutate", name + ".hash"), "rb").read()
    return np.frombuffer(frame, np.uint8), digest


def _oracle_batch(src, off, lens, cap):
    n = len(lens)
    u64 = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
    P64, P32 = C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    p = lambda a, t: a.ctypes.data_as(t)
    out = np.zeros(n * cap + 8, np.uint8)
    d_off, d_cap = u64(np.arange(n) * cap), u64(np.full(n, cap))
    o_len, st = np.zeros(n, np.uint64), np.zeros(n, np.int32)
    s_off, s_len = u64(off), u64(lens)
    ob.lib().orc_decode_batch(src.ctypes.data, p(s_off, P64), p(s_len, P64), out.ctypes.data, p(d_off, P64), p(d_cap, P64), p(o_len, P64), p(st, P32), n,
                              os.cpu_count() or 1)
    return out, o_len, st


def _check(dec, frames, lens, pristine, digest, must_fail=False):
    """frames: uint8 [n, width] (row i holds lens[i] valid bytes).  Every 64th row is replaced by the pristine frame."""
    n, width = frames.shape
    lens = np.asarray(lens, np.int64).copy()
    sentinel = np.arange(0, n, 64)
    frames[sentinel, : len(pristine)] = pristine
    lens[sentinel] = len(pristine)
    st0, good = ob.decode(pristine.tobytes())
    assert st0 == 0 and tk.sha256(good) == digest
    cap = 2 * len(good) + 64
    src = np.ascontiguousarray(frames).reshape(-1)
    off = np.arange(n, dtype=np.int64) * width
    # GPU: output regions separated by canaries
    stride = cap + GAP
    out = np.full(n * stride, CANARY, np.uint8)
    o_len, status = dec.decode_batch_into(src, off, lens, out, np.arange(n, dtype=np.int64) * stride, np.full(n, cap, np.int64))
    want_out, want_len, want_st = _oracle_batch(src, off, lens, cap)
    bad = np.flatnonzero(status != want_st)[:8]
    assert len(bad) == 0, "status mismatch (row, gpu, oracle): %s" % [(int(i), int(status[i]), int(want_st[i])) for i in bad]
    ok = want_st == 0
    assert np.array_equal(o_len[ok], want_len[ok])
    got = out.reshape(n, stride)
    assert (got[:, cap:] == CANARY).all(), "a stream wrote past its output region"
    want = want_out[: n * cap].reshape(n, cap)
    for i in np.flatnonzero(ok):
        assert np.array_equal(got[i, : int(want_len[i])], want[i, : int(want_len[i])]), i
    assert ok[sentinel].all()
    for i in sentinel[:: max(1, len(sentinel) // 16)]:
        assert tk.sha256(got[i, : len(good)].tobytes()) == digest
    if must_fail:
        mut = np.ones(n, bool); mut[sentinel] = False
        assert (want_st[mut] != 0).all()
    return int((~ok).sum())


@pytest.mark.parametrize("name", FIXTURES)
def test_mutate_0_every_bit(dec, golden_dir, name):
    """test/src/mutate_0.rs: sequential bit mutation, every bit of the frame."""
    f, digest = _fixture(golden_dir, name)
    n = len(f) * 8
    frames = np.tile(f, (n, 1))
    idx = np.arange(n)
    frames[idx, idx // 8] ^= (1 << (idx % 8)).astype(np.uint8)
    assert _check(dec, frames, np.full(n, len(f)), f, digest) > 0


@pytest.mark.parametrize("name", FIXTURES)
def test_mutate_1_every_byte_value(dec, golden_dir, name):
    """test/src/mutate_1.rs: sequential byte mutation (XOR with 1..255)."""
    f, digest = _fixture(golden_dir, name)
    if len(f) <= 1024:
        pos = np.arange(len(f))
    else:
        pos = np.unique(np.concatenate([np.arange(128), np.arange(len(f) - 32, len(f)), np.arange(128, len(f) - 32, 7)]))
    vals = np.arange(1, 256, dtype=np.uint8)
    n = len(pos) * len(vals)
    frames = np.tile(f, (n, 1))
    rows = np.arange(n)
    frames[rows, np.repeat(pos, len(vals))] ^= np.tile(vals, len(pos))
    assert _check(dec, frames, np.full(n, len(f)), f, digest) > 0


@pytest.mark.parametrize("width", [2, 4])
@pytest.mark.parametrize("name", FIXTURES)
def test_mutate_2_3_min_max_words(dec, golden_dir, name, width):
    """test/src/mutate_2.rs / mutate_3.rs: every 2- and 4-byte window forced to 0x00.. and 0xFF.."""
    f, digest = _fixture(golden_dir, name)
    starts = np.arange(len(f) - width + 1)
    n = 2 * len(starts)
    frames = np.tile(f, (n, 1))
    for k in range(width):
        frames[np.arange(0, n, 2), starts + k] = 0x00
        frames[np.arange(1, n, 2), starts + k] = 0xFF
    assert _check(dec, frames, np.full(n, len(f)), f, digest) > 0


@pytest.mark.parametrize("bytewise", [False, True])
@pytest.mark.parametrize("name", FIXTURES)
def test_mutate_4_5_compound_random(dec, golden_dir, name, bytewise):
    """test/src/mutate_4.rs / mutate_5.rs: 256 seeds x 256 cumulative random bit (byte) mutations."""
    f, digest = _fixture(golden_dir, name)
    seeds, steps = 256, 256
    rng = np.random.default_rng(0xC0FFEE + bytewise)
    frames = np.empty((seeds * steps, len(f)), np.uint8)
    cur = np.tile(f, (seeds, 1))
    rows = np.arange(seeds)
    for s in range(steps):
        where = rng.integers(0, len(f), seeds)
        if bytewise:
            cur[rows, where] = rng.integers(0, 256, seeds, dtype=np.uint8)
        else:
            cur[rows, where] ^= (1 << rng.integers(0, 8, seeds)).astype(np.uint8)
        frames[s * seeds:(s + 1) * seeds] = cur
    assert _check(dec, frames, np.full(len(frames), len(f)), f, digest) > 0


@pytest.mark.parametrize("name", FIXTURES)
def test_mutate_6_7_truncated_and_twinned(dec, golden_dir, name):
    """test/src/mutate_6.rs: every proper prefix fails; mutate_7.rs: frame + any non-empty prefix of itself fails."""
    f, digest = _fixture(golden_dir, name)
    L = len(f)
    twin = np.concatenate([f, f])
    frames = np.tile(twin, (2 * L - 1, 1))
    lens = np.concatenate([np.arange(0, L - 1), np.arange(L + 1, 2 * L + 1)])  # mutate_6: [..index], index < len - 1
    assert len(lens) == len(frames)
    _check(dec, frames, lens, f, digest, must_fail=True)
