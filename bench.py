#!/usr/bin/env python3
"""bench.py -- batched LZFSE decode (and encode) on B200, the workloads of BASELINE.json.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|4|5]

--config 2 (default; BASELINE.json configs[1] and [2]): 16384 independent 64 KiB chunks of synthetic text per GPU
           (1 GiB), weak scaling.  One step = one batched decode of all frames; `encode` reports the other direction.
--config 4 (configs[3]): 8 streams of 16 MiB per GPU (multi-block frames, matches across block boundaries).
--config 5 (configs[4]): the mixed corpus (text / incompressible / LZVN-sized inputs / highly repetitive), ONE global
           descriptor array cut into contiguous ranges by lzfse_rust_b200.sharding.shard_ranges (host-side scatter, no
           collective), every rank working through its range in resident waves.  Strong scaling: --total-gib is the whole
           job (default 64 for N > 1, 8 for N = 1).

Frames are produced by this repo's GPU encoder and checked against the CPU oracle (byte-identical on a sample, every
stream round-trips).  `value` = uncompressed GB/s with the frames resident in HBM (CUDA events, max over ranks);
`e2e` = the same through the C-ABI host entry point with pinned host buffers (H2D of the frames and D2H of the output
inside the timed region), next to a plain-memcpy control of the same byte counts (`pcie_control`) and the same call on
pageable caller memory (`pageable`).

--impl reference times the reference's CPU algorithm (the C port under oracle/, all host threads) on a bounded sample
of the same workload."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = 65536
METRIC = "decode_uncompressed_throughput"
UNIT = "GB/s"
WORKLOADS = {
    2: "batched decode of 1 GiB synthetic text split into 64 KiB independent LZFSE streams per GPU (BASELINE.json configs[1])",
    4: "decode of 8 x 16 MiB synthetic-text streams per GPU: multi-block frames, matches across block boundaries (BASELINE.json configs[3])",
    5: "mixed synthetic corpus (text / incompressible / LZVN-sized small inputs / highly repetitive), one global descriptor array sharded over the GPUs (BASELINE.json configs[4])",
}
KERNEL_OF = {"literals": "k_fse_literals", "lmds": "k_fse_lmds", "expand": "k_expand"}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled in the background; stop(t0, t1) keeps the samples
    taken while the timed region ran (the sampler is started before warm-up so it is up by then)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def wait_first(self, timeout=10.0):
        t = time.time()
        while self.proc and not self.rows and time.time() - t < timeout:
            time.sleep(0.02)

    def stop(self, windows):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.05)
        self.proc.terminate()
        self.t.join(timeout=2)
        inside = [r for ts, r in self.rows if len(r) >= 6 and any(t0 <= ts <= t1 + 0.03 for t0, t1 in windows)]
        window = "timed region"
        if not inside:
            inside, window = [r for _, r in self.rows if len(r) >= 6], "whole run (timed region shorter than the sampling period)"
        num = lambda x: x.replace(".", "", 1).isdigit()
        sm = [float(r[0]) for r in inside if num(r[0])]
        mx = [float(r[1]) for r in inside if num(r[1])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------
# workloads: (raw uint8[], offsets int64[], lengths int64[]) determined by global stream indices
# ------------------------------------------------------------------------------------------------
def _mix_bytes(lo, hi, salt):
    """Incompressible bytes lo..hi of a stream of pseudo-random bytes that only depends on the byte index (splitmix64)."""
    x = np.arange(lo, hi, dtype=np.uint64) + np.uint64(salt)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return ((x ^ (x >> np.uint64(31))) >> np.uint64(56)).astype(np.uint8)


SMALL_SLOT = 4096 + 32   # bytes of text generated per small-class stream; its length picks a prefix
REP_PERIODS = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 32, 64)


MIX_PARTS = 64   # every class is cut into this many parts; the global order takes one part of each class in turn


def mixed_lengths(bytes_per_class):
    """The ONE global descriptor array of config 5: (lens, cls, kidx) = length, class and index inside its class of every
    stream.  The four classes are interleaved part by part (MIX_PARTS parts each), so that a contiguous range of the
    array -- what a rank gets from shard_ranges -- holds the same mix as the whole corpus."""
    n64 = bytes_per_class // CHUNK
    rng = np.random.default_rng(0x5EED)
    small = (21 + rng.integers(0, 4076, bytes_per_class // 2048)).astype(np.int64)   # LZVN range, mean ~2 KiB
    small[rng.integers(0, len(small), len(small) // 64)] = rng.integers(0, 21, len(small) // 64)   # a sprinkle of raw-range inputs
    per_class = [np.full(n64, CHUNK, np.int64), np.full(n64, CHUNK, np.int64), small, np.full(n64, CHUNK, np.int64)]
    lens, cls, kidx = [], [], []
    for part in range(MIX_PARTS):
        for c, lc in enumerate(per_class):
            a, b = len(lc) * part // MIX_PARTS, len(lc) * (part + 1) // MIX_PARTS
            lens.append(lc[a:b]); cls.append(np.full(b - a, c, np.int8)); kidx.append(np.arange(a, b, dtype=np.int64))
    return np.concatenate(lens), np.concatenate(cls), np.concatenate(kidx)


def mixed_data(lens, cls, kidx, lo, hi, pool, woff):
    """Bytes of global streams [lo, hi) of config 5 (every stream's content depends on its class and its index there)."""
    from bench_support import workload as W

    out = np.empty(int(lens[lo:hi].sum()), np.uint8)
    pos = 0
    i = lo
    while i < hi:
        c = int(cls[i])
        j = i
        while j < hi and cls[j] == c and kidx[j] == kidx[i] + (j - i):
            j += 1
        n, nbytes = j - i, int(lens[i:j].sum())
        k0 = int(kidx[i])
        if c == 0:
            W.text_chunks(pool, woff, n, CHUNK, seed0=0x5EED0000 + k0, out=out[pos:pos + nbytes])
        elif c == 1:
            out[pos:pos + nbytes] = _mix_bytes(k0 * CHUNK, (k0 + n) * CHUNK, 0xA11CE)
        elif c == 2:
            slots = W.text_chunks(pool, woff, n, SMALL_SLOT, seed0=0x12340000 + k0).reshape(n, SMALL_SLOT)
            keep = np.arange(SMALL_SLOT)[None, :] < lens[i:j, None]
            out[pos:pos + nbytes] = slots[keep]
        else:
            seeds = _mix_bytes(k0 * 64, (k0 + n) * 64, 0xBEEF).reshape(n, 64)
            idx = np.arange(CHUNK)
            blk = out[pos:pos + nbytes].reshape(n, CHUNK)
            for t in range(n):
                p = REP_PERIODS[(k0 + t) % len(REP_PERIODS)]
                blk[t] = seeds[t, idx % p]
        pos += nbytes
        i = j
    return out


# ------------------------------------------------------------------------------------------------
# CPU leg (oracle port): used by --impl reference and by the cpu_baseline object
# ------------------------------------------------------------------------------------------------
def cpu_codec_bench(raw, offs, lens, threads, repeat=1):
    """Encodes then decodes the given streams with the oracle; returns (decode GB/s, encode GB/s, compressed bytes, frames, c_off, c_len)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob

    lib = ob.lib()
    n = len(lens)
    u64 = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
    P64, P32 = C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    p = lambda a, t: a.ctypes.data_as(t)
    s_off, s_len = u64(offs), u64(lens)
    ul, inv = np.unique(lens, return_inverse=True)
    caps = u64(np.array([lib.orc_encode_bound(int(l)) for l in ul])[inv])
    c_off = u64(np.concatenate([[0], np.cumsum(caps)[:-1]]))
    comp = np.empty(int(caps.sum()), np.uint8)
    c_len, st = np.zeros(n, np.uint64), np.zeros(n, np.int32)
    t0 = time.perf_counter()
    lib.orc_encode_batch(raw.ctypes.data, p(s_off, P64), p(s_len, P64), comp.ctypes.data, p(c_off, P64), p(caps, P64), p(c_len, P64), p(st, P32), n, threads)
    t_enc = time.perf_counter() - t0
    assert not st.any()
    U = int(np.asarray(lens).sum())
    out = np.empty(U + 8, np.uint8)
    d_off = u64(np.concatenate([[0], np.cumsum(np.asarray(lens, np.int64))[:-1]]))
    o_len, st2 = np.zeros(n, np.uint64), np.zeros(n, np.int32)
    best = 1e30
    for _ in range(repeat):
        t0 = time.perf_counter()
        lib.orc_decode_batch(comp.ctypes.data, p(c_off, P64), p(c_len, P64), out.ctypes.data, p(d_off, P64), p(s_len, P64), p(o_len, P64), p(st2, P32), n, threads)
        best = min(best, time.perf_counter() - t0)
    assert not st2.any()
    if n and np.array_equal(np.asarray(offs, np.int64) - int(offs[0]), d_off.astype(np.int64)):   # contiguous streams: one compare
        assert np.array_equal(out[:U], raw[int(offs[0]):int(offs[0]) + U]), "oracle round trip"
    return U / best / 1e9, U / t_enc / 1e9, int(c_len.sum()), comp, c_off, c_len


def reference_pool():
    """Word pool without the GPU (the reference arm must not touch our kernels): oracle-decoded fixtures."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from bench_support import workload as W

    toks = []
    for name in W.TEXT_FIXTURES:
        st, raw = ob.decode(open(os.path.join(ROOT, "tests", "golden", "data", "snappy", name + ".lzfse"), "rb").read())
        assert st == 0
        toks.extend(raw.split())
    off = np.zeros(len(toks) + 1, dtype=np.uint32)
    np.cumsum([len(t) for t in toks], out=off[1:])
    return np.frombuffer(b"".join(toks), dtype=np.uint8).copy(), off


def sample_workload(a, pool, woff, budget_bytes):
    """A bounded sample of the config's workload for the CPU legs: (raw, offs, lens, description)."""
    from bench_support import workload as W

    if a.config == 2:
        n = min(a.chunks, budget_bytes // CHUNK)
        raw = W.text_chunks(pool, woff, n, CHUNK)
        return raw, np.arange(n, dtype=np.int64) * CHUNK, np.full(n, CHUNK, np.int64), "first %d of %d streams (%d MiB)" % (n, a.chunks, n * CHUNK >> 20)
    if a.config == 4:
        n, cl = a.streams, a.stream_mib << 20
        raw = np.concatenate([W.text_chunks(pool, woff, 1, cl, seed0=0x16000000 + i) for i in range(n)])
        return raw, np.arange(n, dtype=np.int64) * cl, np.full(n, cl, np.int64), "all %d streams of %d MiB (one host thread per stream)" % (n, a.stream_mib)
    lens, cls, kidx = mixed_lengths(budget_bytes // 4)
    raw = mixed_data(lens, cls, kidx, 0, len(lens), pool, woff)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    return raw, offs, lens, "the same four classes at %d MiB per class (%d streams)" % (budget_bytes // 4 >> 20, len(lens))


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pool, woff = reference_pool()
    raw, offs, lens, what = sample_workload(a, pool, woff, 128 << 20)
    threads = min(os.cpu_count() or 1, len(lens))
    U = int(lens.sum())
    for _ in range(a.warmup):
        k = max(1, len(lens) // 8)
        cpu_codec_bench(raw, offs[:k], lens[:k], threads)
    vals, encs = [], []
    t0 = time.perf_counter()
    for _ in range(a.steps):
        d, e, cb, _, _, _ = cpu_codec_bench(raw, offs, lens, threads)
        vals.append(d); encs.append(e)
    dt = (time.perf_counter() - t0) / a.steps
    v = float(np.median(vals))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(U / v / 1e6, 3), "higher_is_better": True, "scaling": "strong" if a.config == 5 else "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": {"workload": WORKLOADS[a.config], "bench_config": a.config, "sampled": what},
        "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%s, oracle/lzfse_oracle.c -O3, %d threads" % (what, threads)},
        "encode": {"value": round(float(np.median(encs)), 4), "unit": UNIT, "ratio": round(U / cb, 4)},
        "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s_per_step": round(dt, 3),
    }))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Batch:
    """One resident batch: raw bytes in pinned host memory and on the device, frames produced by the GPU encoder."""

    def __init__(self, torch, dev, enc, raw_np, offs, lens):
        self.torch, self.dev, self.n = torch, dev, len(lens)
        self.offs, self.lens = np.asarray(offs, np.int64), np.asarray(lens, np.int64)
        self.U = int(self.lens.sum())
        self.raw_h = torch.empty(max(self.U, 1), dtype=torch.uint8).pin_memory()
        self.raw_h.numpy()[: self.U] = raw_np[: self.U]
        self.raw_d = self.raw_h.to(dev)
        i64 = lambda x: torch.tensor(np.asarray(x, dtype=np.int64), device=dev)
        self.i64 = i64
        self.r_off, self.r_len = i64(self.offs), i64(self.lens)
        ul, inv = np.unique(self.lens, return_inverse=True)
        self.caps = np.array([enc.encode_bound(int(l)) for l in ul], np.int64)[inv]
        self.coff = np.concatenate([[0], np.cumsum(self.caps)[:-1]]).astype(np.int64)
        self.comp_d = torch.empty(int(self.caps.sum()), dtype=torch.uint8, device=dev)
        self.c_off, self.c_cap = i64(self.coff), i64(self.caps)

    def encode(self, enc, steps):
        torch = self.torch
        enc.set_timing(True)
        c_len, st = enc.encode_batch_device(self.raw_d, self.r_off, self.r_len, self.comp_d, self.c_off, self.c_cap)  # warm-up + allocation
        assert int((st != 0).sum()) == 0
        ms = []
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            c_len, st = enc.encode_batch_device(self.raw_d, self.r_off, self.r_len, self.comp_d, self.c_off, self.c_cap)
            e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        self.enc_ms, self.enc_stage, self.enc_launches = float(np.median(ms)), enc.last_stage_ms(), enc.last_launches * (steps + 1)
        # pack the frames tightly (what a caller would store); decode reads this buffer
        self.c_len_h = c_len.cpu().numpy().astype(np.int64)
        self.p_off_h = np.concatenate([[0], np.cumsum(self.c_len_h)[:-1]]).astype(np.int64)
        self.Cb = int(self.c_len_h.sum())
        idx = torch.repeat_interleave(self.i64(self.coff - self.p_off_h), c_len.to(torch.int64)) + torch.arange(self.Cb, device=self.dev)
        self.packed_d = self.comp_d[idx].contiguous()
        del idx
        self.comp_d = None
        self.p_off, self.p_len = self.i64(self.p_off_h), self.i64(self.c_len_h)
        self.out_d = torch.empty(max(self.U, 1), dtype=torch.uint8, device=self.dev)

    def parity(self, dec, check_oracle, sample=256):
        """decode(encode(x)) == x for all streams; `sample` frames byte-identical to the oracle encoder's.  Returns the
        compressed size of the sampled streams under both encoders."""
        torch = self.torch
        o_len, st = dec.decode_batch_device(self.packed_d, self.p_off, self.p_len, self.out_d, self.r_off, self.r_len)
        assert int((st != 0).sum()) == 0 and bool(torch.equal(self.out_d[: self.U], self.raw_d[: self.U])), "GPU round trip failed"
        assert bool(torch.equal(o_len.to(torch.int64), self.r_len))
        if not check_oracle:
            return None
        pick = np.unique(np.linspace(0, self.n - 1, min(sample, self.n)).astype(np.int64))
        raw = self.raw_h.numpy()
        _, _, cb, comp, c_off, c_len = cpu_codec_bench(raw, self.offs[pick], self.lens[pick], os.cpu_count() or 1)
        packed_h = self.packed_d.cpu().numpy()
        for k, i in enumerate(pick):
            ours = packed_h[self.p_off_h[i]:self.p_off_h[i] + self.c_len_h[i]]
            theirs = comp[int(c_off[k]):int(c_off[k]) + int(c_len[k])]
            assert len(ours) == len(theirs) and np.array_equal(ours, theirs), "frame %d differs from the oracle encoder's" % i
        return {"frames_compared": int(len(pick)), "gpu_bytes": int(self.c_len_h[pick].sum()), "oracle_bytes": int(cb)}

    def decode_resident(self, dec, steps, warmup, barrier):
        torch = self.torch
        dec.set_timing(True)
        for _ in range(warmup):
            dec.decode_batch_device(self.packed_d, self.p_off, self.p_len, self.out_d, self.r_off, self.r_len)
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = {}
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            dec.decode_batch_device(self.packed_d, self.p_off, self.p_len, self.out_d, self.r_off, self.r_len)
            for k, v in dec.last_stage_ms().items():
                acc[k] = acc.get(k, 0.0) + v
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        return e0.elapsed_time(e1) / steps, {k: v / steps for k, v in acc.items()}, dec.last_launches * steps, (t0, t1)

    def decode_host(self, dec, steps, barrier, pinned=True):
        """Seconds per call of the C-ABI host entry point (H2D + kernels + D2H inside)."""
        torch = self.torch
        if pinned:
            src = torch.empty(self.Cb, dtype=torch.uint8).pin_memory(); src.copy_(self.packed_d)
            dst = torch.empty(max(self.U, 1), dtype=torch.uint8).pin_memory()
            src_np, dst_np = src.numpy(), dst.numpy()
        else:
            src_np, dst_np = self.packed_d.cpu().numpy().copy(), np.empty(max(self.U, 1), np.uint8)
            dst_np[:: 4096] = 0  # touch the pages: what is timed is the copy, not the first page fault
        ho = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
        args = (src_np, ho(self.p_off_h), ho(self.c_len_h), dst_np, ho(self.offs), ho(self.lens))
        dec.set_timing(False)
        dec.decode_batch_into(*args)  # warm-up (staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            hl, hs = dec.decode_batch_into(*args)
        dt = (time.perf_counter() - t0) / steps
        assert not hs.any() and np.array_equal(dst_np[: self.U], self.raw_h.numpy()[: self.U])
        return dt

    def pcie_control(self, steps, barrier):
        """Plain pinned cudaMemcpyAsync of the same byte counts, H2D and D2H on two streams at once."""
        torch = self.torch
        src = torch.empty(self.Cb, dtype=torch.uint8).pin_memory()
        dst = torch.empty(max(self.U, 1), dtype=torch.uint8).pin_memory()
        d_in = torch.empty(self.Cb, dtype=torch.uint8, device=self.dev)
        s_in, s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        def once():
            with torch.cuda.stream(s_in):
                d_in.copy_(src, non_blocking=True)
            with torch.cuda.stream(s_out):
                dst.copy_(self.out_d, non_blocking=True)
            s_in.synchronize(); s_out.synchronize()
        once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            once()
        return (time.perf_counter() - t0) / steps


def host_memcpy_gbps(nbytes=256 << 20):
    """Plain host-to-host copy of pageable memory (numpy): what any staging of a caller's pageable buffer is bound by."""
    a, b = np.ones(nbytes, np.uint8), np.empty(nbytes, np.uint8)
    b[:: 4096] = 0
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter(); np.copyto(b, a); best = min(best, time.perf_counter() - t0)
    return nbytes / best / 1e9


def run_ours(a):
    import torch
    import torch.distributed as dist

    import lzfse_rust_b200 as L
    from lzfse_rust_b200.sharding import shard_ranges
    from bench_support import workload as W

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)
    dec, enc = L.LzfseDecoder(local), L.LzfseEncoder(local)
    pool, woff = W.word_pool(dec)

    # ---- the rank's streams, as a list of waves (raw, offs, lens) ----
    extra = {}
    if a.config == 2:
        n = a.chunks
        waves = [lambda: (W.text_chunks(pool, woff, n, CHUNK, seed0=0x5EED0000 + rank * n), np.arange(n, dtype=np.int64) * CHUNK, np.full(n, CHUNK, np.int64))]
        extra = {"streams_per_gpu": n, "chunk_bytes": CHUNK}
        scaling = "weak"
    elif a.config == 4:
        n, cl = a.streams, a.stream_mib << 20
        waves = [lambda: (np.concatenate([W.text_chunks(pool, woff, 1, cl, seed0=0x16000000 + rank * n + i) for i in range(n)]),
                          np.arange(n, dtype=np.int64) * cl, np.full(n, cl, np.int64))]
        extra = {"streams_per_gpu": n, "stream_bytes": cl}
        scaling = "weak"
    else:
        total = int((a.total_gib if a.total_gib else (64 if world > 1 else 8)) * (1 << 30))
        lens, cls, kidx = mixed_lengths(total // 4)
        lo, hi = shard_ranges(lens, world)[rank]   # weights: bytes each stream moves (its uncompressed size; C is not known yet)
        wave_bytes = int(a.wave_gib * (1 << 30))
        cum = np.cumsum(lens[lo:hi])
        cuts = [lo] + [lo + int(np.searchsorted(cum, wave_bytes * k, side="left")) for k in range(1, int(cum[-1] // wave_bytes) + 1)] + [hi]
        cuts = sorted(set(cuts))
        def mk(i0, i1):
            def f():
                raw = mixed_data(lens, cls, kidx, i0, i1, pool, woff)
                ln = lens[i0:i1]
                return raw, np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.int64), ln
            return f
        waves = [mk(cuts[k], cuts[k + 1]) for k in range(len(cuts) - 1)]
        extra = {"total_uncompressed_bytes": int(lens.sum()), "global_streams": int(len(lens)), "rank0_range": [int(lo), int(hi)], "waves_per_rank": len(waves),
                 "sharding": "lzfse_rust_b200.sharding.shard_ranges over one global descriptor array (the four classes interleaved in %d parts); contiguous ranges, no collective" % MIX_PARTS}
        scaling = "strong"

    sampler = ClockSampler(local)
    sampler.start()
    enc_steps = max(1, min(a.steps, 3))
    e2e_steps = max(1, min(a.steps, 5))
    tot = {"U": 0, "Cb": 0, "n": 0, "dec_ms": 0.0, "enc_ms": 0.0, "e2e_s": 0.0, "page_s": 0.0, "ctl_s": 0.0, "launches": 0}
    stage_ms, enc_stage, windows, parity = {}, {}, [], None
    for w, make in enumerate(waves):
        raw, offs, lens_w = make()
        b = Batch(torch, dev, enc, raw, offs, lens_w)
        del raw
        b.encode(enc, enc_steps)
        p = b.parity(dec, check_oracle=(rank == 0 and w == 0))
        parity = p or parity
        if w == 0:
            sampler.wait_first()
        ms, st, launches, win = b.decode_resident(dec, a.steps, a.warmup, barrier)
        windows.append(win)
        tot["dec_ms"] += ms; tot["enc_ms"] += b.enc_ms; tot["U"] += b.U; tot["Cb"] += b.Cb; tot["n"] += b.n
        tot["launches"] += launches + b.enc_launches
        for k, v in st.items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v
        for k, v in b.enc_stage.items():
            enc_stage[k] = enc_stage.get(k, 0.0) + v
        if w == 0 or a.config != 5:   # host-buffer legs: first wave only for the multi-wave config (said in the line)
            tot["e2e_s"] += b.decode_host(dec, e2e_steps, barrier, pinned=True)
            tot["ctl_s"] += b.pcie_control(e2e_steps, barrier)
            tot["page_s"] += b.decode_host(dec, max(1, e2e_steps // 2), barrier, pinned=False)
            e2e_U, e2e_Cb, e2e_n = b.U, b.Cb, b.n
        if w + 1 < len(waves):
            del b
            torch.cuda.empty_cache()
    clocks = sampler.stop(windows)

    # ---- reduce over ranks: times MAX, bytes SUM ----
    tv = torch.tensor([tot["dec_ms"], tot["enc_ms"], tot["e2e_s"], tot["page_s"], tot["ctl_s"]], device=dev, dtype=torch.float64)
    bv = torch.tensor([tot["U"], tot["Cb"], tot["n"], tot["launches"], e2e_U], device=dev, dtype=torch.int64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        dist.all_reduce(bv, op=dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    dec_ms, enc_ms, e2e_s, page_s, ctl_s = [float(x) for x in tv.cpu()]
    U_all, Cb_all, n_all, launches_all, e2e_U_all = [int(x) for x in bv.cpu()]
    peak, peak_kind = peaks()
    dom = max((k for k in stage_ms if k in KERNEL_OF), key=lambda k: stage_ms[k])
    traffic = None
    if a.config == 2 and a.chunks == 16384:
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(dom)
        except Exception:
            pass
    U, Cb = tot["U"], tot["Cb"]   # rank 0's own bytes: the kernel figures below are per launch on one GPU
    achieved = (U + Cb) / (stage_ms[dom] * 1e-3) / 1e9
    cpu = None
    if world == 1 and not a.no_cpu:
        raw_s, offs_s, lens_s, what = sample_workload(a, pool, woff, 128 << 20)
        threads = min(os.cpu_count() or 1, len(lens_s))
        d_all, e_all, cb_s, _, _, _ = cpu_codec_bench(raw_s, offs_s, lens_s, threads)
        k1 = max(1, len(lens_s) // 8) if a.config != 4 else 1
        d_one, e_one, _, _, _, _ = cpu_codec_bench(raw_s, offs_s[:k1], lens_s[:k1], 1)
        cpu = {"value": round(d_all, 4), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%s decoded by oracle/lzfse_oracle.c, %d threads" % (what, threads),
               "single_thread_value": round(d_one, 4), "encode_value": round(e_all, 4), "encode_single_thread_value": round(e_one, 4),
               "ratio": round(int(lens_s.sum()) / cb_s, 4),
               "note": "C port of lzfse_rust; on BASELINE configs[0] (scripts/cpu_anchor.py) it runs above the crate's published i5-2500K decode figure and at ~0.8x of its encode figure on this pool's host CPU"}
    config = {"workload": WORKLOADS[a.config], "bench_config": a.config, "uncompressed_bytes_rank0": U, "compressed_bytes_rank0": Cb,
              "uncompressed_bytes_all_ranks": U_all, "compression_ratio": round(U / Cb, 4),
              "l2": "inputs (%.0f MiB per launch) larger than L2, no flush" % ((U + Cb) / max(len(waves), 1) / 2**20),
              "frames": "GPU encoder output, byte-identical to the oracle encoder's on %d sampled streams; every stream round-trips" % (parity["frames_compared"] if parity else 0),
              "encode_GBps": round(U_all / (enc_ms * 1e-3) / 1e9, 3)}
    config.update(extra)
    if len(waves) > 1:
        config["timing"] = "every wave is timed on its own (CUDA events around K steps) and the per-step times are added; max over ranks"
    line = {
        "metric": METRIC, "value": round(U_all / (dec_ms * 1e-3) / 1e9, 3), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(dec_ms, 4), "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config,
        "roofline": {"bound": "hbm", "kernel": KERNEL_OF[dom], "achieved": round(achieved, 2), "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                     "frac": round(achieved / peak, 5), "traffic": traffic, "algorithmic_bytes": U + Cb, "kernel_ms": round(stage_ms[dom], 4),
                     "whole_step_frac": round((U + Cb) / (tot["dec_ms"] * 1e-3) / 1e9 / peak, 5), "frac_of_8TBps_nominal": round(achieved / 8000.0, 5),
                     "traffic_frac_of_peak": round(traffic / (stage_ms[dom] * 1e-3) / 1e9 / peak, 5) if traffic else None},
        "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
        "e2e": {"value": round(e2e_U_all / e2e_s / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": e2e_Cb + 4 * 8 * e2e_n, "d2h_bytes_per_step": e2e_U + 12 * e2e_n,
                "ms_per_step": round(e2e_s * 1e3, 3), "api": "lzfse_b200_decode_batch_host (pinned host buffers)",
                "scope": "first wave of every rank" if len(waves) > 1 else "the whole batch",
                "pcie_control": {"value": round(e2e_U_all / ctl_s / 1e9, 3), "unit": UNIT, "ms_per_step": round(ctl_s * 1e3, 3),
                                 "what": "plain pinned cudaMemcpyAsync of the same byte counts (H2D of the frames and D2H of the output on two streams at once), all ranks at the same time: the ceiling of the host path"},
                "frac_of_pcie_control": round(ctl_s / e2e_s, 4),
                "pageable": {"value": round(e2e_U_all / page_s / 1e9, 3), "unit": UNIT, "ms_per_step": round(page_s * 1e3, 3),
                             "what": "the same call on pageable caller buffers (what a Vec<u8> is)",
                             "host_memcpy_GBps": round(host_memcpy_gbps(), 2),
                             "note": "bytes of a pageable buffer cross host memory once more on their way to / from pinned staging: bound by the host's own copy rate (host_memcpy_GBps, one thread; more threads do not raise it on this VM)"}},
        "encode": {"value": round(U_all / (enc_ms * 1e-3) / 1e9, 3), "unit": UNIT, "ms_per_step": round(enc_ms, 3),
                   "stage_ms": {k: round(v, 3) for k, v in enc_stage.items()},
                   "ratio_vs_reference_encoder": round(parity["gpu_bytes"] / parity["oracle_bytes"], 6) if parity else None,
                   "frames_compared_with_oracle": parity["frames_compared"] if parity else 0,
                   "note": "ratio = compressed bytes of the sampled streams, GPU encoder / oracle encoder (the frames are byte-identical)"},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5], help="BASELINE.json workload (see the module docstring)")
    ap.add_argument("--chunks", type=int, default=16384, help="config 2: 64 KiB streams per GPU (16384 = 1 GiB)")
    ap.add_argument("--streams", type=int, default=8, help="config 4: streams per GPU")
    ap.add_argument("--stream-mib", type=int, default=16, help="config 4: MiB per stream")
    ap.add_argument("--total-gib", type=float, default=0, help="config 5: whole job in GiB (default 64 for N > 1, 8 for N = 1)")
    ap.add_argument("--wave-gib", type=float, default=1.0, help="config 5: resident wave per rank in GiB")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
