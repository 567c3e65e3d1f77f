#!/usr/bin/env python3
"""bench.py -- headline benchmark: batched LZFSE decode (and encode) of synthetic text in 64 KiB streams.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--chunks C]

Workload (BASELINE.json configs[1]/[2]): C = 16384 independent 64 KiB chunks of
synthetic text per GPU (1 GiB), frames produced by this repo's GPU encoder (bit-identical to the
reference encoder's frames, tests/test_gpu_encode.py).  One step = one batched decode of all frames.
`value` = uncompressed GB/s with frames resident in HBM; `e2e` = the same through the C-ABI host entry
point with pinned host buffers (H2D of the frames and D2H of the output inside the timed region).
The `encode` object reports the encode direction of the same workload.  Multi-GPU: every rank owns its
own 1 GiB of streams (weak scaling), no collective on the data path; time = max over ranks.

--impl reference times the reference's CPU algorithm (the C port under oracle/, all host threads) on a
bounded sample of the same workload."""
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

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.05)
        self.proc.terminate()
        self.t.join(timeout=2)
        inside = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.03 and len(r) >= 6]
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
# reference arm: the reference's CPU algorithm (oracle port), all host threads, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_decode_bench(raw, n, threads, repeat=1):
    """Encodes (untimed) then decodes n chunks with the oracle; returns (decode GB/s, encode GB/s, ratio)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob

    lib = ob.lib()
    u64 = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
    P64, P32 = C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    bound = lib.orc_encode_bound(CHUNK)
    src_off, src_len = u64(np.arange(n) * CHUNK), u64(np.full(n, CHUNK))
    comp = np.empty(n * bound, dtype=np.uint8)
    c_off, c_cap, c_len, st = u64(np.arange(n) * bound), u64(np.full(n, bound)), np.zeros(n, np.uint64), np.zeros(n, np.int32)
    p = lambda a, t: a.ctypes.data_as(t)
    t0 = time.perf_counter()
    lib.orc_encode_batch(raw.ctypes.data, p(src_off, P64), p(src_len, P64), comp.ctypes.data, p(c_off, P64), p(c_cap, P64), p(c_len, P64), p(st, P32), n, threads)
    t_enc = time.perf_counter() - t0
    assert not st.any()
    out = np.empty(n * CHUNK, dtype=np.uint8)
    o_len, st2 = np.zeros(n, np.uint64), np.zeros(n, np.int32)
    best = 1e30
    for _ in range(repeat):
        t0 = time.perf_counter()
        lib.orc_decode_batch(comp.ctypes.data, p(c_off, P64), p(c_len, P64), out.ctypes.data, p(src_off, P64), p(src_len, P64), p(o_len, P64), p(st2, P32), n, threads)
        best = min(best, time.perf_counter() - t0)
    assert not st2.any() and np.array_equal(out, raw[: n * CHUNK])
    return n * CHUNK / best / 1e9, n * CHUNK / t_enc / 1e9, n * CHUNK / float(c_len.sum())


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


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bench_support import workload as W

    threads = os.cpu_count() or 1
    n = min(a.chunks, 2048)  # bounded sample: 128 MiB of the same workload
    pool, woff = reference_pool()
    raw = W.text_chunks(pool, woff, n, CHUNK)
    for _ in range(a.warmup):
        cpu_decode_bench(raw, min(n, 256), threads)
    vals, encs = [], []
    t0 = time.perf_counter()
    for _ in range(a.steps):
        d, e, ratio = cpu_decode_bench(raw, n, threads)
        vals.append(d); encs.append(e)
    dt = (time.perf_counter() - t0) / a.steps
    v = float(np.median(vals))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(n * CHUNK / v / 1e6, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": {"workload": "batched decode of 1 GiB synthetic text split into 64 KiB independent LZFSE streams per GPU (BASELINE.json configs[1])",
                                         "chunk_bytes": CHUNK, "streams_per_gpu": a.chunks, "streams_sampled_per_step": n},
        "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of %d streams (%d MiB), oracle/lzfse_oracle.c -O3, %d threads" % (n, a.chunks, n * CHUNK >> 20, threads)},
        "encode": {"value": round(float(np.median(encs)), 4), "unit": UNIT, "ratio": round(ratio, 4)},
        "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s_per_step": round(dt, 3),
    }))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    import lzfse_rust_b200 as L
    from bench_support import workload as W

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dec, enc = L.LzfseDecoder(local), L.LzfseEncoder(local)
    n = a.chunks
    pool, woff = W.word_pool(dec)
    # every rank owns its own streams: chunk seeds are offset by the rank (host-side scatter, no collective)
    raw_h = torch.empty(n * CHUNK, dtype=torch.uint8).pin_memory()
    W.text_chunks(pool, woff, n, CHUNK, seed0=0x5EED0000 + rank * n, out=raw_h.numpy())
    raw_d = raw_h.to(dev, non_blocking=True)
    i64 = lambda x: torch.tensor(np.asarray(x, dtype=np.int64), device=dev)
    r_off, r_len = i64(np.arange(n) * CHUNK), i64(np.full(n, CHUNK))
    bound = enc.encode_bound(CHUNK)
    comp_d = torch.empty(n * bound, dtype=torch.uint8, device=dev)
    c_off, c_cap = i64(np.arange(n) * bound), i64(np.full(n, bound))

    # ---- encode direction (also produces the frames the decode direction consumes) ----
    enc.set_timing(True)
    enc_steps = max(1, min(a.steps, 3))
    c_len, st = enc.encode_batch_device(raw_d, r_off, r_len, comp_d, c_off, c_cap)  # warm-up + allocation
    assert int((st != 0).sum()) == 0
    enc_ms = []
    for _ in range(enc_steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        c_len, st = enc.encode_batch_device(raw_d, r_off, r_len, comp_d, c_off, c_cap)
        e1.record(); torch.cuda.synchronize()
        enc_ms.append(e0.elapsed_time(e1))
    enc_stage = enc.last_stage_ms()
    enc_launches = enc.last_launches
    # pack the frames tightly (what a caller would store); decode reads this buffer
    c_len_h = c_len.cpu().numpy().astype(np.int64)
    p_off_h = np.concatenate([[0], np.cumsum(c_len_h)[:-1]])
    total_c = int(c_len_h.sum())
    idx = torch.repeat_interleave(i64(np.arange(n) * bound - p_off_h), c_len) + torch.arange(total_c, device=dev)
    packed_d = comp_d[idx].contiguous()
    del idx, comp_d
    p_off, p_len = i64(p_off_h), i64(c_len_h)
    out_d = torch.empty(n * CHUNK, dtype=torch.uint8, device=dev)

    # ---- parity gates (size independent): decode(encode(x)) == x for all streams; oracle spot check on rank 0 ----
    o_len, st = dec.decode_batch_device(packed_d, p_off, p_len, out_d, r_off, r_len)
    assert int((st != 0).sum()) == 0 and bool(torch.equal(out_d, raw_d)), "GPU round trip failed"
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_binding as ob

        packed_h = packed_d.cpu().numpy()
        oenc = ob.Encoder()
        for i in (0, n // 2, n - 1):
            frame = packed_h[p_off_h[i]:p_off_h[i] + c_len_h[i]].tobytes()
            chunk = raw_h.numpy()[i * CHUNK:(i + 1) * CHUNK].tobytes()
            assert oenc.encode(chunk)[1] == frame and ob.decode(frame) == (0, chunk), "oracle spot check failed"

    # ---- decode, frames resident in HBM ----
    U, Cb = n * CHUNK, total_c
    dec.set_timing(True)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(a.warmup):
        dec.decode_batch_device(packed_d, p_off, p_len, out_d, r_off, r_len)
    sampler.wait_first()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {}
    t_wall0 = time.time()
    e0.record()
    for _ in range(a.steps):
        dec.decode_batch_device(packed_d, p_off, p_len, out_d, r_off, r_len)
        for k, v in dec.last_stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop(t_wall0, time.time())
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / a.steps
    launches = dec.last_launches * a.steps
    stage_ms = {k: v / a.steps for k, v in stage_acc.items()}

    # ---- e2e: C-ABI host entry point, pinned host buffers, H2D + D2H inside the timed region ----
    packed_hp = torch.empty(total_c, dtype=torch.uint8).pin_memory(); packed_hp.copy_(packed_d)
    out_hp = torch.empty(U, dtype=torch.uint8).pin_memory()
    ho = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
    h_args = (packed_hp.numpy(), ho(p_off_h), ho(c_len_h), out_hp.numpy(), ho(np.arange(n) * CHUNK), ho(np.full(n, CHUNK)))
    e2e_steps = max(1, min(a.steps, 5))
    dec.set_timing(False)
    dec.decode_batch_into(*h_args)  # warm-up (staging buffers)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hl, hs = dec.decode_batch_into(*h_args)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert not hs.any() and np.array_equal(out_hp.numpy(), raw_h.numpy())
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = peaks()
    dom = max((k for k in stage_ms if k in ("literals", "lmds", "expand")), key=lambda k: stage_ms[k])
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(dom)
    except Exception:
        pass
    achieved = (U + Cb) / (stage_ms[dom] * 1e-3) / 1e9
    cpu = None
    if world == 1 and not a.no_cpu:
        threads = os.cpu_count() or 1
        ns = min(n, 2048)
        d_all, e_all, ratio = cpu_decode_bench(raw_h.numpy(), ns, threads)
        d_one, e_one, _ = cpu_decode_bench(raw_h.numpy(), min(ns, 256), 1)
        cpu = {"value": round(d_all, 4), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %d of %d streams (%d MiB) decoded by oracle/lzfse_oracle.c, %d threads" % (ns, n, ns * CHUNK >> 20, threads),
               "single_thread_value": round(d_one, 4), "encode_value": round(e_all, 4), "encode_single_thread_value": round(e_one, 4), "ratio": round(ratio, 4)}
    enc_ms_step = float(np.median(enc_ms))
    line = {
        "metric": METRIC, "value": round(world * U / (ms_step * 1e-3) / 1e9, 3), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "batched decode of 1 GiB synthetic text split into 64 KiB independent LZFSE streams per GPU (BASELINE.json configs[1])",
                   "streams_per_gpu": n, "chunk_bytes": CHUNK, "uncompressed_bytes_per_gpu": U, "compressed_bytes_per_gpu": Cb,
                   "compression_ratio": round(U / Cb, 4), "l2": "inputs (%.0f MiB) larger than L2, no flush" % ((U + Cb) / 2**20),
                   "frames": "GPU encoder output, byte-identical to the reference encoder's"},
        "roofline": {"bound": "hbm", "kernel": {"literals": "k_fse_literals", "lmds": "k_fse_lmds", "expand": "k_expand"}[dom],
                     "achieved": round(achieved, 2), "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": round(achieved / peak, 5),
                     "traffic": traffic, "algorithmic_bytes": U + Cb, "kernel_ms": round(stage_ms[dom], 4),
                     "whole_step_frac": round((U + Cb) / (ms_step * 1e-3) / 1e9 / peak, 5), "frac_of_8TBps_nominal": round(achieved / 8000.0, 5),
                     # what the kernel really moves (ncu dram bytes of profiles/roofline_traffic.json over the live kernel time):
                     # mostly random 32-byte sectors of match sources that miss L2, see DESIGN.md section 3
                     "traffic_frac_of_peak": round(traffic / (stage_ms[dom] * 1e-3) / 1e9 / peak, 5) if traffic else None},
        "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
        "e2e": {"value": round(world * U / e2e_s / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": Cb + 4 * 8 * n, "d2h_bytes_per_step": U + 12 * n,
                "ms_per_step": round(e2e_s * 1e3, 3), "api": "lzfse_b200_decode_batch_host (pinned host buffers)"},
        "encode": {"value": round(world * U / (enc_ms_step * 1e-3) / 1e9, 3), "unit": UNIT, "ms_per_step": round(enc_ms_step, 3),
                   "stage_ms": {k: round(v, 3) for k, v in enc_stage.items()}, "ratio_vs_reference_encoder": 1.0,
                   "note": "frames are byte-identical to the oracle's (spot-checked here, exhaustively in tests)"},
        "gpu_launches": int(launches + enc_launches * (enc_steps + 1)),
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
    ap.add_argument("--chunks", type=int, default=16384, help="64 KiB streams per GPU (16384 = 1 GiB)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
