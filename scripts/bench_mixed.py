#!/usr/bin/env python3
"""BASELINE.json configs[4]: the mixed synthetic corpus (incompressible / LZVN-sized small inputs / highly repetitive /
text), sharded over the ranks of a torchrun launch with no collective on the data path.

  python scripts/bench_mixed.py [--mib-per-class M] [--waves W] [--cpu-mib S]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_mixed.py ...

Every rank owns M MiB per class and wave (weak scaling: rank r, wave w use their own seeds), encodes and decodes each
class as one resident batch with the data in HBM, and checks decode(encode(x)) == x for every stream of every wave; rank 0
also checks sampled frames against the oracle (frame bytes equal, oracle decode equal).  Times are CUDA events, summed
over waves, max over ranks.  The 64 GiB of the config do not fit one GPU with outputs, hence resident waves (SURVEY.md
section 8d); W waves of 4 x M MiB per rank are what this run covers and the JSON line says how much that is.
--cpu-mib S: rank 0 also times the oracle port (all host threads) on the first S MiB of every class."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lzfse_rust_b200 as L  # noqa: E402
from bench_support import workload as W  # noqa: E402

CLASSES = ("text", "noise", "small", "rep")
CHUNK = 65536


def make(kind, total, seed, pool, woff):
    """(raw uint8[], offsets int64[], lengths int64[]) of one class: `total` bytes, deterministic in `seed`."""
    rng = np.random.default_rng(seed)
    if kind == "text":  # as configs[1]: 64 KiB chunks of pool words
        n = total // CHUNK
        raw = W.text_chunks(pool, woff, n, CHUNK, seed0=0x5EED0000 + (seed << 16))
        lens = np.full(n, CHUNK, np.int64)
    elif kind == "noise":  # incompressible 64 KiB chunks
        n = total // CHUNK
        raw = rng.integers(0, 256, total, dtype=np.uint8)
        lens = np.full(n, CHUNK, np.int64)
    elif kind == "small":  # LZVN range 21..4096 bytes of text, with a sprinkle of raw-range inputs (<= 20 bytes)
        lens = (21 + rng.integers(0, 4076, total // 2048)).astype(np.int64)
        lens[rng.integers(0, len(lens), len(lens) // 64)] = rng.integers(0, 21, len(lens) // 64)
        src = W.text_chunks(pool, woff, 1, int(lens.sum()) + 8, seed0=0x1234 + seed)
        raw = src[: int(lens.sum())]
    else:  # highly repetitive: period-p repeats of random bytes, 64 KiB chunks
        n = total // CHUNK
        raw = np.empty(total, np.uint8)
        periods = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 32, 64]
        for i in range(n):
            p = periods[i % len(periods)]
            raw[i * CHUNK:(i + 1) * CHUNK] = np.resize(rng.integers(0, 256, p, dtype=np.uint8), CHUNK)
        lens = np.full(n, CHUNK, np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    return np.ascontiguousarray(raw), offs, lens


def cpu_leg(raw, offs, lens, sample_bytes, threads):
    """Oracle port, all host threads, on the streams that fit in the first `sample_bytes`: (encode GB/s, decode GB/s)."""
    import oracle_binding as ob

    lib = ob.lib()
    n = max(1, int(np.searchsorted(np.cumsum(lens), sample_bytes, side="right")))
    u64 = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
    P64, P32 = C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    p = lambda a, t: a.ctypes.data_as(t)
    s_off, s_len = u64(offs[:n]), u64(lens[:n])
    caps = u64([lib.orc_encode_bound(int(l)) for l in lens[:n]])
    c_off = u64(np.concatenate([[0], np.cumsum(caps)[:-1]]))
    comp = np.empty(int(caps.sum()), np.uint8)
    c_len, st = np.zeros(n, np.uint64), np.zeros(n, np.int32)
    t0 = time.perf_counter()
    lib.orc_encode_batch(raw.ctypes.data, p(s_off, P64), p(s_len, P64), comp.ctypes.data, p(c_off, P64), p(caps, P64), p(c_len, P64), p(st, P32), n, threads)
    t_enc = time.perf_counter() - t0
    assert not st.any()
    U = int(lens[:n].sum())
    out = np.empty(U + 8, np.uint8)
    o_len, st2 = np.zeros(n, np.uint64), np.zeros(n, np.int32)
    t0 = time.perf_counter()
    lib.orc_decode_batch(comp.ctypes.data, p(c_off, P64), p(c_len, P64), out.ctypes.data, p(s_off, P64), p(s_len, P64), p(o_len, P64), p(st2, P32), n, threads)
    t_dec = time.perf_counter() - t0
    assert not st2.any() and np.array_equal(out[:U], raw[:U])
    return U / t_enc / 1e9, U / t_dec / 1e9, U


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib-per-class", type=int, default=1024)
    ap.add_argument("--waves", type=int, default=1)
    ap.add_argument("--cpu-mib", type=int, default=0)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    enc, dec = L.LzfseEncoder(local), L.LzfseDecoder(local)
    enc.set_timing(True); dec.set_timing(True)
    pool, woff = W.word_pool(dec)
    total = a.mib_per_class << 20
    t = lambda x: torch.from_numpy(np.asarray(x, np.int64)).to(dev)
    enc_ms = {k: 0.0 for k in CLASSES}
    dec_ms = {k: 0.0 for k in CLASSES}
    u_bytes = {k: 0 for k in CLASSES}
    c_bytes = {k: 0 for k in CLASSES}
    n_streams = {k: 0 for k in CLASSES}
    cpu = {}
    launches = 0
    for wave in range(a.waves):
        for kind in CLASSES:
            raw, offs, lens = make(kind, total, 1 + rank * 1000 + wave, pool, woff)
            n = len(lens)
            d_raw = torch.from_numpy(raw).to(dev)
            ulens, inv = np.unique(lens, return_inverse=True)
            caps = np.array([enc.encode_bound(int(l)) for l in ulens], np.int64)[inv]
            coff = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
            d_comp = torch.empty(int(caps.sum()), dtype=torch.uint8, device=dev)
            d_offs, d_lens, d_coff, d_caps = t(offs), t(lens), t(coff), t(caps)
            reps = 2 if wave == 0 else 1  # the first call of a class sizes the handle's scratch
            for it in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                c_len, st = enc.encode_batch_device(d_raw, d_offs, d_lens, d_comp, d_coff, d_caps)
                e1.record(); torch.cuda.synchronize()
            enc_ms[kind] += e0.elapsed_time(e1)
            if rank == 0: sys.stderr.write("%s encode stages %s\n" % (kind, {k: round(v, 2) for k, v in enc.last_stage_ms().items() if v > 0.01}))
            launches += enc.last_launches
            assert int((st != 0).sum()) == 0, "encode status"
            d_out = torch.zeros(len(raw), dtype=torch.uint8, device=dev)
            d_clen = c_len.to(torch.int64)
            for it in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                out_len, dst = dec.decode_batch_device(d_comp, d_coff, d_clen, d_out, d_offs, d_lens)
                e1.record(); torch.cuda.synchronize()
            dec_ms[kind] += e0.elapsed_time(e1)
            if rank == 0: sys.stderr.write("%s decode stages %s\n" % (kind, {k: round(v, 2) for k, v in dec.last_stage_ms().items() if v > 0.01}))
            launches += dec.last_launches
            assert int((dst != 0).sum()) == 0 and bool(torch.equal(d_out, d_raw)) and bool(torch.equal(out_len.to(torch.int64), d_lens)), "round trip"
            u_bytes[kind] += len(raw); c_bytes[kind] += int(c_len.sum()); n_streams[kind] += n
            if rank == 0 and wave == 0:  # oracle spot check: same frame bytes, oracle decodes them to the input
                import oracle_binding as ob
                oenc = ob.Encoder()
                c_len_h = c_len.cpu().numpy()
                for i in (0, n // 3, n - 1):
                    frame = d_comp[coff[i]:coff[i] + int(c_len_h[i])].cpu().numpy().tobytes()
                    chunk = raw[offs[i]:offs[i] + lens[i]].tobytes()
                    assert oenc.encode(chunk)[1] == frame and ob.decode(frame) == (0, chunk), "oracle spot check (%s, stream %d)" % (kind, i)
                if a.cpu_mib:
                    e, d, u = cpu_leg(raw, offs, lens, a.cpu_mib << 20, os.cpu_count() or 1)
                    cpu[kind] = {"encode": round(e, 4), "decode": round(d, 4), "sample_bytes": u}
            del d_raw, d_comp, d_out
    vec = torch.tensor([enc_ms[k] for k in CLASSES] + [dec_ms[k] for k in CLASSES], device=dev, dtype=torch.float64)
    tot = torch.tensor([u_bytes[k] for k in CLASSES] + [c_bytes[k] for k in CLASSES] + [n_streams[k] for k in CLASSES], device=dev, dtype=torch.int64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        v, s = vec.cpu().numpy(), tot.cpu().numpy()
        per = {}
        for i, k in enumerate(CLASSES):
            per[k] = {"streams": int(s[8 + i]), "uncompressed_bytes": int(s[i]), "ratio": round(float(s[i]) / float(s[4 + i]), 3),
                      "encode_GBps": round(float(s[i]) / v[i] / 1e6, 2), "decode_GBps": round(float(s[i]) / v[4 + i] / 1e6, 2),
                      "encode_ms": round(float(v[i]), 2), "decode_ms": round(float(v[4 + i]), 2)}
        U = float(s[:4].sum())
        line = {"workload": "mixed corpus (BASELINE.json configs[4]): 4 equal classes, resident waves", "n_gpus": world, "waves": a.waves,
                "mib_per_class_per_gpu_per_wave": a.mib_per_class, "uncompressed_bytes_total": int(U), "scaling": "weak",
                "decode_GBps": round(U / float(v[4:].sum()) / 1e6, 2), "encode_GBps": round(U / float(v[:4].sum()) / 1e6, 2),
                "ratio": round(U / float(s[4:8].sum()), 3), "classes": per, "gpu_launches": launches,
                "parity": "decode(encode(x)) == x on every stream; sampled frames byte-identical to the oracle encoder's and decoded by the oracle"}
        if cpu:
            line["cpu_baseline"] = {"kind": "port", "cores": os.cpu_count(), "unit": "GB/s", "classes": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
