#!/usr/bin/env python3
"""BASELINE.json configs[0]: the reference's own CPU-runnable case (bench/src/bench.rs:182,195-209, `snap_uflat00_html`):
encode + decode of the 102 400-byte Snappy html payload on ONE host thread, objects reused between iterations like the
criterion loop does.  lzfse_rust cannot be built in this image (no rustc/cargo), so what is timed is the C port under
oracle/; its README figures on an i5-2500K are 945.7 MB/s decode and 118.9 MB/s encode (README.md:155,166).

  python scripts/cpu_anchor.py [--seconds 2.0]

Prints one JSON line.  The decode output is checked against tests/golden/data/snappy/html.hash first."""
import argparse, ctypes as C, hashlib, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=2.0)
ap.add_argument("--name", default="html")
a = ap.parse_args()
d = os.path.join(ROOT, "tests", "golden", "data", "snappy")
frame = open(os.path.join(d, a.name + ".lzfse"), "rb").read()
digest = open(os.path.join(d, a.name + ".hash"), "rb").read()
st, raw = ob.decode(frame)
assert st == 0 and hashlib.sha256(raw).digest() == digest, "oracle decode does not reproduce the fixture's SHA-256"
lib = ob.lib()
enc = lib.orc_encoder_create()
cap = lib.orc_encode_bound(len(raw))
dst = (C.c_uint8 * cap)()
n = C.c_size_t(0)


def loop(fn):
    fn()
    k, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < a.seconds:
        fn(); k += 1
    return (time.perf_counter() - t0) / k


t_enc = loop(lambda: lib.orc_encode(enc, raw, len(raw), dst, cap, C.byref(n)))
own = bytes(dst[: n.value])
assert ob.decode(own) == (0, raw)
out = (C.c_uint8 * (len(raw) + 64))()
m = C.c_size_t(0)
t_dec_own = loop(lambda: lib.orc_decode(own, len(own), out, len(raw), C.byref(m)))
t_dec_c = loop(lambda: lib.orc_decode(frame, len(frame), out, len(raw), C.byref(m)))
lib.orc_encoder_destroy(enc)
print(json.dumps({
    "config": "BASELINE.json configs[0]: snap_uflat00_%s, %d bytes, one thread, oracle/lzfse_oracle.c (C port of lzfse_rust)" % (a.name, len(raw)),
    "encode_MBps": round(len(raw) / t_enc / 1e6, 1), "decode_MBps_own_frame": round(len(raw) / t_dec_own / 1e6, 1),
    "decode_MBps_apple_frame": round(len(raw) / t_dec_c / 1e6, 1), "ratio": round(len(raw) / len(own), 4),
    "readme_i5_2500K": {"encode_MBps": 118.9, "decode_MBps": 945.7, "source": "/root/reference/README.md:155,166"},
    "cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t") if os.path.exists("/proc/cpuinfo") else None,
}))
