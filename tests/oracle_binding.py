"""ctypes binding for the CPU oracle (oracle/liblzfse_oracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ORACLE_DIR, "liblzfse_oracle.so")


def build(force=False):
    src = [os.path.join(_ORACLE_DIR, f) for f in ("lzfse_oracle.c", "lzfse_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"])
    return _SO


class Lmd(C.Structure):
    _fields_ = [("literal_len", C.c_uint32), ("match_len", C.c_uint32), ("match_distance", C.c_uint32)]


NO_L = 0xFFFFFFFF
_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build())
        u8p, szp = C.c_char_p, C.POINTER(C.c_size_t)
        l.orc_decode.argtypes = [u8p, C.c_size_t, C.c_void_p, C.c_size_t, szp]
        l.orc_decode_trace.argtypes = [u8p, C.c_size_t, C.c_void_p, C.c_size_t, szp, C.POINTER(Lmd), C.c_size_t, szp]
        l.orc_probe.argtypes = [u8p, C.c_size_t, szp, szp]
        l.orc_encoder_create.restype = C.c_void_p
        l.orc_encoder_destroy.argtypes = [C.c_void_p]
        l.orc_encode_bound.argtypes = [C.c_size_t]
        l.orc_encode_bound.restype = C.c_size_t
        l.orc_encode.argtypes = [C.c_void_p, u8p, C.c_size_t, C.c_void_p, C.c_size_t, szp]
        l.orc_fse_encode_lmds.argtypes = [C.c_void_p, u8p, C.c_size_t, C.POINTER(Lmd), C.c_size_t, C.c_void_p, C.c_size_t, szp]
        l.orc_vn_encode_lmds.argtypes = [u8p, C.c_size_t, C.POINTER(Lmd), C.c_size_t, C.c_void_p, C.c_size_t, szp]
        l.orc_frontend_lmds.argtypes = [C.c_void_p, u8p, C.c_size_t, C.c_int, C.POINTER(Lmd), C.c_size_t, szp]
        l.orc_vn_op_class.argtypes = [C.c_uint8]
        l.orc_normalize_m1.argtypes = [C.POINTER(C.c_uint16), C.c_size_t, C.c_uint32, C.c_uint32]
        u64p, i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
        for f in (l.orc_decode_batch, l.orc_encode_batch):
            f.argtypes = [C.c_void_p, u64p, u64p, C.c_void_p, u64p, u64p, u64p, i32p, C.c_size_t, C.c_int]
        _lib = l
    return _lib


def probe(frame):
    raw, nb = C.c_size_t(0), C.c_size_t(0)
    st = lib().orc_probe(frame, len(frame), C.byref(raw), C.byref(nb))
    return st, raw.value, nb.value


def decode(frame, cap=None):
    """Returns (status, bytes)."""
    if cap is None:
        st, raw, _ = probe(frame)
        cap = raw if st == 0 else max(4 * len(frame), 1 << 16)
    buf = C.create_string_buffer(max(cap, 1))
    n = C.c_size_t(0)
    st = lib().orc_decode(frame, len(frame), buf, cap, C.byref(n))
    return st, buf.raw[: n.value]


def decode_trace(frame, cap, trace_cap=1 << 22):
    buf = C.create_string_buffer(max(cap, 1))
    tr = (Lmd * trace_cap)()
    n, tn = C.c_size_t(0), C.c_size_t(0)
    st = lib().orc_decode_trace(frame, len(frame), buf, cap, C.byref(n), tr, trace_cap, C.byref(tn))
    return st, buf.raw[: n.value], [(t.literal_len, t.match_len, t.match_distance) for t in tr[: tn.value]]


class Encoder:
    def __init__(self):
        self.h = lib().orc_encoder_create()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_encoder_destroy(self.h)
            self.h = None

    def encode(self, data, cap=None):
        cap = lib().orc_encode_bound(len(data)) if cap is None else cap
        buf = C.create_string_buffer(max(cap, 1))
        n = C.c_size_t(0)
        st = lib().orc_encode(self.h, data, len(data), buf, cap, C.byref(n))
        return st, buf.raw[: min(n.value, cap)]

    def frontend_lmds(self, data, vn=False, cap=1 << 20):
        arr = (Lmd * cap)()
        n = C.c_size_t(0)
        st = lib().orc_frontend_lmds(self.h, data, len(data), int(vn), arr, cap, C.byref(n))
        return st, [(t.literal_len, t.match_len, t.match_distance) for t in arr[: n.value]]

    def fse_encode_lmds(self, literals, lmds, cap=1 << 22):
        arr = (Lmd * max(len(lmds), 1))(*[Lmd(*t) for t in lmds])
        buf = C.create_string_buffer(cap)
        n = C.c_size_t(0)
        st = lib().orc_fse_encode_lmds(self.h, literals, len(literals), arr, len(lmds), buf, cap, C.byref(n))
        return st, buf.raw[: n.value]


def vn_encode_lmds(literals, lmds, cap=1 << 20):
    arr = (Lmd * max(len(lmds), 1))(*[Lmd(*t) for t in lmds])
    buf = C.create_string_buffer(cap)
    n = C.c_size_t(0)
    st = lib().orc_vn_encode_lmds(literals, len(literals), arr, len(lmds), buf, cap, C.byref(n))
    return st, buf.raw[: n.value]


def encode(data):
    return Encoder().encode(data)
