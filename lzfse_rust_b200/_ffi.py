"""ctypes loader for liblzfse_b200.so.  Fails loudly: there is no CPU fallback behind this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LZB_SO") or os.path.join(_HERE, "liblzfse_b200.so")  # LZB_SO: tuning builds only

# Every symbol include/lzfse_b200.h declares (tests check the library exports all of them).
SYMBOLS = [
    "lzfse_b200_version", "lzfse_b200_status_string", "lzfse_b200_decoder_last_error", "lzfse_b200_encoder_last_error",
    "lzfse_b200_decoder_create", "lzfse_b200_decoder_destroy", "lzfse_b200_decode_bytes", "lzfse_b200_decode_batch_device",
    "lzfse_b200_decode_batch_host", "lzfse_b200_decode_probe_batch_device", "lzfse_b200_decode_probe_batch_host",
    "lzfse_b200_decoder_last_launches", "lzfse_b200_encoder_create", "lzfse_b200_encoder_destroy", "lzfse_b200_encode_bound",
    "lzfse_b200_encode_bytes", "lzfse_b200_encode_batch_device", "lzfse_b200_encode_batch_host", "lzfse_b200_encoder_last_launches",
    "lzfse_b200_decoder_set_timing", "lzfse_b200_decoder_last_stage_ms", "lzfse_b200_encoder_set_timing", "lzfse_b200_encoder_last_stage_ms",
    "lzfse_b200_encode_bound_strict", "lzfse_b200_decode_prefix_batch_device", "lzfse_b200_decode_prefix_batch_host",
    "lzfse_b200_decode_batch_device_async", "lzfse_b200_decoder_sync", "lzfse_b200_encode_batch_device_async", "lzfse_b200_encoder_sync",
]

_lib = None


def load(build_if_missing=True):
    """Returns the loaded library; builds it with nvcc when the .so is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and not os.environ.get("LZB_SO"):
        from . import build as _build

        try:
            _build.build()
        except Exception as e:  # no nvcc on this box: use the prebuilt .so if there is one
            if not os.path.exists(SO_PATH):
                raise RuntimeError("liblzfse_b200.so is missing and could not be built: %s" % e)
    if not os.path.exists(SO_PATH):
        raise RuntimeError("liblzfse_b200.so not found at %s (run python -m lzfse_rust_b200.build)" % SO_PATH)
    lib = C.CDLL(SO_PATH)
    vp, u64p, i32p, u32p, szp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)
    lib.lzfse_b200_version.restype = C.c_char_p
    lib.lzfse_b200_status_string.restype = C.c_char_p
    lib.lzfse_b200_status_string.argtypes = [C.c_int]
    for n in ("decoder", "encoder"):
        getattr(lib, "lzfse_b200_%s_last_error" % n).restype = C.c_char_p
        getattr(lib, "lzfse_b200_%s_last_error" % n).argtypes = [vp]
        getattr(lib, "lzfse_b200_%s_create" % n).argtypes = [C.c_int, C.POINTER(vp)]
        getattr(lib, "lzfse_b200_%s_destroy" % n).argtypes = [vp]
        getattr(lib, "lzfse_b200_%s_destroy" % n).restype = None
        getattr(lib, "lzfse_b200_%s_last_launches" % n).argtypes = [vp]
        getattr(lib, "lzfse_b200_%s_last_launches" % n).restype = C.c_uint64
        getattr(lib, "lzfse_b200_%s_set_timing" % n).argtypes = [vp, C.c_int]
        getattr(lib, "lzfse_b200_%s_set_timing" % n).restype = None
        getattr(lib, "lzfse_b200_%s_last_stage_ms" % n).argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    lib.lzfse_b200_encode_bound.argtypes = [C.c_size_t]
    lib.lzfse_b200_encode_bound.restype = C.c_size_t
    lib.lzfse_b200_encode_bound_strict.argtypes = [C.c_size_t]
    lib.lzfse_b200_encode_bound_strict.restype = C.c_size_t
    for n in ("decode", "encode"):
        getattr(lib, "lzfse_b200_%s_bytes" % n).argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, szp]
        getattr(lib, "lzfse_b200_%s_batch_device" % n).argtypes = [vp, vp, u64p, u64p, vp, u64p, u64p, u64p, i32p, C.c_size_t, vp]
        getattr(lib, "lzfse_b200_%s_batch_host" % n).argtypes = [vp, vp, u64p, u64p, vp, u64p, u64p, u64p, i32p, C.c_size_t]
        getattr(lib, "lzfse_b200_%s_batch_device_async" % n).argtypes = [vp, vp, u64p, u64p, vp, u64p, u64p, u64p, i32p, C.c_size_t, vp]
        getattr(lib, "lzfse_b200_%sr_sync" % n).argtypes = [vp]
    lib.lzfse_b200_decode_prefix_batch_device.argtypes = [vp, vp, u64p, u64p, vp, u64p, u64p, u64p, i32p, vp, C.c_size_t, vp]
    lib.lzfse_b200_decode_prefix_batch_host.argtypes = [vp, vp, u64p, u64p, vp, u64p, u64p, u64p, i32p, vp, C.c_size_t]
    lib.lzfse_b200_decode_probe_batch_device.argtypes = [vp, vp, u64p, u64p, u64p, u32p, i32p, C.c_size_t, vp]
    lib.lzfse_b200_decode_probe_batch_host.argtypes = [vp, vp, u64p, u64p, u64p, u32p, i32p, C.c_size_t]
    _lib = lib
    return lib
