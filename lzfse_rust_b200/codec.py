"""Host-side mirror of lzfse_rust's memory-buffer API over the C-ABI (include/lzfse_b200.h).

Same names and argument meaning as the reference:
  encode_bytes(src, dst) / LzfseEncoder.encode_bytes   (lzfse_rust src/encode/mod.rs:58, encoder.rs:49)
  decode_bytes(src, dst) / LzfseDecoder.decode_bytes   (src/decode/mod.rs:49, decoder.rs:61)
`dst` is a bytearray that is appended to; the return value is the number of bytes appended.
Errors raise LzfseError carrying the status that mirrors lzfse_rust::Error (src/error/mod.rs:40-61).
New: decode_batch / encode_batch over many independent streams, on host buffers or on CUDA tensors.
"""
import ctypes as C

import numpy as np

from . import _ffi

OK = 0
STATUS_NAMES = {
    0: "Ok", 1: "BadBlock", 2: "BadBitStream", 3: "BadDValue", 4: "BadReaderState", 5: "BufferOverflow", 6: "PayloadOverflow",
    7: "PayloadUnderflow", 16: "Fse(BadLiteralBits)", 17: "Fse(BadLiteralCount)", 18: "Fse(BadLiteralPayload)",
    19: "Fse(BadLiteralState)", 20: "Fse(BadLmdBits)", 21: "Fse(BadLmdCount)", 22: "Fse(BadLmdPayload)", 23: "Fse(BadLmdState)",
    24: "Fse(BadPayloadCount)", 25: "Fse(BadRawByteCount)", 26: "Fse(BadReaderState)", 27: "Fse(BadWeightPayload)",
    28: "Fse(BadWeightPayloadCount)", 29: "Fse(WeightPayloadOverflow)", 30: "Fse(WeightPayloadUnderflow)",
    32: "Vn(BadPayloadCount)", 33: "Vn(BadPayload)", 34: "Vn(BadOpcode)", 64: "InvalidArgument", 65: "NoDevice", 66: "CudaError",
    67: "OutOfMemory",
}


class LzfseError(Exception):
    """lzfse_rust::Error.  `status` is the numeric code of include/lzfse_b200.h."""

    def __init__(self, status, detail=""):
        self.status = int(status)
        name = STATUS_NAMES.get(self.status, "Unknown(%d)" % self.status)
        super().__init__(name + (": " + detail if detail else ""))


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_u8(b):
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(b, dtype=np.uint8)


class _Handle:
    _kind = None

    def __init__(self, device=0):
        self._lib = _ffi.load()
        self._h = C.c_void_p()
        rc = getattr(self._lib, "lzfse_b200_%s_create" % self._kind)(int(device), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise LzfseError(rc, "cannot create %s on cuda:%d (no CPU fallback exists)" % (self._kind, device))
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            getattr(self._lib, "lzfse_b200_%s_destroy" % self._kind)(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __repr__(self):  # the reference types are Debug (decoder.rs:176, encoder.rs:56)
        return "Lzfse%s(cuda:%d)" % (self._kind.capitalize(), self.device)

    def _check(self, rc):
        if rc != 0:
            raise LzfseError(rc, getattr(self._lib, "lzfse_b200_%s_last_error" % self._kind)(self._h).decode())

    @property
    def last_launches(self):
        return int(getattr(self._lib, "lzfse_b200_%s_last_launches" % self._kind)(self._h))

    def set_timing(self, enabled=True):
        """Measurement aid: bracket every pipeline stage of *_batch_device calls with CUDA events."""
        getattr(self._lib, "lzfse_b200_%s_set_timing" % self._kind)(self._h, int(bool(enabled)))

    STAGES = ()

    def last_stage_ms(self):
        buf = (C.c_float * 16)()
        n = getattr(self._lib, "lzfse_b200_%s_last_stage_ms" % self._kind)(self._h, buf, 16)
        return dict(zip(self.STAGES, [float(buf[i]) for i in range(n)]))

    # ---- shared batch plumbing -------------------------------------------------------------
    def _batch_host(self, fn, src, src_off, src_len, dst, dst_off, dst_cap):
        n = len(src_off)
        src_off, src_len, dst_off, dst_cap = _u64(src_off), _u64(src_len), _u64(dst_off), _u64(dst_cap)
        out_len = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        self._check(fn(self._h, _ptr(src), _ptr(src_off), _ptr(src_len), _ptr(dst), _ptr(dst_off), _ptr(dst_cap), _ptr(out_len), _ptr(status), n))
        return out_len, status

    def sync(self):
        """Waits for the last `wait=False` batch call on this object (lzfse_b200_{decoder,encoder}_sync)."""
        self._check(getattr(self._lib, "lzfse_b200_%s_sync" % self._kind)(self._h))

    def _batch_device(self, fn, src, src_off, src_len, dst, dst_off, dst_cap, stream=None):
        import torch

        n = int(src_off.numel())
        for t in (src_off, src_len, dst_off, dst_cap):
            assert t.is_cuda and t.dtype in (torch.int64, torch.uint64) and t.is_contiguous() and t.numel() == n
        assert src.is_cuda and dst.is_cuda and src.dtype == torch.uint8 and dst.dtype == torch.uint8
        out_len = torch.empty(n, dtype=torch.int64, device=src.device)
        status = torch.empty(n, dtype=torch.int32, device=src.device)
        s = stream if stream is not None else torch.cuda.current_stream(src.device)
        self._check(fn(self._h, src.data_ptr(), src_off.data_ptr(), src_len.data_ptr(), dst.data_ptr(), dst_off.data_ptr(), dst_cap.data_ptr(),
                       out_len.data_ptr(), status.data_ptr(), n, s.cuda_stream))
        return out_len, status


class LzfseDecoder(_Handle):
    """LZFSE decoder (lzfse_rust::LzfseDecoder).  Reusable; one call at a time per object."""

    _kind = "decoder"
    STAGES = ("scan", "literals", "lmds", "expand", "finish")

    def decode_bytes(self, src, dst):
        """Decode the frame `src` and append it to the bytearray `dst`; returns the bytes appended."""
        src = _as_u8(src)
        raw, _, st = self.probe_batch(src, [0], [len(src)])
        cap = min(int(raw[0]), len(src) * self.MAX_RATIO + self.MAX_SLACK) if st[0] == 0 else 0
        out = np.empty(max(cap, 1), dtype=np.uint8)
        n = C.c_size_t(0)
        rc = self._lib.lzfse_b200_decode_bytes(self._h, _ptr(src), len(src), _ptr(out), cap, C.byref(n))
        if rc != 0:
            raise LzfseError(rc, self._lib.lzfse_b200_decoder_last_error(self._h).decode() if rc >= 64 else "")
        dst += out[: n.value].tobytes()
        return n.value

    def probe_batch(self, src, src_off, src_len):
        """Header walk: (raw_len[n], n_blocks[n], status[n]) for frames inside one host buffer."""
        src = _as_u8(src)
        n = len(src_off)
        src_off, src_len = _u64(src_off), _u64(src_len)
        raw, nb, st = np.zeros(n, np.uint64), np.zeros(n, np.uint32), np.zeros(n, np.int32)
        self._check(self._lib.lzfse_b200_decode_probe_batch_host(self._h, _ptr(src), _ptr(src_off), _ptr(src_len), _ptr(raw), _ptr(nb), _ptr(st), n))
        return raw, nb, st

    def decode_batch_into(self, src, src_off, src_len, dst, dst_off, dst_cap):
        """Batched decode on host buffers (numpy uint8 arrays).  Returns (out_len[n], status[n])."""
        return self._batch_host(self._lib.lzfse_b200_decode_batch_host, _as_u8(src), src_off, src_len, dst, dst_off, dst_cap)

    # The most an LZFSE frame can legitimately expand: 10 000 matches of 2 359 bytes from a block of a few hundred bytes.
    # decode_batch / decode_bytes size their buffers from the frames' own (untrusted) headers, so what a frame may
    # announce is capped at MAX_RATIO x its size + MAX_SLACK; a frame that really produces more fails with
    # BufferOverflow on its own, the rest of the batch is untouched.  Pass `caps` to lift the cap.
    MAX_RATIO, MAX_SLACK = 65536, 1 << 16

    def decode_batch(self, frames, caps=None):
        """Decode a list of frames; returns (list of bytes-or-None, status array)."""
        n = len(frames)
        lens = np.array([len(f) for f in frames], dtype=np.uint64)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
        src = np.frombuffer(b"".join(bytes(f) for f in frames) or b"\0", dtype=np.uint8)
        if caps is None:
            raw, _, _ = self.probe_batch(src, offs, lens)
            caps = np.minimum(raw, lens * np.uint64(self.MAX_RATIO) + np.uint64(self.MAX_SLACK))
        caps = _u64(caps)
        doff = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
        dst = np.empty(max(int(caps.sum()), 1), dtype=np.uint8)
        out_len, status = self.decode_batch_into(src, offs, lens, dst, doff, caps)
        outs = [dst[int(doff[i]) : int(doff[i]) + int(out_len[i])].tobytes() if status[i] == 0 else None for i in range(n)]
        return outs, status

    def decode_prefix_batch(self, frames, limits):
        """Bounded decode (the reference's decode_n as a batch call): the first limits[i] bytes of every frame.
        Returns (list of bytes-or-None, status array, more array); more[i] = 1 if the frame goes on after what was
        returned.  Errors beyond the blocks that had to be decoded are not seen."""
        n = len(frames)
        lens = np.array([len(f) for f in frames], dtype=np.uint64)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
        src = np.frombuffer(b"".join(bytes(f) for f in frames) or b"\0", dtype=np.uint8)
        limits = _u64(limits)
        doff = np.concatenate([[0], np.cumsum(limits)[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
        dst = np.empty(max(int(limits.sum()), 1), dtype=np.uint8)
        out_len, status, more = np.zeros(n, np.uint64), np.zeros(n, np.int32), np.zeros(n, np.uint8)
        self._check(self._lib.lzfse_b200_decode_prefix_batch_host(self._h, _ptr(src), _ptr(offs), _ptr(lens), _ptr(dst), _ptr(doff), _ptr(limits),
                                                                  _ptr(out_len), _ptr(status), _ptr(more), n))
        outs = [dst[int(doff[i]) : int(doff[i]) + int(out_len[i])].tobytes() if status[i] == 0 else None for i in range(n)]
        return outs, status, more

    def decode_batch_device(self, src, src_off, src_len, dst, dst_off, dst_cap, stream=None, wait=True):
        """Batched decode on CUDA tensors (uint8 data, int64 descriptors).  Returns (out_len, status) tensors.  The work
        is enqueued on `stream` (default: torch's current stream).  wait=False returns once everything is enqueued; the
        results are valid after sync() or for work ordered behind it on the same stream."""
        fn = self._lib.lzfse_b200_decode_batch_device if wait else self._lib.lzfse_b200_decode_batch_device_async
        return self._batch_device(fn, src, src_off, src_len, dst, dst_off, dst_cap, stream)


class LzfseEncoder(_Handle):
    """LZFSE encoder (lzfse_rust::LzfseEncoder).  Reusable; one call at a time per object."""

    _kind = "encoder"
    STAGES = ("prep", "find", "replay", "parse", "long_chain", "long_find", "long_replay", "long_stitch", "long_packs", "fse_blocks", "assemble")

    def encode_bound(self, n):
        return int(self._lib.lzfse_b200_encode_bound(int(n)))

    def encode_bytes(self, src, dst):
        """Encode `src` into one LZFSE frame appended to the bytearray `dst`; returns the bytes appended."""
        src = _as_u8(src) if len(src) else np.zeros(0, np.uint8)
        cap = self.encode_bound(len(src))
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        rc = self._lib.lzfse_b200_encode_bytes(self._h, _ptr(src) if len(src) else None, len(src), _ptr(out), cap, C.byref(n))
        if rc == 5:  # BufferOverflow under the working bound: once more with the format's worst case
            cap = int(self._lib.lzfse_b200_encode_bound_strict(len(src)))
            out = np.empty(cap, dtype=np.uint8)
            rc = self._lib.lzfse_b200_encode_bytes(self._h, _ptr(src) if len(src) else None, len(src), _ptr(out), cap, C.byref(n))
        if rc != 0:
            raise LzfseError(rc, self._lib.lzfse_b200_encoder_last_error(self._h).decode() if rc >= 64 else "")
        dst += out[: n.value].tobytes()
        return n.value

    def encode_batch_into(self, src, src_off, src_len, dst, dst_off, dst_cap):
        return self._batch_host(self._lib.lzfse_b200_encode_batch_host, _as_u8(src), src_off, src_len, dst, dst_off, dst_cap)

    def encode_batch(self, chunks):
        """Encode a list of byte strings into a list of frames; returns (frames, status array)."""
        n = len(chunks)
        lens = np.array([len(c) for c in chunks], dtype=np.uint64)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
        src = np.frombuffer(b"".join(bytes(c) for c in chunks) or b"\0", dtype=np.uint8)
        caps = np.array([self.encode_bound(int(l)) for l in lens], dtype=np.uint64)
        doff = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
        dst = np.empty(max(int(caps.sum()), 1), dtype=np.uint8)
        out_len, status = self.encode_batch_into(src, offs, lens, dst, doff, caps)
        outs = [dst[int(doff[i]) : int(doff[i]) + int(out_len[i])].tobytes() if status[i] == 0 else None for i in range(n)]
        over = [i for i in range(n) if status[i] == 5]  # BufferOverflow under the working bound: once more with the format's worst case
        if over:
            caps2 = np.array([int(self._lib.lzfse_b200_encode_bound_strict(int(lens[i]))) for i in over], dtype=np.uint64)
            doff2 = np.concatenate([[0], np.cumsum(caps2)[:-1]]).astype(np.uint64)
            dst2 = np.empty(int(caps2.sum()), dtype=np.uint8)
            ol2, st2 = self.encode_batch_into(src, offs[over], lens[over], dst2, doff2, caps2)
            for k, i in enumerate(over):
                status[i] = st2[k]
                outs[i] = dst2[int(doff2[k]) : int(doff2[k]) + int(ol2[k])].tobytes() if st2[k] == 0 else None
        return outs, status

    def encode_batch_device(self, src, src_off, src_len, dst, dst_off, dst_cap, stream=None, wait=True):
        fn = self._lib.lzfse_b200_encode_batch_device if wait else self._lib.lzfse_b200_encode_batch_device_async
        return self._batch_device(fn, src, src_off, src_len, dst, dst_off, dst_cap, stream)


def decode_bytes(src, dst, device=0):
    """lzfse_rust::decode_bytes (src/decode/mod.rs:49): temporary decoder, one frame."""
    d = LzfseDecoder(device)
    try:
        return d.decode_bytes(src, dst)
    finally:
        d.close()


def encode_bytes(src, dst, device=0):
    """lzfse_rust::encode_bytes (src/encode/mod.rs:58): temporary encoder, one frame."""
    e = LzfseEncoder(device)
    try:
        return e.encode_bytes(src, dst)
    finally:
        e.close()
