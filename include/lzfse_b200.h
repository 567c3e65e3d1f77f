/*
 * lzfse_b200.h -- C ABI of the B200-native batched LZFSE codec (liblzfse_b200.so).
 *
 * Drop-in boundary for lzfse_rust's memory-buffer engine.  Each entry point names the reference
 * interface it replaces (paths relative to the lzfse_rust v0.2.0 source tree):
 *
 *   lzfse_b200_decoder_create/destroy  <->  LzfseDecoder::default() / Drop   src/decode/decoder.rs:16-24
 *   lzfse_b200_decode_bytes            <->  LzfseDecoder::decode_bytes       src/decode/decoder.rs:61
 *                                           (and the free fn decode_bytes    src/decode/mod.rs:49)
 *   lzfse_b200_encoder_create/destroy  <->  LzfseEncoder::default() / Drop   src/encode/encoder.rs:13-18
 *   lzfse_b200_encode_bytes            <->  LzfseEncoder::encode_bytes       src/encode/encoder.rs:49
 *                                           (and the free fn encode_bytes    src/encode/mod.rs:58)
 *   lzfse_b200_{decode,encode}_batch_* <->  NEW: the same call over n independent LZFSE streams
 *   lzfse_b200_decode_probe_batch_*    <->  decode::probe                    src/decode/probe.rs:11-35
 *   lzfse_b200_decode_prefix_batch_*   <->  FseCore::decode_n / VnCore::decode_n / RawBlock::decode_n (bounded decode)
 *                                           src/fse/fse_core.rs:143-198, src/vn/vn_core.rs:67-73, src/raw/block.rs:59-68
 *   status codes                       <->  lzfse_rust::Error                src/error/mod.rs:40-61,
 *                                           FseErrorKind src/fse/error_kind.rs:9-39,
 *                                           VnErrorKind  src/vn/error_kind.rs:9-16
 *
 * Differences from the Rust signatures, all forced by the C boundary:
 *   - `dst: &mut Vec<u8>` (append, grows) becomes `dst, dst_cap, *dst_len`: the frame is written at
 *     dst[0..*dst_len).  A Rust wrapper does `dst.reserve(bound)`, passes the spare capacity and
 *     `set_len`s; see INTEGRATION.md.  Too small a capacity yields LZFSE_B200_BUFFER_OVERFLOW.
 *   - Match distances may not reach before dst[0] (the reference lets them reach into bytes that were
 *     already in the Vec, src/lz/writer.rs:156-157).
 *
 * All work runs on the GPU.  There is no CPU fallback: without a usable CUDA device every call
 * returns LZFSE_B200_NO_DEVICE (create) or LZFSE_B200_CUDA_ERROR.
 *
 * Threading: a handle is single-caller (`&mut self` in the reference).  Distinct handles may be
 * used concurrently.  *_batch_device calls enqueue on the given CUDA stream, synchronise it once
 * internally after the header scan (scratch sizes depend on what the headers announce) and once at
 * the end; results are complete when they return.  The *_batch_device_async variants skip the
 * second wait: they return as soon as every kernel is enqueued (the header scan, a few tens of
 * microseconds of device time, is the only thing they wait for); out_len / status / dst are valid
 * once `cuda_stream` reaches that point -- lzfse_b200_{decoder,encoder}_sync waits for it, or the
 * caller orders its own work on the same stream.  One asynchronous call may be in flight per handle
 * (its scratch is the handle's); use one handle per stream to overlap batches.
 *
 * `cuda_stream` is a cudaStream_t.  NULL is CUDA's legacy default stream (stream 0), with its usual
 * ordering against other blocking streams -- e.g. what torch.cuda.current_stream() is by default.
 */
#ifndef LZFSE_B200_H
#define LZFSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-stream status == lzfse_rust::Error discriminants (0 = Ok). */
enum lzfse_b200_status {
    LZFSE_B200_OK = 0,
    LZFSE_B200_BAD_BLOCK = 1,             /* Error::BadBlock */
    LZFSE_B200_BAD_BITSTREAM = 2,         /* Error::BadBitStream */
    LZFSE_B200_BAD_D_VALUE = 3,           /* Error::BadDValue */
    LZFSE_B200_BAD_READER_STATE = 4,      /* Error::BadReaderState (never produced here) */
    LZFSE_B200_BUFFER_OVERFLOW = 5,       /* Error::BufferOverflow: dst_cap too small */
    LZFSE_B200_PAYLOAD_OVERFLOW = 6,      /* Error::PayloadOverflow */
    LZFSE_B200_PAYLOAD_UNDERFLOW = 7,     /* Error::PayloadUnderflow */
    LZFSE_B200_FSE_BAD_LITERAL_BITS = 16, /* Error::Fse(FseErrorKind::...) = 16 + discriminant */
    LZFSE_B200_FSE_BAD_LITERAL_COUNT = 17,
    LZFSE_B200_FSE_BAD_LITERAL_PAYLOAD = 18,
    LZFSE_B200_FSE_BAD_LITERAL_STATE = 19,
    LZFSE_B200_FSE_BAD_LMD_BITS = 20,
    LZFSE_B200_FSE_BAD_LMD_COUNT = 21,
    LZFSE_B200_FSE_BAD_LMD_PAYLOAD = 22,
    LZFSE_B200_FSE_BAD_LMD_STATE = 23,
    LZFSE_B200_FSE_BAD_PAYLOAD_COUNT = 24,
    LZFSE_B200_FSE_BAD_RAW_BYTE_COUNT = 25,
    LZFSE_B200_FSE_BAD_READER_STATE = 26,
    LZFSE_B200_FSE_BAD_WEIGHT_PAYLOAD = 27,
    LZFSE_B200_FSE_BAD_WEIGHT_PAYLOAD_COUNT = 28,
    LZFSE_B200_FSE_WEIGHT_PAYLOAD_OVERFLOW = 29,
    LZFSE_B200_FSE_WEIGHT_PAYLOAD_UNDERFLOW = 30,
    LZFSE_B200_VN_BAD_PAYLOAD_COUNT = 32, /* Error::Vn(VnErrorKind::...) = 32 + discriminant */
    LZFSE_B200_VN_BAD_PAYLOAD = 33,
    LZFSE_B200_VN_BAD_OPCODE = 34,
    /* call-level failures (return values; per-stream only where a function says so) */
    LZFSE_B200_INVALID_ARGUMENT = 64,
    LZFSE_B200_NO_DEVICE = 65,
    LZFSE_B200_CUDA_ERROR = 66,
    LZFSE_B200_OUT_OF_MEMORY = 67         /* io::ErrorKind::Other in the reference's encode path */
};

typedef struct lzfse_b200_decoder lzfse_b200_decoder;
typedef struct lzfse_b200_encoder lzfse_b200_encoder;

const char *lzfse_b200_version(void);
const char *lzfse_b200_status_string(int status);
/* Last CUDA error text seen by this handle's calls ("" if none). */
const char *lzfse_b200_decoder_last_error(const lzfse_b200_decoder *d);
const char *lzfse_b200_encoder_last_error(const lzfse_b200_encoder *e);

/* ---- decoder ---------------------------------------------------------------------------- */
int lzfse_b200_decoder_create(int cuda_device, lzfse_b200_decoder **out);
void lzfse_b200_decoder_destroy(lzfse_b200_decoder *d);

/* decode_bytes: one frame, host buffers.  Returns the stream's status. */
int lzfse_b200_decode_bytes(lzfse_b200_decoder *d, const uint8_t *src, size_t src_len, uint8_t *dst,
                            size_t dst_cap, size_t *dst_len);

/* Batched decode of n independent frames.
 * Stream i reads  src_base[src_off[i] .. src_off[i]+src_len[i])  and writes its output at
 * dst_base[dst_off[i] ..), at most dst_cap[i] bytes; out_len[i] and status[i] receive the result.
 * Output regions of different streams must not overlap.  A failing stream does not disturb others;
 * its output region's contents are unspecified (as in the reference, src/lz/writer.rs:18,28).
 * _device: every pointer is device memory on the handle's GPU; `cuda_stream` is a cudaStream_t (NULL = stream 0).
 * _host:   every pointer is host memory; H2D/D2H copies happen inside the call.
 * Return value: call-level status (LZFSE_B200_OK even if individual streams failed).
 * Limit (no reference counterpart): a frame that decodes to more than 0xF0000000 bytes fails with
 * LZFSE_B200_BUFFER_OVERFLOW at the block that crosses the limit (32-bit stream positions on the device). */
int lzfse_b200_decode_batch_device(lzfse_b200_decoder *d, const uint8_t *src_base, const uint64_t *src_off,
                                   const uint64_t *src_len, uint8_t *dst_base, const uint64_t *dst_off,
                                   const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n,
                                   void *cuda_stream);
int lzfse_b200_decode_batch_host(lzfse_b200_decoder *d, const uint8_t *src_base, const uint64_t *src_off,
                                 const uint64_t *src_len, uint8_t *dst_base, const uint64_t *dst_off,
                                 const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n);
/* Asynchronous form of _device (see "Threading" above) and the wait that completes it. */
int lzfse_b200_decode_batch_device_async(lzfse_b200_decoder *d, const uint8_t *src_base, const uint64_t *src_off,
                                         const uint64_t *src_len, uint8_t *dst_base, const uint64_t *dst_off,
                                         const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n,
                                         void *cuda_stream);
int lzfse_b200_decoder_sync(lzfse_b200_decoder *d);

/* Bounded decode: the first limit[i] bytes of every stream -- the batched counterpart of the reference's incremental
 * decode_n (src/fse/fse_core.rs:143-198, src/vn/vn_core.rs:67-73, src/raw/block.rs:59-68), which its streaming reader
 * uses to produce a frame piecewise.  Blocks are decoded (whole) until they cover limit[i] bytes; out_len[i] =
 * min(limit[i], bytes those blocks produce) bytes are written at dst_base[dst_off[i] ..); more[i] = 1 if the frame goes
 * on after what was returned (output was cut off, or blocks that produce bytes follow), 0 if it ended there (blocks
 * that produce nothing and the end-of-stream marker are then still checked as in a full decode).
 * status[i] covers the decoded blocks only: damage further on is not seen, as with decode_n.  The internal buffer is
 * sized by what the decoded blocks announce, not by the whole frame, so this is also the safe way to look into frames
 * of unknown provenance.  `more` is n bytes. */
int lzfse_b200_decode_prefix_batch_device(lzfse_b200_decoder *d, const uint8_t *src_base, const uint64_t *src_off,
                                          const uint64_t *src_len, uint8_t *dst_base, const uint64_t *dst_off,
                                          const uint64_t *limit, uint64_t *out_len, int32_t *status, uint8_t *more,
                                          size_t n, void *cuda_stream);
int lzfse_b200_decode_prefix_batch_host(lzfse_b200_decoder *d, const uint8_t *src_base, const uint64_t *src_off,
                                        const uint64_t *src_len, uint8_t *dst_base, const uint64_t *dst_off,
                                        const uint64_t *limit, uint64_t *out_len, int32_t *status, uint8_t *more,
                                        size_t n);

/* Header-only walk: raw_len[i] = sum of the blocks' n_raw_bytes, n_blocks[i] = block count.
 * (Unlike the reference's dead-code probe, LZVN blocks advance by n_payload_bytes; cf. src/vn/ops.rs:13.) */
int lzfse_b200_decode_probe_batch_device(lzfse_b200_decoder *d, const uint8_t *src_base, const uint64_t *src_off,
                                         const uint64_t *src_len, uint64_t *raw_len, uint32_t *n_blocks,
                                         int32_t *status, size_t n, void *cuda_stream);
int lzfse_b200_decode_probe_batch_host(lzfse_b200_decoder *d, const uint8_t *src_base, const uint64_t *src_off,
                                       const uint64_t *src_len, uint64_t *raw_len, uint32_t *n_blocks,
                                       int32_t *status, size_t n);

/* Kernel launches issued by the last batch call on this handle (bench.py's gpu_launches). */
uint64_t lzfse_b200_decoder_last_launches(const lzfse_b200_decoder *d);

/* Measurement aid (no reference counterpart): when enabled, the next *_batch_device calls bracket every
 * pipeline stage with CUDA events on the launching stream.  stage_ms receives up to `cap` durations in
 * milliseconds in pipeline order; the return value is the number of stages
 * (decoder: scan, literals, lmds, expand, finish; encoder: prep, parse, fse_blocks, assemble). */
void lzfse_b200_decoder_set_timing(lzfse_b200_decoder *d, int enabled);
int lzfse_b200_decoder_last_stage_ms(const lzfse_b200_decoder *d, float *stage_ms, int cap);

/* ---- encoder ---------------------------------------------------------------------------- */
int lzfse_b200_encoder_create(int cuda_device, lzfse_b200_encoder **out);
void lzfse_b200_encoder_destroy(lzfse_b200_encoder *e);

/* Frame-size bounds for src_len input bytes.
 * lzfse_b200_encode_bound is the working bound: 1.25 x the input plus 768 bytes per 16 KiB, which no input we know of
 * exceeds (incompressible data costs ~8.2 bits per byte plus one block header per 40 000 literals) but which is not
 * derived from the format's worst case; a frame that did exceed it fails with LZFSE_B200_BUFFER_OVERFLOW, where the
 * reference's Vec would have grown.  lzfse_b200_encode_bound_strict is that worst case (10 bits per literal, 54 bits per
 * L/M/D triple, a header and weight table per block: ~3 x the input); a caller that must never see BufferOverflow retries
 * the failing streams with it (the Python wrappers do).
 * Inputs of more than 0x7FFFFFFF bytes per stream are not supported (the reference repositions its history there,
 * src/encode/frontend_bytes.rs:348-375): such a stream gets the per-stream status LZFSE_B200_INVALID_ARGUMENT -- the one
 * case where a code >= 64 appears in status[]. */
size_t lzfse_b200_encode_bound(size_t src_len);
size_t lzfse_b200_encode_bound_strict(size_t src_len);

/* encode_bytes: one frame, host buffers. */
int lzfse_b200_encode_bytes(lzfse_b200_encoder *e, const uint8_t *src, size_t src_len, uint8_t *dst,
                            size_t dst_cap, size_t *dst_len);

/* Batched encode: n independent inputs -> n independent frames (same layout rules as decode).
 * Device memory: the handle's scratch grows to about 19 bytes per input byte of the largest batch it has seen (a 32-bit word per
 * position, the per-segment match lists of the front end, packs, literals and the blocks' output scratch; 27 bytes per input
 * byte for streams longer than 64 KiB, whose hash chain lives in HBM) and is kept until the handle is destroyed; a batch that does
 * not fit returns LZFSE_B200_OUT_OF_MEMORY as the call's status -- split it.  A stream may be as long as 0x7FFFFFFF bytes. */
int lzfse_b200_encode_batch_device(lzfse_b200_encoder *e, const uint8_t *src_base, const uint64_t *src_off,
                                   const uint64_t *src_len, uint8_t *dst_base, const uint64_t *dst_off,
                                   const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n,
                                   void *cuda_stream);
int lzfse_b200_encode_batch_host(lzfse_b200_encoder *e, const uint8_t *src_base, const uint64_t *src_off,
                                 const uint64_t *src_len, uint8_t *dst_base, const uint64_t *dst_off,
                                 const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n);
int lzfse_b200_encode_batch_device_async(lzfse_b200_encoder *e, const uint8_t *src_base, const uint64_t *src_off,
                                         const uint64_t *src_len, uint8_t *dst_base, const uint64_t *dst_off,
                                         const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n,
                                         void *cuda_stream);
int lzfse_b200_encoder_sync(lzfse_b200_encoder *e);
uint64_t lzfse_b200_encoder_last_launches(const lzfse_b200_encoder *e);
void lzfse_b200_encoder_set_timing(lzfse_b200_encoder *e, int enabled);
int lzfse_b200_encoder_last_stage_ms(const lzfse_b200_encoder *e, float *stage_ms, int cap);

#ifdef __cplusplus
}
#endif
#endif
