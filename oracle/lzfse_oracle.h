/*
 * lzfse_oracle.h -- CPU oracle for the LZFSE hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of shampoofactory/lzfse_rust v0.2.0's memory-buffer engine
 * (`LzfseDecoder::decode_bytes`, src/decode/decoder.rs:61 and `LzfseEncoder::encode_bytes`,
 * src/encode/encoder.rs:49).  It exists to check the CUDA path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it.
 * The shipped library (lzfse_rust_b200/csrc) never links, includes or falls back to this file.
 *
 * Parity pin: every golden vector the reference's own tests hold for this path
 * (data/{snappy,mutate,special} SHA-256 hashes, data/snappy/lmdy_output LMD dumps, the
 * encoder byte-exact KATs in src/encode/frontend_bytes.rs:455-531 and the doc-test frame in
 * src/encode/mod.rs:50-54) is checked in tests/test_oracle_*.py.
 * The Rust reference itself cannot be compiled here (no rustc/cargo), and Apple's C lzfse
 * (lzfse_sys) is not vendored, so there is no oracle/_ref build.
 */
#ifndef LZFSE_ORACLE_H
#define LZFSE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes.  Numerically identical to include/lzfse_b200.h (checked by tests). They mirror
 * lzfse_rust's `Error` (src/error/mod.rs:40-61), `FseErrorKind` (src/fse/error_kind.rs:9-39) and
 * `VnErrorKind` (src/vn/error_kind.rs:9-16). */
enum {
    ORC_OK = 0,
    ORC_BAD_BLOCK = 1,
    ORC_BAD_BITSTREAM = 2,
    ORC_BAD_D_VALUE = 3,
    ORC_BUFFER_OVERFLOW = 5,
    ORC_PAYLOAD_OVERFLOW = 6,
    ORC_PAYLOAD_UNDERFLOW = 7,
    ORC_FSE_BAD_LITERAL_BITS = 16,
    ORC_FSE_BAD_LITERAL_COUNT = 17,
    ORC_FSE_BAD_LITERAL_PAYLOAD = 18,
    ORC_FSE_BAD_LITERAL_STATE = 19,
    ORC_FSE_BAD_LMD_BITS = 20,
    ORC_FSE_BAD_LMD_COUNT = 21,
    ORC_FSE_BAD_LMD_PAYLOAD = 22,
    ORC_FSE_BAD_LMD_STATE = 23,
    ORC_FSE_BAD_PAYLOAD_COUNT = 24,
    ORC_FSE_BAD_RAW_BYTE_COUNT = 25,
    ORC_FSE_BAD_WEIGHT_PAYLOAD = 27,
    ORC_FSE_WEIGHT_PAYLOAD_OVERFLOW = 29,
    ORC_FSE_WEIGHT_PAYLOAD_UNDERFLOW = 30,
    ORC_VN_BAD_PAYLOAD_COUNT = 32,
    ORC_VN_BAD_PAYLOAD = 33,
    ORC_VN_BAD_OPCODE = 34
};

/* One decoded LZ77 step, as the reference's decoder hands it to its LzWriter. */
typedef struct {
    uint32_t literal_len;
    uint32_t match_len;
    uint32_t match_distance; /* already substituted (D==0 => previous), 0 when match_len==0 */
} orc_lmd_t;

/* decode_bytes.  Appends nothing: writes the decoded frame to dst[0..cap) and its size to *out_len.
 * If trace != NULL, every (L,M,D) step is appended to it (up to trace_cap entries); *trace_len gets
 * the number of steps the decoder executed. */
int orc_decode(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *out_len);
int orc_decode_trace(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *out_len,
                     orc_lmd_t *trace, size_t trace_cap, size_t *trace_len);

/* Header-only walk: total raw bytes and block count of a frame (what the GPU pre-pass computes). */
int orc_probe(const uint8_t *src, size_t src_len, size_t *raw_len, size_t *n_blocks);

/* encode_bytes.  Opaque encoder object owns the 512 KiB history table and the FSE block buffers
 * exactly like `LzfseEncoder` (src/encode/encoder.rs:14-18). */
typedef struct orc_encoder orc_encoder;
orc_encoder *orc_encoder_create(void);
void orc_encoder_destroy(orc_encoder *e);
size_t orc_encode_bound(size_t src_len);
int orc_encode(orc_encoder *e, const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap,
               size_t *out_len);

/* FSE stage only ("given LMD stream" harness, src/test_utils/lmds.rs:36-61): push an explicit
 * LMD list + literal bytes through the FSE backend (src/fse/backend.rs) and emit the bvx2
 * block(s), no EOS.  lmds[i].match_len==0 => push_literals. */
int orc_fse_encode_lmds(orc_encoder *e, const uint8_t *literals, size_t n_literals, const orc_lmd_t *lmds,
                        size_t n_lmds, uint8_t *dst, size_t dst_cap, size_t *out_len);
/* Same for the LZVN backend (src/vn/backend.rs). */
int orc_vn_encode_lmds(const uint8_t *literals, size_t n_literals, const orc_lmd_t *lmds, size_t n_lmds,
                       uint8_t *dst, size_t dst_cap, size_t *out_len);

/* Front-end only: run the match finder of `FrontendBytes` with the FSE (vn=0) or VN (vn=1)
 * parameters and return the LMD list it pushes to its backend (before any splitting). */
int orc_frontend_lmds(orc_encoder *e, const uint8_t *src, size_t src_len, int vn, orc_lmd_t *lmds,
                      size_t lmd_cap, size_t *n_lmds);

/* LZVN opcode class of a first byte (vn/constants.rs:39-72): 0 SmlL 1 LrgL 2 SmlM 3 LrgM 4 PreD 5 SmlD
 * 6 MedD 7 LrgD 8 Eos 9 Udef 10 Nop. */
int orc_vn_op_class(uint8_t b);

/* Weight normalisation (src/fse/weights.rs:218-278) exposed for unit tests. */
void orc_normalize_m1(uint16_t *weights, size_t n, uint32_t in_total, uint32_t out_total);

/* Batched helpers for the CPU baseline: n independent streams, `n_threads` worker threads pulling
 * stream indices from an atomic counter, one encoder/decoder object per thread. */
int orc_decode_batch(const uint8_t *src_base, const uint64_t *src_off, const uint64_t *src_len,
                     uint8_t *dst_base, const uint64_t *dst_off, const uint64_t *dst_cap,
                     uint64_t *out_len, int32_t *status, size_t n, int n_threads);
int orc_encode_batch(const uint8_t *src_base, const uint64_t *src_off, const uint64_t *src_len,
                     uint8_t *dst_base, const uint64_t *dst_off, const uint64_t *dst_cap,
                     uint64_t *out_len, int32_t *status, size_t n, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
