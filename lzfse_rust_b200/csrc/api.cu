// api.cu -- C-ABI entry points of liblzfse_b200.so (see include/lzfse_b200.h).
//
// Host side of the drop-in boundary: owns the CUDA device scratch (what `LzfseDecoder` / `LzfseEncoder`
// own as FseCore / HistoryTable, decode/decoder.rs:16-24, encode/encoder.rs:14-18) and sequences the
// kernels in decode.cu / encode.cu.  No CPU codec lives here: if CUDA is unavailable every call fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <string>

#include "common.cuh"
#include "host_util.h"

namespace lzb {
// decode.cu
uint32_t long_stream_threshold(size_t, int);
void launch_scan_count(const uint8_t *, const uint64_t *, const uint64_t *, const uint64_t *, size_t, StreamCounts *, uint32_t *, uint64_t *,
                       uint32_t *, StreamCounts *, uint32_t *, uint32_t, const uint64_t *, uint8_t *, cudaStream_t);
void launch_scan_fill(const uint8_t *, const uint64_t *, const uint64_t *, const uint64_t *, const uint64_t *, size_t, StreamCounts *, BlockDesc *, FseDesc *,
                      uint32_t *, uint64_t *, uint32_t *, uint32_t, uint64_t *, uint32_t *, uint32_t *, const uint64_t *, cudaStream_t);
void launch_scan_u64(const uint64_t *, uint64_t *, size_t, uint64_t *, cudaStream_t);
void launch_prefix_copy(const uint8_t *, const uint64_t *, const uint64_t *, const int32_t *, const uint64_t *, uint8_t *, const uint64_t *, uint64_t *,
                        uint8_t *, size_t, int, cudaStream_t);
int setup_decode_kernels();
void launch_fse_stages(const uint8_t *, const uint64_t *, const uint64_t *, const uint64_t *, const uint64_t *, const BlockDesc *, FseDesc *, uint32_t, uint8_t *,
                       LmdRec *, uint32_t *, uint32_t *, int, cudaStream_t, cudaEvent_t, cudaStream_t, cudaEvent_t, cudaEvent_t);
void launch_expand(const uint8_t *, const uint64_t *, const uint64_t *, uint8_t *, const uint64_t *, const uint64_t *, const StreamCounts *, const BlockDesc *,
                   const FseDesc *, const uint8_t *, const LmdRec *, uint32_t *, size_t, uint32_t *, const uint64_t *, int, cudaStream_t);
void launch_expand_vn(const uint8_t *, const uint64_t *, const uint64_t *, uint8_t *, const uint64_t *, const uint64_t *, const StreamCounts *,
                      const BlockDesc *, uint32_t *, size_t, uint32_t *, int, cudaStream_t);
void launch_finish(const uint32_t *, const uint64_t *, uint64_t *, int32_t *, size_t, cudaStream_t);
// expand.cu
int setup_expand_kernel();
void launch_expand_cta(const uint8_t *, const uint64_t *, const uint64_t *, uint8_t *, const uint64_t *, const uint64_t *, const StreamCounts *,
                       const BlockDesc *, const FseDesc *, const uint8_t *, const LmdRec *, const uint64_t *, uint32_t *, size_t, uint32_t *,
                       const uint64_t *, int, cudaStream_t);
// expand_long.cu
int launch_expand_long(const uint8_t *, const uint64_t *, const uint64_t *, uint8_t *, const uint64_t *, const uint64_t *, const StreamCounts *,
                        const BlockDesc *, const FseDesc *, const uint8_t *, const LmdRec *, const uint32_t *, const uint32_t *, const uint64_t *, uint32_t *,
                        uint32_t *, uint32_t, uint32_t, uint32_t *, int, cudaStream_t);
}  // namespace lzb

using namespace lzb;

// Device scratch of one kernel chain.  The host entry point runs several chains (one per slice of the batch) at the
// same time, each with its own scratch and stream; everything else uses chain 0.
struct DecodeScratch {
    DevBuf counts, err, raw_total, totals_dev, blocks, fse, lits, lmds, work;
    DevBuf long_base, long_blocks, long_streams, image;  // two-pass expansion of long streams (expand_long.cu)
    DevBuf inner, inner_off;                              // bounded decode: internal output buffer and its layout
    PinnedBuf totals_host;
    cudaStream_t stream = nullptr;  // chains 1.. only
    cudaStream_t side = nullptr;    // the literal stage of a small batch runs here, next to the LMD stage
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    void release() {
        if (side) cudaStreamDestroy(side);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        side = nullptr; ev_fork = ev_join = nullptr;
        for (DevBuf *b : {&counts, &err, &raw_total, &totals_dev, &blocks, &fse, &lits, &lmds, &work, &long_base, &long_blocks, &long_streams, &image, &inner, &inner_off}) b->release();
        totals_host.release();
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};
constexpr int kMaxChains = 6;

struct lzfse_b200_decoder {
    int device = 0;
    int n_sms = 148;
    cudaStream_t own_stream = nullptr;
    std::string last_error;
    uint64_t launches = 0;
    bool pending = false;             // an *_async call has been enqueued and not yet synchronised
    cudaStream_t pending_stream = nullptr;
    int expand_mode = 0;  // 0 = choose per batch, 1 = warp per stream, 2 = CTA per stream (LZB_EXPAND=warp|cta: measurements only)
    DecodeScratch chain[kMaxChains];
    // staging for the *_host entry points
    HostStage stage;
    StageTimer timer;
};

namespace {

#define CK LZB_CK

// Phase A of a chain: header scan (counts only).  Asynchronous.
int decode_launch_scan(lzfse_b200_decoder *d, DecodeScratch &c, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len,
                       const uint64_t *dst_cap, uint64_t *raw_len, uint32_t *n_blocks, size_t n, cudaStream_t s, bool probe_only,
                       const uint64_t *limit = nullptr, uint8_t *more = nullptr) {
    CK(d, c.counts.reserve((n + 1 + n / 1024 + 2) * sizeof(StreamCounts)));  // + tile sums of the scan
    CK(d, c.err.reserve(n * sizeof(uint32_t)));
    CK(d, c.raw_total.reserve(n * sizeof(uint64_t)));
    CK(d, c.totals_dev.reserve(sizeof(StreamCounts)));
    CK(d, c.totals_host.reserve(sizeof(StreamCounts) + sizeof(LongTotals) + sizeof(uint64_t)));
    CK(d, c.work.reserve(kWorkWords * sizeof(uint32_t)));
    CK(d, cudaMemsetAsync(c.work.p, 0, kWorkWords * sizeof(uint32_t), s));
    uint64_t *raw_total = raw_len ? raw_len : c.raw_total.as<uint64_t>();
    // Totals go straight into pinned host memory (UVA): no device-to-host copy that could queue behind a bulk
    // download on the copy engine.
    launch_scan_count(src, src_off, src_len, probe_only ? nullptr : dst_cap, n, c.counts.as<StreamCounts>(), c.err.as<uint32_t>(), raw_total,
                      n_blocks, c.totals_host.as<StreamCounts>(), c.work.as<uint32_t>(), long_stream_threshold(n, d->n_sms), limit, more, s);
    d->launches += n > 8192 ? 4 : 2;  // k_scan + the exclusive scan (three launches for large batches)
    return LZFSE_B200_OK;
}

// Phase B: waits for the scan (the one host round trip: scratch sizes depend on what the headers announce), then
// launches the rest of the chain.  Asynchronous after that; the caller synchronises `s`.
int decode_launch_rest(lzfse_b200_decoder *d, DecodeScratch &c, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                       const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n, cudaStream_t s,
                       StageTimer *timer, const uint64_t *limit = nullptr) {
    uint64_t *raw_total = c.raw_total.as<uint64_t>();
    CK(d, cudaStreamSynchronize(s));
    const StreamCounts tot = *c.totals_host.as<StreamCounts>();
    const LongTotals lt = *reinterpret_cast<const LongTotals *>(c.totals_host.as<StreamCounts>() + 1);
    if (tot.n_fse > 0xFFFFFFFFull) { d->last_error = "too many FSE blocks in one batch"; return LZFSE_B200_INVALID_ARGUMENT; }
    CK(d, c.blocks.reserve((tot.n_blocks + 1) * sizeof(BlockDesc)));
    CK(d, c.fse.reserve((tot.n_fse + 1) * sizeof(FseDesc)));
    CK(d, c.lits.reserve(tot.n_literals + 512));  // slack: the expander prefetches up to 128 bytes past a block's run
    CK(d, c.lmds.reserve((tot.n_lmds + 1) * sizeof(LmdRec)));
    CK(d, c.long_base.reserve(n * sizeof(uint64_t)));
    CK(d, c.long_blocks.reserve(((size_t)lt.n_blocks + 1) * sizeof(uint32_t)));
    CK(d, c.long_streams.reserve(((size_t)lt.n_streams + 1) * sizeof(uint32_t)));
    CK(d, c.image.reserve((lt.elements + 4) * sizeof(uint32_t)));  // four bytes per output byte of the long streams

    launch_scan_fill(src, src_off, src_len, dst_off, dst_cap, n, c.counts.as<StreamCounts>(), c.blocks.as<BlockDesc>(), c.fse.as<FseDesc>(),
                     c.err.as<uint32_t>(), raw_total, c.work.as<uint32_t>(), long_stream_threshold(n, d->n_sms), c.long_base.as<uint64_t>(),
                     c.long_blocks.as<uint32_t>(), c.long_streams.as<uint32_t>(), limit, s);
    d->launches += 1;
    if (timer) timer->mark(s);  // scan (count + host round trip + fill)
    if (tot.n_fse) {
        if (!c.side) {
            CK(d, cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking));
            CK(d, cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
            CK(d, cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
        }
        launch_fse_stages(src, src_off, src_len, dst_off, dst_cap, c.blocks.as<BlockDesc>(), c.fse.as<FseDesc>(), (uint32_t)tot.n_fse,
                          c.lits.as<uint8_t>(), c.lmds.as<LmdRec>(), c.err.as<uint32_t>(), c.work.as<uint32_t>(), d->n_sms, s,
                          timer && timer->enabled ? timer->ev[timer->n] : nullptr, c.side, c.ev_fork, c.ev_join);
        if (timer && timer->enabled) timer->n++;  // literals
        d->launches += 2;
    } else if (timer) {
        timer->mark(s);
    }
    if (timer) timer->mark(s);  // lmds
    // Expansion.  Streams that consist of one small LZVN block have their own kernel (the in-order kernels below skip
    // them); it is only launched when the batch has raw or LZVN blocks at all.
    if (tot.n_blocks > tot.n_fse) {
        launch_expand_vn(src, src_off, src_len, dst, dst_off, dst_cap, c.counts.as<StreamCounts>(), c.blocks.as<BlockDesc>(), c.err.as<uint32_t>(), n,
                         c.work.as<uint32_t>() + 3, d->n_sms, s);
        d->launches += 1;
    }
    // Everything else: a warp per stream when there are enough streams to fill the machine that way (32 warps x 148
    // SMs); otherwise a CTA per stream, whose 7 worker warps share one stream through a shared-memory window.
    // Long streams (many blocks): two passes, all their blocks side by side (expand_long.cu); the kernels below skip them.
    if (lt.n_streams) {
        CK(d, (cudaError_t)launch_expand_long(src, src_off, src_len, dst, dst_off, dst_cap, c.counts.as<StreamCounts>(), c.blocks.as<BlockDesc>(),
                                              c.fse.as<FseDesc>(), c.lits.as<uint8_t>(), c.lmds.as<LmdRec>(), c.long_blocks.as<uint32_t>(),
                                              c.long_streams.as<uint32_t>(), c.long_base.as<uint64_t>(), c.image.as<uint32_t>(), c.err.as<uint32_t>(),
                                              lt.n_streams, lt.n_blocks, c.work.as<uint32_t>(), d->n_sms, s));
        d->launches += 2;
    }
    const bool use_cta = d->expand_mode == 2 || (d->expand_mode == 0 && n < (size_t)d->n_sms * 16);
    if (lt.n_streams == n) {
        // nothing left for the in-order kernels
    } else if (!use_cta)
        launch_expand(src, src_off, src_len, dst, dst_off, dst_cap, c.counts.as<StreamCounts>(), c.blocks.as<BlockDesc>(), c.fse.as<FseDesc>(),
                      c.lits.as<uint8_t>(), c.lmds.as<LmdRec>(), c.err.as<uint32_t>(), n, c.work.as<uint32_t>() + 2, c.long_base.as<uint64_t>(), d->n_sms, s);
    else
        launch_expand_cta(src, src_off, src_len, dst, dst_off, dst_cap, c.counts.as<StreamCounts>(), c.blocks.as<BlockDesc>(),
                          c.fse.as<FseDesc>(), c.lits.as<uint8_t>(), c.lmds.as<LmdRec>(), raw_total, c.err.as<uint32_t>(), n, c.work.as<uint32_t>() + 2,
                          c.long_base.as<uint64_t>(), d->n_sms, s);
    if (timer) timer->mark(s);  // expand
    launch_finish(c.err.as<uint32_t>(), raw_total, out_len, status, n, s);
    if (timer) timer->mark(s);  // finish
    d->launches += lt.n_streams == n ? 1 : 2;
    CK(d, cudaGetLastError());
    return LZFSE_B200_OK;
}

int decode_batch_device_impl(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                             const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, uint64_t *raw_len,
                             uint32_t *n_blocks, size_t n, cudaStream_t s, bool probe_only, bool wait = true) {
    d->launches = 0;
    d->pending = false;
    if (n == 0) return LZFSE_B200_OK;
    DecodeScratch &c = d->chain[0];
    d->timer.begin(s);
    int rc = decode_launch_scan(d, c, src, src_off, src_len, dst_cap, raw_len, n_blocks, n, s, probe_only);
    if (rc) return rc;
    if (probe_only) {
        launch_finish(c.err.as<uint32_t>(), raw_len ? raw_len : c.raw_total.as<uint64_t>(), nullptr, status, n, s);
        d->launches += 1;
        CK(d, cudaStreamSynchronize(s));
        CK(d, cudaGetLastError());
        return LZFSE_B200_OK;
    }
    rc = decode_launch_rest(d, c, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status, n, s, &d->timer);
    if (rc) return rc;
    if (!wait) {  // lzfse_b200_decoder_sync (or the caller's own wait on `s`) completes the call
        d->pending = true;
        d->pending_stream = s;
        return LZFSE_B200_OK;
    }
    CK(d, cudaStreamSynchronize(s));
    d->timer.finish();
    return LZFSE_B200_OK;
}

// Bounded decode: the blocks that cover limit[i] bytes are decoded into an internal buffer (the block that crosses the
// limit is decoded whole, like the reference's decode_n overshoots by one LMD), then the prefix is copied out.
int decode_prefix_device_impl(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                              const uint64_t *dst_off, const uint64_t *limit, uint64_t *out_len, int32_t *status, uint8_t *more, size_t n,
                              cudaStream_t s) {
    d->launches = 0;
    d->pending = false;
    if (n == 0) return LZFSE_B200_OK;
    DecodeScratch &c = d->chain[0];
    CK(d, c.inner_off.reserve(n * sizeof(uint64_t)));
    int rc = decode_launch_scan(d, c, src, src_off, src_len, nullptr, nullptr, nullptr, n, s, false, limit, more);
    if (rc) return rc;
    uint64_t *inner_total = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(c.totals_host.p) + sizeof(StreamCounts) + sizeof(LongTotals));
    launch_scan_u64(c.raw_total.as<uint64_t>(), c.inner_off.as<uint64_t>(), n, inner_total, s);
    d->launches += 1;
    CK(d, cudaStreamSynchronize(s));
    CK(d, c.inner.reserve(*inner_total + 64));
    // the chain runs against the internal buffer: stream i owns raw_total[i] bytes at inner_off[i]
    rc = decode_launch_rest(d, c, src, src_off, src_len, c.inner.as<uint8_t>(), c.inner_off.as<uint64_t>(), c.raw_total.as<uint64_t>(), out_len, status, n, s,
                            nullptr, limit);
    if (rc) return rc;
    launch_prefix_copy(c.inner.as<uint8_t>(), c.inner_off.as<uint64_t>(), c.raw_total.as<uint64_t>(), status, limit, dst, dst_off, out_len, more, n, d->n_sms, s);
    d->launches += 1;
    CK(d, cudaGetLastError());
    CK(d, cudaStreamSynchronize(s));
    return LZFSE_B200_OK;
}

}  // namespace

extern "C" {

const char *lzfse_b200_version(void) { return "lzfse_b200 0.1.0 (sm_100a)"; }

const char *lzfse_b200_status_string(int st) {
    switch (st) {
    case LZFSE_B200_OK: return "ok";
    case LZFSE_B200_BAD_BLOCK: return "bad block";
    case LZFSE_B200_BAD_BITSTREAM: return "bad bitstream";
    case LZFSE_B200_BAD_D_VALUE: return "bad D value";
    case LZFSE_B200_BAD_READER_STATE: return "bad reader state";
    case LZFSE_B200_BUFFER_OVERFLOW: return "buffer overflow";
    case LZFSE_B200_PAYLOAD_OVERFLOW: return "bad payload overflow";
    case LZFSE_B200_PAYLOAD_UNDERFLOW: return "bad payload underflow";
    case LZFSE_B200_FSE_BAD_LITERAL_BITS: return "FSE: bad literal bits";
    case LZFSE_B200_FSE_BAD_LITERAL_COUNT: return "FSE: bad literal count";
    case LZFSE_B200_FSE_BAD_LITERAL_PAYLOAD: return "FSE: bad literal payload";
    case LZFSE_B200_FSE_BAD_LITERAL_STATE: return "FSE: bad literal state";
    case LZFSE_B200_FSE_BAD_LMD_BITS: return "FSE: bad LMD bits";
    case LZFSE_B200_FSE_BAD_LMD_COUNT: return "FSE: bad LMD count";
    case LZFSE_B200_FSE_BAD_LMD_PAYLOAD: return "FSE: bad LMD payload";
    case LZFSE_B200_FSE_BAD_LMD_STATE: return "FSE: bad LMD state";
    case LZFSE_B200_FSE_BAD_PAYLOAD_COUNT: return "FSE: bad payload count";
    case LZFSE_B200_FSE_BAD_RAW_BYTE_COUNT: return "FSE: bad raw byte count";
    case LZFSE_B200_FSE_BAD_READER_STATE: return "FSE: bad reader state";
    case LZFSE_B200_FSE_BAD_WEIGHT_PAYLOAD: return "FSE: bad weight payload";
    case LZFSE_B200_FSE_BAD_WEIGHT_PAYLOAD_COUNT: return "FSE: bad weight payload count";
    case LZFSE_B200_FSE_WEIGHT_PAYLOAD_OVERFLOW: return "FSE: weight payload overflow";
    case LZFSE_B200_FSE_WEIGHT_PAYLOAD_UNDERFLOW: return "FSE: weight payload underflow";
    case LZFSE_B200_VN_BAD_PAYLOAD_COUNT: return "VN: bad payload count";
    case LZFSE_B200_VN_BAD_PAYLOAD: return "VN: bad payload";
    case LZFSE_B200_VN_BAD_OPCODE: return "VN: bad opcode";
    case LZFSE_B200_INVALID_ARGUMENT: return "invalid argument";
    case LZFSE_B200_NO_DEVICE: return "no usable CUDA device";
    case LZFSE_B200_CUDA_ERROR: return "CUDA error";
    case LZFSE_B200_OUT_OF_MEMORY: return "out of device memory";
    default: return "unknown status";
    }
}

int lzfse_b200_decoder_create(int device, lzfse_b200_decoder **out) {
    if (!out) return LZFSE_B200_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return LZFSE_B200_NO_DEVICE;
    DeviceGuard g(device);
    if (!g.ok) return LZFSE_B200_NO_DEVICE;
    lzfse_b200_decoder *d = new (std::nothrow) lzfse_b200_decoder();
    if (!d) return LZFSE_B200_OUT_OF_MEMORY;
    d->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete d; return LZFSE_B200_NO_DEVICE; }
    d->n_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&d->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete d; return LZFSE_B200_CUDA_ERROR; }
    const char *xe = getenv("LZB_EXPAND");
    d->expand_mode = xe && strcmp(xe, "warp") == 0 ? 1 : (xe && strcmp(xe, "cta") == 0 ? 2 : 0);
    if (setup_decode_kernels() != 0 || setup_expand_kernel() != 0) {
        // no sm_100a image for this device, or not enough shared memory: there is no fallback path
        cudaStreamDestroy(d->own_stream);
        delete d;
        cudaGetLastError();
        return LZFSE_B200_NO_DEVICE;
    }
    *out = d;
    return LZFSE_B200_OK;
}

void lzfse_b200_decoder_destroy(lzfse_b200_decoder *d) {
    if (!d) return;
    DeviceGuard g(d->device);
    for (auto &c : d->chain) c.release();
    d->stage.release();
    d->timer.release();
    if (d->own_stream) cudaStreamDestroy(d->own_stream);
    delete d;
}

const char *lzfse_b200_decoder_last_error(const lzfse_b200_decoder *d) { return d ? d->last_error.c_str() : ""; }
void lzfse_b200_decoder_set_timing(lzfse_b200_decoder *d, int enabled) { if (d) d->timer.enabled = enabled != 0; }
int lzfse_b200_decoder_last_stage_ms(const lzfse_b200_decoder *d, float *ms, int cap) {
    if (!d) return 0;
    for (int i = 0; i < d->timer.n_done && i < cap; i++) ms[i] = d->timer.ms[i];
    return d->timer.n_done;
}
uint64_t lzfse_b200_decoder_last_launches(const lzfse_b200_decoder *d) { return d ? d->launches : 0; }
// Test hook (not part of the public header): the work counters of chain 0 after the last call (common.cuh, kWorkWords).
int lzfse_b200_debug_decoder_counters(lzfse_b200_decoder *d, uint32_t *out) {
    if (!d || !d->chain[0].work.p) return 0;
    DeviceGuard g(d->device);
    return cudaMemcpy(out, d->chain[0].work.p, kWorkWords * sizeof(uint32_t), cudaMemcpyDeviceToHost) == cudaSuccess ? (int)kWorkWords : 0;
}

int lzfse_b200_decode_batch_device(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                   const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n, void *stream) {
    if (!d || (n && (!src_off || !src_len || !dst_off || !dst_cap || !out_len || !status))) return LZFSE_B200_INVALID_ARGUMENT;
    DeviceGuard g(d->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    return decode_batch_device_impl(d, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status, nullptr, nullptr, n, (cudaStream_t)stream, false);
}

int lzfse_b200_decode_batch_device_async(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                         const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n, void *stream) {
    if (!d || (n && (!src_off || !src_len || !dst_off || !dst_cap || !out_len || !status))) return LZFSE_B200_INVALID_ARGUMENT;
    DeviceGuard g(d->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    return decode_batch_device_impl(d, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status, nullptr, nullptr, n, (cudaStream_t)stream, false, false);
}

int lzfse_b200_decoder_sync(lzfse_b200_decoder *d) {
    if (!d) return LZFSE_B200_INVALID_ARGUMENT;
    if (!d->pending) return LZFSE_B200_OK;
    DeviceGuard g(d->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    d->pending = false;
    CK(d, cudaStreamSynchronize(d->pending_stream));
    d->timer.finish();
    return LZFSE_B200_OK;
}

int lzfse_b200_decode_probe_batch_device(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len,
                                         uint64_t *raw_len, uint32_t *n_blocks, int32_t *status, size_t n, void *stream) {
    if (!d || (n && (!src_off || !src_len || !raw_len || !status))) return LZFSE_B200_INVALID_ARGUMENT;
    DeviceGuard g(d->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    return decode_batch_device_impl(d, src, src_off, src_len, nullptr, nullptr, nullptr, nullptr, status, raw_len, n_blocks, n, (cudaStream_t)stream, true);
}

// Host-buffer variants: stage the bytes on the device (host_util.h), run the device path, copy back.
//
// Large dense batches are cut into consecutive slices, each with its own kernel chain (stream + scratch), so that
// the upload of later slices, the kernels and the download of earlier slices overlap (PCIe is full duplex) -- and so
// that the chains themselves overlap: the entropy stages are bound by the serial latency of one block, not by the
// machine, hence a chain takes ~2.5 ms however small its slice is, and running the chains one after the other would
// leave the download engine idle between slices.  The download (the bound: ~56 GB/s) starts as soon as the first,
// small slice is done; by then the other chains have been running next to it.
static int decode_batch_host_pipelined_impl(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                            const uint64_t *dst_off, uint64_t *out_len, int32_t *status, size_t n);
static int decode_batch_host_pipelined(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                       const uint64_t *dst_off, uint64_t *out_len, int32_t *status, size_t n) {
    const int rc = decode_batch_host_pipelined_impl(d, src, src_off, src_len, dst, dst_off, out_len, status, n);
    if (rc) {  // a call-level failure: nothing of this call may still be running (or copying into the caller's buffers) when it returns
        for (auto &c : d->chain)
            if (c.stream) cudaStreamSynchronize(c.stream);
        if (d->stage.copy_in) cudaStreamSynchronize(d->stage.copy_in);
        if (d->stage.copy_out) cudaStreamSynchronize(d->stage.copy_out);
        cudaStreamSynchronize(d->own_stream);
        cudaGetLastError();
    }
    return rc;
}
static int decode_batch_host_pipelined_impl(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                            const uint64_t *dst_off, uint64_t *out_len, int32_t *status, size_t n) {
    uint32_t share_pct[kMaxChains] = {4, 12, 28, 52, 76, 100};  // cumulative share of the bytes crossing the bus
    int n_slices = kMaxChains;
    if (const char *ov = getenv("LZB_SLICES")) {  // tuning aid: comma-separated cumulative percentages
        int k = 0;
        for (const char *p = ov; *p && k < kMaxChains; k++) {
            share_pct[k] = (uint32_t)strtoul(p, const_cast<char **>(&p), 10);
            if (*p == ',') p++;
        }
        if (k > 0) { n_slices = k; share_pct[k - 1] = 100; }
    }
    HostStage &st = d->stage;
    if (!st.copy_in) {
        CK(d, cudaStreamCreateWithFlags(&st.copy_in, cudaStreamNonBlocking));
        CK(d, cudaStreamCreateWithFlags(&st.copy_out, cudaStreamNonBlocking));
        for (auto &e : st.ev_in) CK(d, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : st.ev_scan) CK(d, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(d, cudaEventCreateWithFlags(&st.ev_desc, cudaEventDisableTiming));
        int lo = 0, hi = 0;  // earlier slices get the higher priority: their output is needed first
        CK(d, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int k = 0; k < kMaxChains; k++) {
            const int prio = hi + k < lo ? hi + k : lo;
            CK(d, cudaStreamCreateWithPriority(&d->chain[k].stream, cudaStreamNonBlocking, prio));
        }
    }
    const uint64_t *pin = st.pin.as<uint64_t>();
    uint64_t *dd = st.desc.as<uint64_t>();
    // per-stream results are written by k_finish directly into pinned host memory (after the 4n descriptor words)
    uint64_t *d_out_len = st.pin.as<uint64_t>() + 4 * n;
    int32_t *d_status = reinterpret_cast<int32_t *>(d_out_len + n);
    size_t cut[kMaxChains + 1];
    {
        uint64_t total = 0, acc = 0;
        for (size_t i = 0; i < n; i++) total += src_len[i] + pin[3 * n + i];
        int k = 1;
        cut[0] = 0;
        for (size_t i = 0; i < n && k < n_slices; i++) {
            acc += src_len[i] + pin[3 * n + i];
            while (k < n_slices && acc >= total / 100 * share_pct[k - 1]) cut[k++] = i + 1;
        }
        for (; k < n_slices; k++) cut[k] = n;
        cut[n_slices] = n;
    }
    // the descriptor upload was enqueued on the caller-facing stream: every chain waits for it
    CK(d, cudaEventRecord(st.ev_desc, d->own_stream));
    // uploads, in slice order, on their own stream
    for (int k = 0; k < n_slices; k++) {
        const size_t i0 = cut[k], i1 = cut[k + 1];
        if (i1 > i0) {
            const uint64_t lo = src_off[i0], hi = src_off[i1 - 1] + src_len[i1 - 1];
            if (hi > lo) CK(d, cudaMemcpyAsync(st.src.as<uint8_t>() + (lo - st.src_lo), src + lo, hi - lo, cudaMemcpyHostToDevice, st.copy_in));
        }
        CK(d, cudaEventRecord(st.ev_in[k], st.copy_in));
    }
    d->launches = 0;
    // Phase A of every chain (asynchronous: each chain's stream waits for its own upload).
    for (int k = 0; k < n_slices; k++) {
        const size_t i0 = cut[k], m = cut[k + 1] - i0;
        if (m == 0) continue;
        DecodeScratch &c = d->chain[k];
        CK(d, cudaStreamWaitEvent(c.stream, st.ev_desc, 0));
        CK(d, cudaStreamWaitEvent(c.stream, st.ev_in[k], 0));
        int rc = decode_launch_scan(d, c, st.src.as<uint8_t>(), dd + i0, dd + n + i0, dd + 3 * n + i0, nullptr, nullptr, m, c.stream, false);
        if (rc) return rc;
        CK(d, cudaEventRecord(st.ev_scan[k], c.stream));
    }
    // Phase B of every chain in order, each as soon as its scan is through (i.e. its upload has arrived); while the
    // host waits for that it enqueues, in slice order, the downloads of the chains that have finished meanwhile.
    int next_dl = 0, launched = 0;
    auto service_downloads = [&](bool block) -> int {
        while (next_dl < launched) {
            const size_t i0 = cut[next_dl], i1 = cut[next_dl + 1], m = i1 - i0;
            if (m != 0) {
                if (block) CK(d, cudaStreamSynchronize(d->chain[next_dl].stream));
                else if (cudaStreamQuery(d->chain[next_dl].stream) != cudaSuccess) { cudaGetLastError(); return LZFSE_B200_OK; }
                memcpy(out_len + i0, d_out_len + i0, m * sizeof(uint64_t));
                memcpy(status + i0, d_status + i0, m * sizeof(int32_t));
                int rc = fetch_outputs(d, st, dst, dst_off, out_len, status, n, st.copy_out, i0, i1);
                if (rc) return rc;
            }
            next_dl++;
        }
        return LZFSE_B200_OK;
    };
    for (int k = 0; k < n_slices; k++) {
        const size_t i0 = cut[k], m = cut[k + 1] - i0;
        if (m != 0) {
            while (cudaEventQuery(st.ev_scan[k]) == cudaErrorNotReady) {
                int rc = service_downloads(false);
                if (rc) return rc;
            }
            DecodeScratch &c = d->chain[k];
            int rc = decode_launch_rest(d, c, st.src.as<uint8_t>(), dd + i0, dd + n + i0, st.dst.as<uint8_t>(), dd + 2 * n + i0, dd + 3 * n + i0,
                                        d_out_len + i0, d_status + i0, m, c.stream, nullptr);
            if (rc) return rc;
        }
        launched = k + 1;
    }
    {
        int rc = service_downloads(true);
        if (rc) return rc;
    }
    const uint64_t launches = d->launches;
    d->launches = launches;
    CK(d, cudaStreamSynchronize(st.copy_out));
    return LZFSE_B200_OK;
}

int lzfse_b200_decode_batch_host(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                 const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n) {
    if (!d || (n && (!src || !src_off || !src_len || !dst_off || !dst_cap || !out_len || !status))) return LZFSE_B200_INVALID_ARGUMENT;
    if (n == 0) return LZFSE_B200_OK;
    DeviceGuard g(d->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    cudaStream_t s = d->own_stream;
    HostStage &st = d->stage;
    // pipelining needs streams laid out in order on both sides (the usual packed batch) and enough bytes to matter
    uint64_t bytes = 0;
    bool ordered = n >= 64;
    for (size_t i = 0; i < n; i++) {
        bytes += src_len[i] + dst_cap[i];
        if (i && (src_off[i] < src_off[i - 1] + src_len[i - 1] || dst_off[i] < dst_off[i - 1] + dst_cap[i - 1])) ordered = false;
    }
    const bool want_pipeline = ordered && bytes >= (256ull << 20);
    // Stage timing is a device-API measurement aid; its timestamp events would serialise the copy/compute overlap here.
    struct TimerOff { StageTimer &t; bool was; explicit TimerOff(StageTimer &x) : t(x), was(x.enabled) { t.enabled = false; } ~TimerOff() { t.enabled = was; } } timer_off(d->timer);
    int rc = stage_sources(d, st, src, src_off, src_len, n, 6 * n + 8, s, want_pipeline);
    if (rc) return rc;
    rc = stage_outputs(d, st, dst_off, dst_cap, n);
    if (rc) return rc;
    CK(d, st.desc.reserve(4 * n * sizeof(uint64_t)));
    CK(d, st.res.reserve(n * (sizeof(uint64_t) + sizeof(int32_t))));
    CK(d, cudaMemcpyAsync(st.desc.p, st.pin.p, 4 * n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    if (want_pipeline && st.src_mirrored && st.dst_mirrored) return decode_batch_host_pipelined(d, src, src_off, src_len, dst, dst_off, out_len, status, n);
    if (want_pipeline && st.src_mirrored) {  // upload was deferred but the output side cannot be sliced: do it now
        uint64_t lo = ~0ull, hi = 0;
        for (size_t i = 0; i < n; i++) { if (src_off[i] < lo) lo = src_off[i]; if (src_off[i] + src_len[i] > hi) hi = src_off[i] + src_len[i]; }
        if (hi > lo) CK(d, cudaMemcpyAsync(st.src.p, src + lo, hi - lo, cudaMemcpyHostToDevice, s));
    }
    uint64_t *dd = st.desc.as<uint64_t>();
    uint64_t *d_out_len = st.res.as<uint64_t>();
    int32_t *d_status = reinterpret_cast<int32_t *>(d_out_len + n);
    rc = decode_batch_device_impl(d, st.src.as<uint8_t>(), dd, dd + n, st.dst.as<uint8_t>(), dd + 2 * n, dd + 3 * n, d_out_len, d_status, nullptr,
                                  nullptr, n, s, false);
    if (rc) return rc;
    CK(d, cudaMemcpyAsync(out_len, d_out_len, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CK(d, cudaMemcpyAsync(status, d_status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CK(d, cudaStreamSynchronize(s));
    rc = fetch_outputs(d, st, dst, dst_off, out_len, status, n, s);
    if (rc) return rc;
    CK(d, cudaStreamSynchronize(s));
    return LZFSE_B200_OK;
}

int lzfse_b200_decode_probe_batch_host(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint64_t *raw_len,
                                       uint32_t *n_blocks, int32_t *status, size_t n) {
    if (!d || (n && (!src || !src_off || !src_len || !raw_len || !status))) return LZFSE_B200_INVALID_ARGUMENT;
    if (n == 0) return LZFSE_B200_OK;
    DeviceGuard g(d->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    cudaStream_t s = d->own_stream;
    HostStage &st = d->stage;
    int rc = stage_sources(d, st, src, src_off, src_len, n, 4 * n, s);
    if (rc) return rc;
    CK(d, st.desc.reserve(4 * n * sizeof(uint64_t)));
    CK(d, cudaMemcpyAsync(st.desc.p, st.pin.p, 2 * n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CK(d, st.res.reserve(n * (sizeof(uint64_t) + sizeof(int32_t) + sizeof(uint32_t))));
    uint64_t *d_raw = st.res.as<uint64_t>();
    int32_t *d_status = reinterpret_cast<int32_t *>(d_raw + n);
    uint32_t *d_nb = reinterpret_cast<uint32_t *>(d_status + n);
    uint64_t *dd = st.desc.as<uint64_t>();
    rc = decode_batch_device_impl(d, st.src.as<uint8_t>(), dd, dd + n, nullptr, nullptr, nullptr, nullptr, d_status, d_raw, d_nb, n, s, true);
    if (rc) return rc;
    CK(d, cudaMemcpyAsync(raw_len, d_raw, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CK(d, cudaMemcpyAsync(status, d_status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (n_blocks) CK(d, cudaMemcpyAsync(n_blocks, d_nb, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(d, cudaStreamSynchronize(s));
    return LZFSE_B200_OK;
}

int lzfse_b200_decode_prefix_batch_device(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                          const uint64_t *dst_off, const uint64_t *limit, uint64_t *out_len, int32_t *status, uint8_t *more, size_t n,
                                          void *stream) {
    if (!d || (n && (!src_off || !src_len || !dst_off || !limit || !out_len || !status || !more))) return LZFSE_B200_INVALID_ARGUMENT;
    DeviceGuard g(d->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    return decode_prefix_device_impl(d, src, src_off, src_len, dst, dst_off, limit, out_len, status, more, n, (cudaStream_t)stream);
}

int lzfse_b200_decode_prefix_batch_host(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                        const uint64_t *dst_off, const uint64_t *limit, uint64_t *out_len, int32_t *status, uint8_t *more, size_t n) {
    if (!d || (n && (!src || !src_off || !src_len || !dst_off || !limit || !out_len || !status || !more))) return LZFSE_B200_INVALID_ARGUMENT;
    if (n == 0) return LZFSE_B200_OK;
    DeviceGuard g(d->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    cudaStream_t s = d->own_stream;
    HostStage &st = d->stage;
    int rc = stage_sources(d, st, src, src_off, src_len, n, 6 * n + 8, s);
    if (rc) return rc;
    rc = stage_outputs(d, st, dst_off, limit, n);  // the staged output regions are the limits
    if (rc) return rc;
    CK(d, st.desc.reserve(4 * n * sizeof(uint64_t)));
    CK(d, st.res.reserve(n * (sizeof(uint64_t) + sizeof(int32_t) + 1)));
    CK(d, cudaMemcpyAsync(st.desc.p, st.pin.p, 4 * n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    uint64_t *dd = st.desc.as<uint64_t>();
    uint64_t *d_out_len = st.res.as<uint64_t>();
    int32_t *d_status = reinterpret_cast<int32_t *>(d_out_len + n);
    uint8_t *d_more = reinterpret_cast<uint8_t *>(d_status + n);
    rc = decode_prefix_device_impl(d, st.src.as<uint8_t>(), dd, dd + n, st.dst.as<uint8_t>(), dd + 2 * n, dd + 3 * n, d_out_len, d_status, d_more, n, s);
    if (rc) return rc;
    CK(d, cudaMemcpyAsync(out_len, d_out_len, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CK(d, cudaMemcpyAsync(status, d_status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CK(d, cudaMemcpyAsync(more, d_more, n, cudaMemcpyDeviceToHost, s));
    CK(d, cudaStreamSynchronize(s));
    rc = fetch_outputs(d, st, dst, dst_off, out_len, status, n, s);
    if (rc) return rc;
    CK(d, cudaStreamSynchronize(s));
    return LZFSE_B200_OK;
}

int lzfse_b200_decode_bytes(lzfse_b200_decoder *d, const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *dst_len) {
    if (!d || (!src && src_len) || (!dst && dst_cap)) return LZFSE_B200_INVALID_ARGUMENT;
    uint64_t so = 0, sl = src_len, doff = 0, dc = dst_cap, ol = 0;
    int32_t st = 0;
    static const uint8_t empty = 0;
    int rc = lzfse_b200_decode_batch_host(d, src ? src : &empty, &so, &sl, dst, &doff, &dc, &ol, &st, 1);
    if (dst_len) *dst_len = (size_t)ol;
    return rc ? rc : st;
}

}  // extern "C"
