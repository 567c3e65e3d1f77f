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
void launch_scan_count(const uint8_t *, const uint64_t *, const uint64_t *, const uint64_t *, size_t, StreamCounts *, uint32_t *, uint64_t *,
                       uint32_t *, StreamCounts *, cudaStream_t);
void launch_scan_fill(const uint8_t *, const uint64_t *, const uint64_t *, const uint64_t *, const uint64_t *, size_t, StreamCounts *, BlockDesc *, FseDesc *,
                      uint32_t *, cudaStream_t);
int setup_decode_kernels();
void launch_fse_stages(const uint8_t *, const uint64_t *, const uint64_t *, const uint64_t *, const uint64_t *, const BlockDesc *, FseDesc *, uint32_t, uint8_t *,
                       LmdRec *, uint32_t *, uint32_t *, int, cudaStream_t, cudaEvent_t);
void launch_expand(const uint8_t *, const uint64_t *, const uint64_t *, uint8_t *, const uint64_t *, const uint64_t *, const StreamCounts *, const BlockDesc *,
                   const FseDesc *, const uint8_t *, const LmdRec *, uint32_t *, size_t, cudaStream_t);
void launch_finish(const uint32_t *, const uint64_t *, uint64_t *, int32_t *, size_t, cudaStream_t);
// expand.cu
int setup_expand_kernel();
void launch_expand_cta(const uint8_t *, const uint64_t *, const uint64_t *, uint8_t *, const uint64_t *, const uint64_t *, const StreamCounts *,
                       const BlockDesc *, const FseDesc *, const uint8_t *, const LmdRec *, const uint64_t *, uint32_t *, size_t, uint32_t *, int,
                       cudaStream_t);
}  // namespace lzb

using namespace lzb;

struct lzfse_b200_decoder {
    int device = 0;
    int n_sms = 148;
    cudaStream_t own_stream = nullptr;
    std::string last_error;
    uint64_t launches = 0;
    int expand_mode = 0;  // 0 = choose per batch, 1 = warp per stream, 2 = CTA per stream (LZB_EXPAND=warp|cta: measurements only)
    // scratch
    DevBuf counts, err, raw_total, totals_dev, blocks, fse, lits, lmds, work;
    PinnedBuf totals_host;
    // staging for the *_host entry points
    HostStage stage;
    StageTimer timer;
};

namespace {

#define CK LZB_CK

int decode_batch_device_impl(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                             const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, uint64_t *raw_len,
                             uint32_t *n_blocks, size_t n, cudaStream_t s, bool probe_only) {
    d->launches = 0;
    if (n == 0) return LZFSE_B200_OK;
    CK(d, d->counts.reserve((n + 1) * sizeof(StreamCounts)));
    CK(d, d->err.reserve(n * sizeof(uint32_t)));
    CK(d, d->raw_total.reserve(n * sizeof(uint64_t)));
    CK(d, d->totals_dev.reserve(sizeof(StreamCounts)));
    CK(d, d->totals_host.reserve(sizeof(StreamCounts)));
    CK(d, d->work.reserve(4 * sizeof(uint32_t)));
    uint64_t *raw_total = raw_len ? raw_len : d->raw_total.as<uint64_t>();
    d->timer.begin(s);

    // Totals go straight into pinned host memory (UVA): no device-to-host copy that could queue behind a bulk
    // download on the copy engine.
    launch_scan_count(src, src_off, src_len, probe_only ? nullptr : dst_cap, n, d->counts.as<StreamCounts>(), d->err.as<uint32_t>(), raw_total,
                      n_blocks, d->totals_host.as<StreamCounts>(), s);
    d->launches += 2;
    if (probe_only) {
        launch_finish(d->err.as<uint32_t>(), raw_total, nullptr, status, n, s);
        d->launches += 1;
        CK(d, cudaStreamSynchronize(s));
        CK(d, cudaGetLastError());
        return LZFSE_B200_OK;
    }
    // The one host round trip: scratch sizes depend on what the headers announce.
    CK(d, cudaStreamSynchronize(s));
    const StreamCounts tot = *d->totals_host.as<StreamCounts>();
    if (tot.n_fse > 0xFFFFFFFFull) { d->last_error = "too many FSE blocks in one batch"; return LZFSE_B200_INVALID_ARGUMENT; }
    CK(d, d->blocks.reserve((tot.n_blocks + 1) * sizeof(BlockDesc)));
    CK(d, d->fse.reserve((tot.n_fse + 1) * sizeof(FseDesc)));
    CK(d, d->lits.reserve(tot.n_literals + 512));  // slack: the expander prefetches up to 128 bytes past a block's run
    CK(d, d->lmds.reserve((tot.n_lmds + 1) * sizeof(LmdRec)));
    CK(d, cudaMemsetAsync(d->work.p, 0, 4 * sizeof(uint32_t), s));

    launch_scan_fill(src, src_off, src_len, dst_off, dst_cap, n, d->counts.as<StreamCounts>(), d->blocks.as<BlockDesc>(), d->fse.as<FseDesc>(),
                     d->err.as<uint32_t>(), s);
    d->launches += 1;
    d->timer.mark(s);  // scan (count + host round trip + fill)
    if (tot.n_fse) {
        launch_fse_stages(src, src_off, src_len, dst_off, dst_cap, d->blocks.as<BlockDesc>(), d->fse.as<FseDesc>(), (uint32_t)tot.n_fse,
                          d->lits.as<uint8_t>(), d->lmds.as<LmdRec>(), d->err.as<uint32_t>(), d->work.as<uint32_t>(), d->n_sms, s,
                          d->timer.enabled ? d->timer.ev[d->timer.n] : nullptr);
        if (d->timer.enabled) d->timer.n++;  // literals
        d->launches += 2;
    } else {
        d->timer.mark(s);
    }
    d->timer.mark(s);  // lmds
    // Expansion: a warp per stream when there are enough streams to fill the machine that way (64 warps x 148 SMs);
    // otherwise a CTA per stream, whose 7 worker warps share one stream through a shared-memory window.
    const bool use_cta = d->expand_mode == 2 || (d->expand_mode == 0 && n < (size_t)d->n_sms * 16);
    if (!use_cta)
        launch_expand(src, src_off, src_len, dst, dst_off, dst_cap, d->counts.as<StreamCounts>(), d->blocks.as<BlockDesc>(), d->fse.as<FseDesc>(),
                      d->lits.as<uint8_t>(), d->lmds.as<LmdRec>(), d->err.as<uint32_t>(), n, s);
    else
        launch_expand_cta(src, src_off, src_len, dst, dst_off, dst_cap, d->counts.as<StreamCounts>(), d->blocks.as<BlockDesc>(),
                          d->fse.as<FseDesc>(), d->lits.as<uint8_t>(), d->lmds.as<LmdRec>(), raw_total, d->err.as<uint32_t>(), n, d->work.as<uint32_t>() + 2,
                          d->n_sms, s);
    d->timer.mark(s);  // expand
    launch_finish(d->err.as<uint32_t>(), raw_total, out_len, status, n, s);
    d->timer.mark(s);  // finish
    d->launches += 2;
    CK(d, cudaGetLastError());
    CK(d, cudaStreamSynchronize(s));
    d->timer.finish();
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
    for (DevBuf *b : {&d->counts, &d->err, &d->raw_total, &d->totals_dev, &d->blocks, &d->fse, &d->lits, &d->lmds, &d->work}) b->release();
    d->totals_host.release();
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

int lzfse_b200_decode_batch_device(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                   const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n, void *stream) {
    if (!d || (n && (!src_off || !src_len || !dst_off || !dst_cap || !out_len || !status))) return LZFSE_B200_INVALID_ARGUMENT;
    DeviceGuard g(d->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    cudaStream_t s = stream ? (cudaStream_t)stream : d->own_stream;
    return decode_batch_device_impl(d, src, src_off, src_len, dst, dst_off, dst_cap, out_len, status, nullptr, nullptr, n, s, false);
}

int lzfse_b200_decode_probe_batch_device(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len,
                                         uint64_t *raw_len, uint32_t *n_blocks, int32_t *status, size_t n, void *stream) {
    if (!d || (n && (!src_off || !src_len || !raw_len || !status))) return LZFSE_B200_INVALID_ARGUMENT;
    DeviceGuard g(d->device);
    if (!g.ok) return LZFSE_B200_CUDA_ERROR;
    cudaStream_t s = stream ? (cudaStream_t)stream : d->own_stream;
    return decode_batch_device_impl(d, src, src_off, src_len, nullptr, nullptr, nullptr, nullptr, status, raw_len, n_blocks, n, s, true);
}

// Host-buffer variants: stage the bytes on the device (host_util.h), run the device path, copy back.
//
// Large dense batches are cut into a few consecutive slices so that the upload of slice k+1, the kernels
// of slice k and the download of slice k-1 overlap (PCIe is full duplex).  Few slices only: the entropy
// stages are bound by the serial latency of one block, so every extra slice costs that latency again.
static int decode_batch_host_pipelined(lzfse_b200_decoder *d, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst,
                                       const uint64_t *dst_off, uint64_t *out_len, int32_t *status, size_t n, int n_slices) {
    cudaStream_t s = d->own_stream;
    HostStage &st = d->stage;
    if (!st.copy_in) {
        CK(d, cudaStreamCreateWithFlags(&st.copy_in, cudaStreamNonBlocking));
        CK(d, cudaStreamCreateWithFlags(&st.copy_out, cudaStreamNonBlocking));
        for (auto &e : st.ev_in) CK(d, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const uint64_t *pin = st.pin.as<uint64_t>();
    uint64_t *dd = st.desc.as<uint64_t>();
    // per-stream results are written by k_finish directly into pinned host memory (after the 4n descriptor words)
    uint64_t *d_out_len = st.pin.as<uint64_t>() + 4 * n;
    int32_t *d_status = reinterpret_cast<int32_t *>(d_out_len + n);
    // Slice boundaries by bytes crossing the bus.  The download engine is the bound, and it can only start
    // once the first slice's kernels are done (a fixed latency), so the first slice is the smallest.
    size_t cut[5] = {0, 0, 0, 0, n};
    {
        static const uint32_t share_pct[3][4] = {{100, 100, 100, 100}, {40, 100, 100, 100}, {20, 55, 100, 100}};
        uint64_t total = 0, acc = 0;
        for (size_t i = 0; i < n; i++) total += src_len[i] + pin[3 * n + i];
        int k = 1;
        for (size_t i = 0; i < n && k < n_slices; i++) {
            acc += src_len[i] + pin[3 * n + i];
            if (acc >= total / 100 * share_pct[n_slices - 1][k - 1]) cut[k++] = i + 1;
        }
        for (; k < n_slices; k++) cut[k] = n;
        cut[n_slices] = n;
    }
    // uploads, in slice order, on their own stream
    for (int k = 0; k < n_slices; k++) {
        const size_t i0 = cut[k], i1 = cut[k + 1];
        if (i1 > i0) {
            const uint64_t lo = src_off[i0], hi = src_off[i1 - 1] + src_len[i1 - 1];
            if (hi > lo) CK(d, cudaMemcpyAsync(st.src.as<uint8_t>() + (lo - st.src_lo), src + lo, hi - lo, cudaMemcpyHostToDevice, st.copy_in));
        }
        CK(d, cudaEventRecord(st.ev_in[k], st.copy_in));
    }
    uint64_t launches = 0;
    const bool dbg = getenv("LZB_DEBUG") != nullptr;
    auto now = []() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
    const double t_start = now();
    for (int k = 0; k < n_slices; k++) {
        const size_t i0 = cut[k], i1 = cut[k + 1], m = i1 - i0;
        if (m == 0) continue;
        if (dbg) fprintf(stderr, "[lzb] slice %d: streams [%zu,%zu) start %.2f ms\n", k, i0, i1, now() - t_start);
        CK(d, cudaStreamWaitEvent(s, st.ev_in[k], 0));
        int rc = decode_batch_device_impl(d, st.src.as<uint8_t>(), dd + i0, dd + n + i0, st.dst.as<uint8_t>(), dd + 2 * n + i0, dd + 3 * n + i0,
                                          d_out_len + i0, d_status + i0, nullptr, nullptr, m, s, false);
        if (rc) return rc;
        launches += d->launches;
        memcpy(out_len + i0, d_out_len + i0, m * sizeof(uint64_t));  // the call above synchronised the stream
        memcpy(status + i0, d_status + i0, m * sizeof(int32_t));
        if (dbg) fprintf(stderr, "[lzb] slice %d: kernels done %.2f ms\n", k, now() - t_start);
        rc = fetch_outputs(d, st, dst, dst_off, out_len, status, n, st.copy_out, i0, i1);  // overlaps the next slice's kernels
        if (rc) return rc;
        if (dbg) fprintf(stderr, "[lzb] slice %d: download enqueued %.2f ms\n", k, now() - t_start);
    }
    d->launches = launches;
    CK(d, cudaStreamSynchronize(st.copy_out));
    if (dbg) fprintf(stderr, "[lzb] all downloads done %.2f ms\n", now() - t_start);
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
    if (want_pipeline && st.src_mirrored && st.dst_mirrored) return decode_batch_host_pipelined(d, src, src_off, src_len, dst, dst_off, out_len, status, n, 3);
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
