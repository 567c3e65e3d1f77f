// host_util.h -- host-side helpers shared by api.cu (decoder) and encode.cu (encoder).
#pragma once
#include <cstdint>
#include <string>
#include <cuda_runtime.h>

#include "../../include/lzfse_b200.h"

namespace lzb {

// A device buffer that only ever grows (no per-call allocation once warm).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); want = bytes; e = cudaMalloc(&p, want); }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};
struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; prev = -1; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define LZB_CK(h, call)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            (h)->last_error = std::string(#call) + ": " + cudaGetErrorString(_e);              \
            cudaGetLastError();                                                                \
            return _e == cudaErrorMemoryAllocation ? LZFSE_B200_OUT_OF_MEMORY : LZFSE_B200_CUDA_ERROR; \
        }                                                                                      \
    } while (0)

// Optional per-stage CUDA-event timing on the launching stream (measurement aid for bench.py).
struct StageTimer {
    static constexpr int kMax = 16;
    bool enabled = false;
    cudaEvent_t ev[kMax + 1] = {};
    int n = 0;          // events recorded in the current call
    int n_done = 0;     // stages of the last finished call
    float ms[kMax] = {};
    void begin(cudaStream_t s) {
        n = 0;
        if (!enabled) return;
        for (int i = 0; i <= kMax; i++)
            if (!ev[i]) cudaEventCreate(&ev[i]);
        cudaEventRecord(ev[0], s); n = 1;
    }
    void mark(cudaStream_t s) { if (enabled && n <= kMax) { cudaEventRecord(ev[n], s); n++; } }
    void finish() {  // call after the stream was synchronised
        if (!enabled) return;
        n_done = n > 0 ? n - 1 : 0;
        for (int i = 0; i < n_done; i++) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
    }
    void release() { for (int i = 0; i <= kMax; i++) if (ev[i]) { cudaEventDestroy(ev[i]); ev[i] = nullptr; } }
};

// Host-buffer staging shared by both directions.  The device copy mirrors the host layout when the
// streams cover their byte range densely (one H2D copy); sparse layouts are packed stream by stream.
struct HostStage {
    DevBuf src, dst, desc, res;
    PinnedBuf pin;
    uint64_t src_lo = 0, dst_lo = 0;
    bool src_mirrored = false, dst_mirrored = false;
    size_t pin_n = 0;  // batch size the pin[] layout was built for
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // side streams of the pipelined host path
    cudaEvent_t ev_in[8] = {};
    cudaEvent_t ev_scan[8] = {};
    cudaEvent_t ev_desc = nullptr;
    void release() {
        src.release(); dst.release(); desc.release(); res.release(); pin.release();
        if (copy_in) cudaStreamDestroy(copy_in);
        if (copy_out) cudaStreamDestroy(copy_out);
        for (auto &e : ev_in) if (e) cudaEventDestroy(e);
        if (ev_desc) cudaEventDestroy(ev_desc);
        for (auto &e : ev_scan) if (e) cudaEventDestroy(e);
        for (auto &e : ev_in) e = nullptr;
        for (auto &e : ev_scan) e = nullptr;
        ev_desc = nullptr;
        copy_in = copy_out = nullptr;
    }
};

// Fills pin[0..n) = device src offsets, pin[n..2n) = src_len and uploads the source bytes.
template <class H>
int stage_sources(H *h, HostStage &st, const uint8_t *src, const uint64_t *src_off, const uint64_t *src_len, size_t n, size_t pin_words,
                  cudaStream_t s, bool defer_dense_upload = false) {
    uint64_t lo = ~0ull, hi = 0, sum = 0;
    for (size_t i = 0; i < n; i++) {
        if (src_off[i] < lo) lo = src_off[i];
        if (src_off[i] + src_len[i] > hi) hi = src_off[i] + src_len[i];
        sum += src_len[i];
    }
    if (n == 0) { lo = 0; hi = 0; }
    LZB_CK(h, st.pin.reserve(pin_words * sizeof(uint64_t)));
    st.pin_n = n;
    uint64_t *pin = st.pin.as<uint64_t>();
    const bool dense = (hi - lo) <= 2 * sum + 4096;
    st.src_mirrored = dense;
    if (dense) {
        LZB_CK(h, st.src.reserve(hi - lo + 64));
        if (hi > lo && !defer_dense_upload) LZB_CK(h, cudaMemcpyAsync(st.src.p, src + lo, hi - lo, cudaMemcpyHostToDevice, s));
        for (size_t i = 0; i < n; i++) { pin[i] = src_off[i] - lo; pin[n + i] = src_len[i]; }
    } else {
        uint64_t off = 0;
        for (size_t i = 0; i < n; i++) { pin[i] = off; pin[n + i] = src_len[i]; off += (src_len[i] + 15) & ~15ull; }
        LZB_CK(h, st.src.reserve(off + 64));
        for (size_t i = 0; i < n; i++)
            if (src_len[i]) LZB_CK(h, cudaMemcpyAsync(st.src.template as<uint8_t>() + pin[i], src + src_off[i], src_len[i], cudaMemcpyHostToDevice, s));
    }
    st.src_lo = lo;
    return LZFSE_B200_OK;
}

// Fills pin[2n..3n) = device dst offsets, pin[3n..4n) = dst_cap and sizes the device output buffer.
template <class H>
int stage_outputs(H *h, HostStage &st, const uint64_t *dst_off, const uint64_t *dst_cap, size_t n) {
    uint64_t lo = ~0ull, hi = 0, sum = 0;
    for (size_t i = 0; i < n; i++) {
        if (dst_off[i] < lo) lo = dst_off[i];
        if (dst_off[i] + dst_cap[i] > hi) hi = dst_off[i] + dst_cap[i];
        sum += dst_cap[i];
    }
    if (n == 0) { lo = 0; hi = 0; }
    uint64_t *pin = st.pin.as<uint64_t>();
    st.dst_mirrored = (hi - lo) <= 2 * sum + 4096;
    if (st.dst_mirrored) {
        LZB_CK(h, st.dst.reserve(hi - lo + 64));
        for (size_t i = 0; i < n; i++) { pin[2 * n + i] = dst_off[i] - lo; pin[3 * n + i] = dst_cap[i]; }
    } else {
        uint64_t off = 0;
        for (size_t i = 0; i < n; i++) { pin[2 * n + i] = off; pin[3 * n + i] = dst_cap[i]; off += (dst_cap[i] + 15) & ~15ull; }
        LZB_CK(h, st.dst.reserve(off + 64));
    }
    st.dst_lo = lo;
    return LZFSE_B200_OK;
}

// Copies every successful stream's bytes back; neighbours that are adjacent on both sides share a copy.
template <class H>
int fetch_outputs(H *h, HostStage &st, uint8_t *dst, const uint64_t *dst_off, const uint64_t *out_len, const int32_t *status, size_t n,
                  cudaStream_t s, size_t i0 = 0, size_t i1 = ~(size_t)0) {
    const uint64_t *pin = st.pin.as<uint64_t>();
    size_t i = i0;
    if (i1 < n) n = i1;  // only streams [i0, i1); pin[] is still indexed with the full batch size
    const size_t full = st.pin_n;
    while (i < n) {
        if (status[i] != 0 || out_len[i] == 0) { i++; continue; }
        uint64_t h0 = dst_off[i], d0 = pin[2 * full + i], run = out_len[i];
        size_t j = i + 1;
        while (j < n && status[j] == 0 && dst_off[j] == h0 + run && pin[2 * full + j] == d0 + run) { run += out_len[j]; j++; }
        LZB_CK(h, cudaMemcpyAsync(dst + h0, st.dst.template as<uint8_t>() + d0, run, cudaMemcpyDeviceToHost, s));
        i = j;
    }
    return LZFSE_B200_OK;
}

}  // namespace lzb
