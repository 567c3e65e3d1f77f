// encode_long.cuh -- the front end for bvx2 streams longer than kFastMaxLen (BASELINE configs[3]: 16 MiB streams).
// Included by encode.cu inside namespace lzb, after k_enc_replay (it uses Match, TSink, hash_u, ld4u, ld8u).
//
// Round 1 parsed such a stream with ONE warp (k_enc_parse): 16 MiB in 3 s.  Two facts let a single stream use the
// whole machine without changing one byte of the frame:
//
//  1. What a position finds in the history does not depend on the parse (top of encode.cu).  The reference's
//     HistoryTable (encode/history.rs:24-31,101-131) is the hash chain prev[p] = newest q < p in p's bucket, and
//     that chain can be built for all 32 Ki-position pieces of a stream at once: k_long_chain links the positions
//     inside a piece (bucket heads in shared memory) and leaves the newest position per bucket, k_long_carry turns
//     those into "newest position before this piece" per bucket (a running maximum over the pieces, one thread per
//     bucket; computed first, from an order-free maximum per piece), and a link that would leave its piece continues there.  k_long_find then evaluates find_match
//     (frontend_bytes.rs:214-244) for EVERY position, one thread each, and writes the same per-position word
//     k_enc_find writes.
//  2. The sequential part (backward limit, Match::select, frontend_bytes.rs:160-211,261-302) started from a clean
//     state at an arbitrary position falls into step with the true parse after a few matches.  k_long_replay runs
//     it for every kRSeg-position segment on its own (thread per segment, speculative), recording each pushed
//     match with the state it leaves behind; k_long_stitch walks a stream's segments in order with the TRUE state,
//     replays from it until that state equals a recorded one (or takes the segment whole when the true state at
//     its border is equivalent to a clean start), and from there on the speculative output IS the true output.
//     tests/model/long_parse_model.c is this algorithm on the CPU, checked against the reference port's sequential front end.
//
//  k_long_seg_stats / k_long_blocks / k_long_write_packs finally turn a stream's match list into packs and block records
//  (Buffer::push with its L / M splits and block closing, fse/buffer.rs:45-117); see there.
//  Everything downstream (k_enc_fse_blocks, k_enc_assemble) is shared with the other paths.

constexpr uint32_t kNoPos = 0xFFFFFFFFu;
constexpr uint32_t kCSeg = kLongCSeg;         // positions per chain piece (16-bit bucket heads)
constexpr uint32_t kRSeg = kLongRSeg;         // positions per speculative replay segment
static_assert(kRSeg % 32 == 0 && kCSeg % kRSeg == 0, "segments are whole 32-position groups");
constexpr uint32_t kEmitCap = kRSeg / 4 + 8;  // matches are >= 4 bytes and do not overlap
constexpr uint32_t kSpecStates = kRSeg / 64 < 256 ? kRSeg / 64 : 256;  // pushed matches per segment that carry the state they leave behind
constexpr uint32_t kSegChunks = (kEmitCap + 31) / 32;  // 32-match chunks of a segment's list (k_long_seg_stats leaves their sums)
constexpr uint32_t kLongLaneCap = 64;

struct LongSeg { uint32_t stream, k; };
struct FrontState { uint32_t cur, lit, p_idx, p_midx, p_len; };
struct LongSegOut {
    uint32_t n_spec;     // matches pushed by the speculative replay
    uint32_t lim0;       // a backward extension before the first push stopped at the literal limit
    uint32_t good0;      // the first candidate met was >= GOOD_MATCH_LEN long
    uint32_t cand0;      // position of the first candidate met (kNoPos: none)
    FrontState exit;     // state when the cursor left the segment
    uint32_t n_fix;      // matches the stitch pushed before it fell into step
    uint32_t from;       // first speculative match that is part of the true parse (n_spec: none)
    FrontState a_out;    // k_long_stitch_a: true state behind this segment IF the previous segment's exit state was true
    uint32_t a_ok;       // k_long_stitch_a: ... and the segment fell into step (a_out == exit)
};

// ---- segment descriptors -----------------------------------------------------------------------
__global__ void k_long_segs(const uint32_t *__restrict__ long_list, uint32_t n_long, const EncStream *__restrict__ streams, LongSeg *cseg, LongSeg *rseg) {
    const uint32_t slot = blockIdx.x;
    if (slot >= n_long) return;
    const uint32_t si = long_list[slot];
    const EncStream st = streams[si];
    for (uint32_t k = threadIdx.x; k < st.n_cseg; k += blockDim.x) cseg[st.cseg_base + k] = LongSeg{si, k};
    for (uint32_t k = threadIdx.x; k < st.n_rseg; k += blockDim.x) rseg[st.rseg_base + k] = LongSeg{si, k};
}

// ---- chain inside a piece ----------------------------------------------------------------------
// One WARP per piece, six of them per SM, bucket heads (16-bit) in shared memory, 32 positions per step in order.  Two
// positions of a step share a bucket in ~3 % of the steps; instead of finding the peers of every step (match.any costs
// ~350 cycles when all 32 values differ, which made a separate info kernel 1.2 ms per 128 MiB) every lane stores its
// position into its bucket head and reads it back: if all 32 read their own position there were no peers and the heads are
// right; otherwise this step is redone with match.any (newest lane owns the head, lanes with a lower peer link to it).
// Writes {link, the four bytes at the position} per position: find_match accepts or rejects a candidate with the same load
// that yields the next hop.
constexpr uint32_t kCSegShift = 15;
static_assert(kCSeg == 1u << kCSegShift, "chain piece size");
constexpr uint32_t kChainSmem = (1u << kHashBits) * 2;
constexpr uint32_t kChainPerSm = 6;
constexpr uint32_t kHeadsSmem = (1u << kHashBits) * 4;
// Newest position per bucket and piece, order-free (a maximum): lets k_long_carry run BEFORE the chain, so that the chain
// can write finished links (a link that leaves its piece continues with the newest position of the bucket before the piece).
__global__ void __launch_bounds__(512)
k_long_heads(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
             const EncStream *__restrict__ streams, const LongSeg *__restrict__ cseg, uint32_t *__restrict__ seg_head) {
    extern __shared__ __align__(16) uint8_t csm[];
    uint32_t *hd = reinterpret_cast<uint32_t *>(csm);
    const LongSeg sg = cseg[blockIdx.x];
    const EncStream st = streams[sg.stream];
    const uint32_t end = (uint32_t)src_len[sg.stream] - 3;
    const uint32_t B = sg.k * kCSeg, n_pos = end - B < kCSeg ? end - B : kCSeg;
    const uint8_t *src = src_base + src_off[sg.stream] + B;
    for (uint32_t t = threadIdx.x; t < (1u << kHashBits); t += 512) hd[t] = 0;
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < n_pos; t += 512) atomicMax(&hd[hash_u(ld4u(src + t), false)], t + 1);
    __syncthreads();
    uint32_t *sh = seg_head + ((size_t)st.cseg_base + sg.k) * (1u << kHashBits);
    for (uint32_t t = threadIdx.x; t < (1u << kHashBits); t += 512) sh[t] = hd[t] ? B + hd[t] - 1 : kNoPos;
}
__global__ void __launch_bounds__(32)
k_long_chain(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
             const EncStream *__restrict__ streams, const LongSeg *__restrict__ cseg, uint32_t n_cseg, uint2 *__restrict__ pairs,
             const uint32_t *__restrict__ seg_head, uint32_t *work_counter) {
    extern __shared__ __align__(16) uint8_t csm[];
    volatile uint16_t *head = reinterpret_cast<volatile uint16_t *>(csm);
    const uint32_t lane = threadIdx.x;
    for (;;) {
        uint32_t cs = 0;
        if (lane == 0) cs = atomicAdd(work_counter, 1u);
        cs = __shfl_sync(0xFFFFFFFFu, cs, 0);
        if (cs >= n_cseg) break;
        const LongSeg sg = cseg[cs];
        const EncStream st = streams[sg.stream];
        const uint32_t end = (uint32_t)src_len[sg.stream] - 3;
        const uint32_t B = sg.k * kCSeg, n_pos = end - B < kCSeg ? end - B : kCSeg;
        const uint8_t *src = src_base + src_off[sg.stream] + B;
        asm volatile("" : "+l"(src));
        uint2 *pv = pairs + st.long_off + B;
        const uint32_t *before = seg_head + ((size_t)st.cseg_base + sg.k) * (1u << kHashBits);  // (after k_long_carry) newest position before the piece
        for (uint32_t t = lane; t < (1u << kHashBits) * 2 / 16; t += 32) reinterpret_cast<uint4 *>(csm)[t] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        __syncwarp();
        // The bytes of the next eight steps are on their way while eight are worked on, and the links of eight steps are
        // stored while the next eight are worked on (a link that leaves the piece is a load from `before`).
        constexpr uint32_t kSteps = 8;
        uint32_t nxt[kSteps], dval[kSteps], dlink[kSteps];
#pragma unroll
        for (uint32_t k = 0; k < kSteps; k++) { const uint32_t p = k * 32 + lane; nxt[k] = ld4u(src + (p < n_pos ? p : n_pos - 1)); dval[k] = 0; dlink[k] = 0; }
        for (uint32_t u0 = 0; u0 < n_pos + 32 * kSteps; u0 += 32 * kSteps) {
            uint32_t val[kSteps], link[kSteps];
#pragma unroll
            for (uint32_t k = 0; k < kSteps; k++) {
                val[k] = nxt[k];
                const uint32_t p = u0 + (kSteps + k) * 32 + lane;
                nxt[k] = ld4u(src + (p < n_pos ? p : n_pos - 1));
            }
#pragma unroll
            for (uint32_t k = 0; k < kSteps; k++) {
                const uint32_t b0 = u0 + k * 32, p = b0 + lane;
                link[k] = kNoPos;
                if (b0 >= n_pos) continue;
                const bool act = p < n_pos;
                const uint32_t h = hash_u(val[k], false);
                uint32_t l16 = 0xFFFFu;
                if (act) l16 = head[h];
                __syncwarp();
                if (act) head[h] = (uint16_t)p;
                __syncwarp();
                const bool lost = act && head[h] != p;
                if (__any_sync(0xFFFFFFFFu, lost)) {  // peers in this step: HistoryTable::push in order
                    const uint32_t m = __match_any_sync(0xFFFFFFFFu, act ? h : 0xFFFF0000u + lane);
                    const uint32_t lower = m & lanemask_lt();
                    if (act && (m >> lane) == 1u) head[h] = (uint16_t)p;  // the newest position of the bucket
                    if (lower) l16 = b0 + (31 - __clz(lower));
                    __syncwarp();
                }
                if (act) link[k] = l16 != 0xFFFFu ? B + l16 : before[h];
            }
            // the previous eight steps' pairs
            if (u0 != 0) {
#pragma unroll
                for (uint32_t k = 0; k < kSteps; k++) {
                    const uint32_t p = u0 - 32 * kSteps + k * 32 + lane;
                    if (p < n_pos) pv[p] = make_uint2(dlink[k], dval[k]);
                }
            }
#pragma unroll
            for (uint32_t k = 0; k < kSteps; k++) { dval[k] = val[k]; dlink[k] = link[k]; }
        }
        __syncwarp();
    }
}

// seg_head[piece][bucket]: newest position of the bucket inside the piece  ->  newest position BEFORE the piece.
__global__ void k_long_carry(const uint32_t *__restrict__ long_list, uint32_t n_long, const EncStream *__restrict__ streams, uint32_t *seg_head) {
    const uint32_t slot = blockIdx.x / ((1u << kHashBits) / 256);
    const uint32_t h = (blockIdx.x % ((1u << kHashBits) / 256)) * 256 + threadIdx.x;
    if (slot >= n_long) return;
    const EncStream st = streams[long_list[slot]];
    uint32_t run = kNoPos;
    uint32_t *q = seg_head + (size_t)st.cseg_base * (1u << kHashBits) + h;
    for (uint32_t k = 0; k < st.n_cseg; k++, q += (1u << kHashBits)) {
        const uint32_t t = *q;
        *q = run;
        if (t != kNoPos) run = t;
    }
}

// ---- find_match for every position ---------------------------------------------------------------
// Forward length of src[a..] against src[b..], whole warp, 8 bytes per lane and step, from `l` up to `lim`.
__device__ __forceinline__ uint32_t gwarp_match_inc(const uint8_t *src, uint32_t a, uint32_t b, uint32_t l, uint32_t lim, uint32_t lane) {
    while (l < lim) {
        const uint32_t off = l + lane * 8;
        uint64_t y = 0;
        if (off + 8 <= lim) y = ld8u(src + a + off) ^ ld8u(src + b + off);
        else for (uint32_t k = 0; off + k < lim; k++) y |= (uint64_t)(src[a + off + k] ^ src[b + off + k]) << (8 * k);
        const uint32_t diff = __ballot_sync(0xFFFFFFFFu, y != 0);
        if (diff) {
            const int j = __ffs(diff) - 1;
            const uint64_t yj = __shfl_sync(0xFFFFFFFFu, y, j);
            return l + j * 8 + ((__ffsll((long long)yj) - 1) >> 3);
        }
        l += 256;
    }
    return lim;
}

constexpr int kLFindThreads = 256;
#ifndef LZB_LFIND_MINB
#define LZB_LFIND_MINB 8   // resident CTAs per SM: 5 / 6 / 8 (47 / 40 / 32 registers) measured 4.03 / 3.76 / 3.64 ms
#endif
__global__ void __launch_bounds__(kLFindThreads, LZB_LFIND_MINB)
k_long_find(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
            const EncStream *__restrict__ streams, const StreamCounts *__restrict__ bases, const LongSeg *__restrict__ cseg,
            const uint2 *__restrict__ pairs, uint32_t *__restrict__ words) {
    const uint32_t lane = threadIdx.x & 31;
    const LongSeg sg = cseg[blockIdx.x / (kCSeg / kLFindThreads)];
    const EncStream st = streams[sg.stream];
    const uint8_t *src = src_base + src_off[sg.stream];
    asm volatile("" : "+l"(src));
    const uint32_t len = (uint32_t)src_len[sg.stream], end = len - 3;
    const uint32_t p = sg.k * kCSeg + (blockIdx.x % (kCSeg / kLFindThreads)) * kLFindThreads + threadIdx.x;
    if (p - lane >= end) return;  // whole warp beyond the last position
    const bool act = p < end;
    const uint2 *pv = pairs + st.long_off;  // {newest earlier position of the bucket inside the piece, the four bytes at this position}
    uint32_t best_len = 0, best_c = 0, n_sat = 0;
    uint32_t cs[4] = {0, 0, 0, 0}, ls[4] = {0, 0, 0, 0};
    const uint32_t maxl = act ? len - p : 0;
    if (act) {
        const uint2 me = pv[p];
        const uint32_t val = me.y;
        const uint32_t lim = maxl < kLongLaneCap ? maxl : kLongLaneCap;
        uint32_t c = me.x;
        const uint64_t p8 = maxl >= 12 ? ld8u(src + p + 4) : 0ull;  // bytes 4..11 of this position, shared by its candidates
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (c == kNoPos || p - c > kMaxDValue) break;  // newest first, up to the first one out of range (frontend_bytes.rs:214-244)
            const uint2 ce = pv[c];
            const uint32_t cn = ce.x;
            if (ce.y == val) {
                uint32_t l = 4;
                if (maxl >= 12) {  // most matches end inside the next eight bytes: one load of the candidate decides
                    const uint64_t y0 = p8 ^ ld8u(src + c + 4);
                    if (y0) { l = 4 + ((__ffsll((long long)y0) - 1) >> 3); goto ext_done; }
                    l = 12;
                }
                while (l + 8 <= lim) {
                    const uint64_t y = ld8u(src + p + l) ^ ld8u(src + c + l);
                    if (y) { l += (__ffsll((long long)y) - 1) >> 3; goto ext_done; }
                    l += 8;
                }
                while (l < lim && src[p + l] == src[c + l]) l++;
            ext_done:
                if (l == kLongLaneCap && l < maxl) { cs[n_sat] = c; n_sat++; }
                if (l > best_len) { best_len = l; best_c = c; }
            }
            c = cn;
        }
    }
    // Candidates that reached the per-lane cap: their lengths up to what the word can hold, measured by the whole warp.
    // Strictly longest wins, newest first.  Two candidates that both pass the word's limit are compared further; once both
    // have matched as many bytes as the larger of their distances the source is periodic with both distances and the two
    // lengths are equal (Fine & Wilf), so the comparison stops there and the newer one wins.
    uint32_t todo = __ballot_sync(0xFFFFFFFFu, n_sat != 0);
    while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t pj = __shfl_sync(0xFFFFFFFFu, p, j), nj = __shfl_sync(0xFFFFFFFFu, n_sat, j), mj = len - pj;
        const uint32_t wl = mj < kWordLenSat ? mj : kWordLenSat;
        uint32_t lj[4] = {0, 0, 0, 0}, cj[4];
        uint32_t n_full = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            cj[k] = __shfl_sync(0xFFFFFFFFu, cs[k], j);
            if ((uint32_t)k < nj) { lj[k] = gwarp_match_inc(src, pj, cj[k], kLongLaneCap, wl, lane); n_full += lj[k] == wl; }
        }
        if (n_full >= 2 && wl < mj) {
            // lockstep beyond the word's limit, 256 bytes per round, only among those still matching
            uint32_t alive = 0, maxd = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) if ((uint32_t)k < nj && lj[k] == wl) { alive |= 1u << k; maxd = max(maxd, pj - cj[k]); }
            uint32_t off = wl;
            while (__popc(alive) >= 2 && off < mj && off < maxd) {
                const uint32_t hi = off + 256 < mj ? off + 256 : mj;
#pragma unroll
                for (int k = 0; k < 4; k++) if (alive & (1u << k)) {
                    const uint32_t e = gwarp_match_inc(src, pj, cj[k], off, hi, lane);
                    lj[k] = e;
                    if (e < hi) alive &= ~(1u << k);
                }
                off = hi;
            }
            // the survivors are the longest (equal among themselves): make the newest of them win outright
            if (alive) { const int w = __ffs(alive) - 1;
#pragma unroll
                for (int k = 0; k < 4; k++) if (k == w) lj[k] = 0xFFFFFFFFu; }
        }
        if (lane == (uint32_t)j) {
#pragma unroll
            for (int k = 0; k < 4; k++) ls[k] = lj[k];
        }
    }
    uint32_t word = 0;
    if (act && best_len) {
        if (n_sat) {
            uint32_t bl = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) if ((uint32_t)k < n_sat && ls[k] > bl) { bl = ls[k]; best_c = cs[k]; }
            best_len = bl;
        }
        // backward length (match_kit/match_fast.rs:61-89) up to what the word can hold: the 16 bytes before the two positions
        uint32_t bw = 0;
        const uint32_t blim = best_c < kWordBwSat ? best_c : kWordBwSat;
        if (best_c >= 16) {
            const uint64_t y1 = ld8u(src + p - 8) ^ ld8u(src + best_c - 8);
            if (y1) bw = (uint32_t)__clzll((long long)y1) >> 3;
            else { const uint64_t y2 = ld8u(src + p - 16) ^ ld8u(src + best_c - 16); bw = 8 + (y2 ? (uint32_t)__clzll((long long)y2) >> 3 : 8u); }
            bw = bw < blim ? bw : blim;
        } else {
            while (bw < blim && src[p - bw - 1] == src[best_c - bw - 1]) bw++;
        }
        word = (p - best_c) | ((best_len < kWordLenSat ? best_len : kWordLenSat) << 18) | (bw << 28);
    }
    {   // positions without a candidate carry the distance to the next position of their 32-group that has one
        const uint32_t nz = __ballot_sync(0xFFFFFFFFu, word != 0);
        if (word == 0) {
            const uint32_t next = lane == 31 ? 0u : nz & (0xFFFFFFFEu << lane);
            word = (next ? (uint32_t)__ffs(next) - 1u - lane : 32u - lane) << 18;
        }
    }
    if (act) words[bases[sg.stream].n_fse + p] = word;
}

// ---- one position of the sequential front end ------------------------------------------------------
// FrontendBytes::match_any's loop body (frontend_bytes.rs:185-211,261-302) over the per-position words.  Returns true
// when a match is pushed to the back end (sel).  `lim_flag` is set when a backward extension stopped at the literal
// limit although the candidate's own start would have allowed more; `cand`/`good` describe the first candidate met.
// A word whose length field is saturated (>= 1 023 bytes) needs the match measured to its end.  `ext_len` != 0 is that
// length, already known; otherwise the caller either lets the step measure it (one thread, 8 bytes per iteration: the stitch,
// where it is rare) or, with kCoop, gets the request back (`need_ext`, state untouched) and has its whole warp measure it --
// on periodic data EVERY segment's clean start finds a match that runs to the end of the stream.
template <bool kCoop>
__device__ __forceinline__ bool front_step(const uint8_t *src, uint32_t len, uint32_t end, uint32_t w, FrontState &s, Match &sel, uint32_t &lim_flag,
                                           uint32_t &cand, uint32_t &good, uint32_t ext_len, bool &need_ext) {
    const uint32_t cur = s.cur;
    if ((w & 0x3FFFFu) == 0) { s.cur = cur + ((w >> 18) & 0x3FFu); return false; }
    Match inc;
    inc.idx = cur;
    inc.match_idx = cur - (w & 0x3FFFFu);
    inc.match_len = (w >> 18) & 0x3FFu;
    if (inc.match_len == kWordLenSat) {  // the word's length field is saturated
        if (ext_len) inc.match_len = ext_len;
        else if (kCoop) { need_ext = true; return false; }
        else {
            const uint32_t maxl = len - cur;
            while (inc.match_len + 8 <= maxl) {
                const uint64_t y = ld8u(src + cur + inc.match_len) ^ ld8u(src + inc.match_idx + inc.match_len);
                if (y) { inc.match_len += (__ffsll((long long)y) - 1) >> 3; goto fwd_done; }
                inc.match_len += 8;
            }
            while (inc.match_len < maxl && src[cur + inc.match_len] == src[inc.match_idx + inc.match_len]) inc.match_len++;
        fwd_done:;
        }
    }
    {   // match_dec (:261-268)
        const uint32_t lit = cur - s.lit;
        const uint32_t lim = lit < inc.match_idx ? lit : inc.match_idx;
        const uint32_t bw = w >> 28;
        uint32_t dec = bw < lim ? bw : lim;
        if (bw == kWordBwSat) while (dec < lim && src[inc.idx - dec - 1] == src[inc.match_idx - dec - 1]) dec++;
        if (dec == lit && dec < inc.match_idx) lim_flag = 1;
        inc.idx -= dec; inc.match_idx -= dec; inc.match_len += dec;
    }
    if (cand == kNoPos) { cand = cur; good = inc.match_len >= kGoodMatchLen; }
    bool have = true;
    sel.idx = s.p_idx; sel.match_idx = s.p_midx; sel.match_len = s.p_len;
    if (inc.match_len >= kGoodMatchLen) { sel = inc; s.p_len = 0; }
    else if (s.p_len == 0) { s.p_idx = inc.idx; s.p_midx = inc.match_idx; s.p_len = inc.match_len; have = false; }
    else if ((int32_t)(s.p_idx + s.p_len - inc.idx) <= 0) { s.p_idx = inc.idx; s.p_midx = inc.match_idx; s.p_len = inc.match_len; }
    else if (inc.match_len > s.p_len) { sel = inc; s.p_len = 0; }
    else { s.p_len = 0; }
    if (!have) { s.cur = cur + 1; return false; }
    s.lit = sel.idx + sel.match_len;
    s.cur = s.lit >= end ? end : (cur + 1 > s.lit ? cur + 1 : s.lit);
    return true;
}

// (out of line: the replay loop is register-bound and only needs this on long matches)
__device__ __noinline__ uint32_t gwarp_match_inc_far(const uint8_t *src, uint32_t a, uint32_t b, uint32_t l, uint32_t lim, uint32_t lane) {
    return gwarp_match_inc(src, a, b, l, lim, lane);
}
// ---- speculative replay, one THREAD per segment ----------------------------------------------------
// Same loop and the same word ring as k_enc_replay (see there for the ring's rules); starts clean at the segment's first
// position and stops when the cursor leaves the segment.
__global__ void __launch_bounds__(kReplayThreads, 32)
k_long_replay(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
              const StreamCounts *__restrict__ bases, const LongSeg *__restrict__ rseg, uint32_t n_rseg, const uint32_t *__restrict__ words,
              uint4 *__restrict__ spec, uint4 *__restrict__ states, LongSegOut *__restrict__ seg_out) {
    const uint32_t rs = blockIdx.x * kReplayThreads + threadIdx.x;
    const bool valid = rs < n_rseg;
    const LongSeg sg = rseg[valid ? rs : 0];
    const uint8_t *src = src_base + src_off[sg.stream];
    const uint32_t len = (uint32_t)src_len[sg.stream], end = len - 3;
    const uint32_t B = sg.k * kRSeg, se = !valid ? 0u : (end - B < kRSeg ? end : B + kRSeg);
    const uint32_t *W = words + bases[sg.stream].n_fse;
    uint4 *out = spec + (size_t)rs * kEmitCap;
    uint4 *sto = states + (size_t)rs * kSpecStates;
    FrontState s = {B, B, 0, 0, 0};
    uint32_t n_out = 0, lim_flag = 0, lim0 = 0, cand = kNoPos, good = 0, ext_len = 0, w_req = 0;
    bool active = valid, need_ext = false;

    __shared__ __align__(16) uint8_t rings[kReplayThreads * kRingStride];
    constexpr uint32_t kLag = 8;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(rings) + threadIdx.x * kRingStride;
    const uint32_t w_limit = valid ? (se + 3u) & ~3u : 0u;
    uint32_t wbase = 0xFFFFFFFFu, fetched = 0, safe = 0, mark = 0, iter = 0;
    uint4 wq = make_uint4(0, 0, 0, 0);
    auto request = [&](bool go) {
        const bool p = go && fetched < w_limit;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                     ::"r"(ring + (fetched & (kRingWords - 1)) * 4), "l"(W + (p ? fetched : 0u)), "r"((uint32_t)p) : "memory");
        fetched += p ? 4u : 0u;
    };
    for (;;) {
        __syncwarp();
        active = active && s.cur < se;
        if (!__any_sync(0xFFFFFFFFu, active)) break;
        {   // ---- ring upkeep, whole warp (k_enc_replay) ----
            const uint32_t wb = s.cur & ~3u;
            const bool restart = active && wb > fetched;
            const bool drain = __any_sync(0xFFFFFFFFu, restart);
            if (drain) asm volatile("cp.async.wait_group 0;" ::: "memory");
            if (restart) fetched = wb;
            if (drain) { safe = fetched; mark = fetched; }
            request(active && fetched < wb + kRingWords);
            request(active && fetched < wb + kRingWords);
            if (iter == 0) {
#pragma unroll
                for (int k = 0; k < 6; k++) request(active && fetched < wb + kRingWords);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(kLag - 1) : "memory");
            if ((iter & (kLag - 1)) == 0) { safe = mark; mark = fetched; }
            const bool need = active && wb + 4 > safe;
            if (__any_sync(0xFFFFFFFFu, need)) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                safe = fetched; mark = fetched;
            }
            iter++;
        }
        {   // ---- matches longer than the word can say are measured by the whole warp, 256 bytes per round ----
            const uint32_t lane = threadIdx.x;  // (one warp per CTA)
            uint32_t rm = __ballot_sync(0xFFFFFFFFu, need_ext);
            while (rm) {
                const int j = __ffs(rm) - 1;
                rm &= rm - 1;
                const uint8_t *sj = reinterpret_cast<const uint8_t *>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(src), j));
                const uint32_t a = __shfl_sync(0xFFFFFFFFu, s.cur, j), lj = __shfl_sync(0xFFFFFFFFu, len, j);
                const uint32_t b = a - (__shfl_sync(0xFFFFFFFFu, w_req, j) & 0x3FFFFu);
                const uint32_t e = gwarp_match_inc_far(sj, a, b, kWordLenSat, lj - a, lane);
                if (lane == (uint32_t)j) { ext_len = e; need_ext = false; }
            }
        }
        if (!active) continue;
        if ((s.cur & ~3u) != wbase) {
            wbase = s.cur & ~3u;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(wq.x), "=r"(wq.y), "=r"(wq.z), "=r"(wq.w) : "r"(ring + (wbase & (kRingWords - 1)) * 4) : "memory");
        }
        const uint32_t k4 = s.cur & 3u;
        const uint32_t w = k4 == 0 ? wq.x : (k4 == 1 ? wq.y : (k4 == 2 ? wq.z : wq.w));
        Match sel;
        const bool pushed = front_step<true>(src, len, end, w, s, sel, lim_flag, cand, good, ext_len, need_ext);
        w_req = w;
        if (!need_ext) ext_len = 0;
        if (pushed) {
            if (n_out == 0) lim0 = lim_flag;
            out[n_out] = make_uint4(sel.idx, sel.match_len, sel.idx - sel.match_idx, s.cur);
            if (n_out < kSpecStates) sto[n_out] = make_uint4(s.p_idx, s.p_midx, s.p_len, 0);
            n_out++;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!valid) return;
    if (n_out == 0) lim0 = lim_flag;
    LongSegOut o;
    o.n_spec = n_out; o.lim0 = lim0; o.good0 = good; o.cand0 = cand; o.exit = s;
    o.n_fix = 0; o.from = sg.k == 0 ? 0u : n_out; o.a_out = s; o.a_ok = 0;
    seg_out[rs] = o;
}

__device__ __forceinline__ bool same_state(const FrontState &a, const FrontState &b) {
    return a.cur == b.cur && a.lit == b.lit && a.p_len == b.p_len && (a.p_len == 0 || (a.p_idx == b.p_idx && a.p_midx == b.p_midx));
}
// ---- stitch ----------------------------------------------------------------------------------------
// The true state T arrives at segment k of a stream; on return T is the true state behind it, n_fix matches have been
// written to the segment's fix list and `from` says where the speculative list joins the true parse.
// kWarp: called by a whole converged warp with identical arguments (k_long_stitch_b): the words are then fetched 32 at a time
// and handed round by shuffle -- one memory round trip per 32 positions instead of one per position, which is what the
// fallback costs when a segment never falls into step.
template <bool kWarp>
__device__ void stitch_segment(const uint8_t *src, uint32_t len, uint32_t end, const uint32_t *W, uint32_t k, FrontState &T, LongSegOut &o,
                               const uint4 *sp, const uint4 *stt, uint4 *fix, uint32_t lane = 0) {
    const uint32_t B = k * kRSeg, se = end - B < kRSeg ? end : B + kRSeg;
    o.n_fix = 0; o.from = o.n_spec;
    if (T.cur >= se) return;  // a match of an earlier segment covers this one
    if (T.cur == B && !o.lim0) {
        // The true parse stands at the segment's first position like the speculative one did, with a literal run at least
        // as long, and no backward extension of the speculative replay before its first push was cut by its shorter run:
        // both see the same candidates with the same lengths.
        if (T.p_len == 0) {
            const uint32_t keep = T.lit;
            o.from = 0; T = o.exit;
            if (o.n_spec == 0) T.lit = keep;
            return;
        }
        if (T.p_idx + T.p_len <= B) {
            // ... and a pending match that ends before the segment: the first candidate either replaces it (>= GOOD) or
            // pushes it out unchanged (Match::select); after that the two parses are in step.
            if (o.cand0 == kNoPos) { T.cur = se; return; }
            uint32_t keep = T.lit;
            if (!o.good0) { fix[0] = make_uint4(T.p_idx, T.p_len, T.p_idx - T.p_midx, 0); o.n_fix = 1; keep = T.p_idx + T.p_len; }
            o.from = 0; T = o.exit;
            if (o.n_spec == 0) T.lit = keep;
            return;
        }
    }
    if (T.cur < o.cand0) T.cur = o.cand0 < se ? o.cand0 : se;  // nothing happens before the segment's first candidate
    uint32_t j = 0, n_fix = 0, lim_unused = 0, cand_unused = 0, good_unused = 0;  // (what only a speculative replay records)
    const uint32_t n_cmp = o.n_spec < kSpecStates ? o.n_spec : kSpecStates;
    uint4 sj = n_cmp ? sp[0] : make_uint4(0, 0, 0, 0);
    uint32_t wreg = 0, wbase = kNoPos;
    while (T.cur < se) {
        Match sel;
        bool need_unused = false;
        uint32_t w;
        if (kWarp) {
            const uint32_t b = T.cur & ~31u;
            if (b != wbase) { wbase = b; wreg = W[b + lane]; }  // (may read up to 31 words past the stream's last position: scratch, never used)
            w = __shfl_sync(0xFFFFFFFFu, wreg, T.cur & 31u);
        } else {
            w = W[T.cur];
        }
        if (!front_step<false>(src, len, end, w, T, sel, lim_unused, cand_unused, good_unused, 0u, need_unused)) continue;
        fix[n_fix++] = make_uint4(sel.idx, sel.match_len, sel.idx - sel.match_idx, T.cur);
        while (j < n_cmp && sj.x + sj.y < T.lit) { j++; if (j < n_cmp) sj = sp[j]; }
        if (j < n_cmp && sj.x + sj.y == T.lit && sj.w == T.cur) {
            const uint4 ps = stt[j];
            if (ps.z == T.p_len && (T.p_len == 0 || (ps.x == T.p_idx && ps.y == T.p_midx))) {  // the same state: in step from here
                o.n_fix = n_fix; o.from = j + 1; T = o.exit;
                return;
            }
        }
    }
    o.n_fix = n_fix;
}
// a: every segment at once, ASSUMING the exit state of the segment before it is true (it is, once that segment is in step).
__global__ void k_long_stitch_a(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
                                const StreamCounts *__restrict__ bases, const LongSeg *__restrict__ rseg, uint32_t n_rseg, const uint32_t *__restrict__ words,
                                const uint4 *__restrict__ spec, const uint4 *__restrict__ states, uint4 *__restrict__ fix, LongSegOut *seg_out) {
    const uint32_t rs = blockIdx.x * blockDim.x + threadIdx.x;
    if (rs >= n_rseg) return;
    const LongSeg sg = rseg[rs];
    if (sg.k == 0) return;
    const uint32_t len = (uint32_t)src_len[sg.stream];
    LongSegOut o = seg_out[rs];
    FrontState T = seg_out[rs - 1].exit;
    stitch_segment<false>(src_base + src_off[sg.stream], len, len - 3, words + bases[sg.stream].n_fse, sg.k, T, o, spec + (size_t)rs * kEmitCap,
                   states + (size_t)rs * kSpecStates, fix + (size_t)rs * kEmitCap);
    seg_out[rs].n_fix = o.n_fix; seg_out[rs].from = o.from; seg_out[rs].a_out = T;  // (.exit is being read by the neighbour)
    seg_out[rs].a_ok = same_state(T, o.exit);
}
// b: one warp per stream walks the segments in order with the true state.  While the true state entering a segment is the exit
// state of the one before, what k_long_stitch_a left is right: the lanes look at 32 segments at a time and the walk jumps to
// the first that did not fall into step; its a_out is still the truth behind it.  Any other segment is stitched again from
// the true state (the warp in step, words fetched 32 at a time).  Leaves the stream's tail (pending match, final literals; frontend_bytes.rs:121-131,271-317).
__global__ void __launch_bounds__(128)
k_long_stitch_b(const uint8_t *__restrict__ src_base, const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
                const EncStream *__restrict__ streams, const StreamCounts *__restrict__ bases, const uint32_t *__restrict__ seg_list,
                uint32_t n_segl, const uint32_t *__restrict__ words, const uint4 *__restrict__ spec, const uint4 *__restrict__ states,
                uint4 *__restrict__ fix, LongSegOut *seg_out, uint4 *tail, uint32_t *redo_count) {
    const uint32_t slot = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (slot >= n_segl) return;
    const uint32_t si = seg_list[slot];
    const EncStream st = streams[si];
    const uint8_t *src = src_base + src_off[si];
    const uint32_t len = (uint32_t)src_len[si], end = len - 3;
    const uint32_t *W = words + bases[si].n_fse;
    FrontState T = seg_out[st.rseg_base].exit;
    uint32_t redo = 0, k = 1;
    while (k < st.n_rseg) {
        const uint32_t rs = st.rseg_base + k;
        if (same_state(T, seg_out[rs - 1].exit)) {
            bool ok = false;
            if (k + lane < st.n_rseg) ok = seg_out[rs + lane].a_ok != 0;
            const uint32_t nb = ~__ballot_sync(0xFFFFFFFFu, ok);
            const uint32_t m = nb ? (uint32_t)__ffs(nb) - 1u : 32u;  // segments k .. k+m-1 fell into step
            if (m == 32 || k + m >= st.n_rseg) { T = seg_out[rs + m - 1].exit; k += m; }
            else { T = seg_out[rs + m].a_out; k += m + 1; }
            continue;
        }
        {   // the whole warp in step (identical state in every lane; every lane stores the same values)
            LongSegOut o = seg_out[rs];
            stitch_segment<true>(src, len, end, W, k, T, o, spec + (size_t)rs * kEmitCap, states + (size_t)rs * kSpecStates, fix + (size_t)rs * kEmitCap, lane);
            __syncwarp();
            if (lane == 0) { seg_out[rs].n_fix = o.n_fix; seg_out[rs].from = o.from; }
        }
        redo++; k++;
    }
    if (lane == 0) {
        uint32_t lit = T.lit, n_tail = 0;
        if (T.p_len != 0) { tail[2 * slot + n_tail++] = make_uint4(T.p_idx, T.p_len, T.p_idx - T.p_midx, 0); lit = T.p_idx + T.p_len; }
        if (len - lit != 0) tail[2 * slot + n_tail++] = make_uint4(len, 0, 1, 0);  // push_literals: (L, 0, 1)
        if (n_tail < 2) tail[2 * slot + 1] = make_uint4(0, 0, 0, 0xFFFFFFFFu);
        if (n_tail < 1) tail[2 * slot] = make_uint4(0, 0, 0, 0xFFFFFFFFu);
        if (redo) atomicAdd(redo_count, redo);
    }
}

// ---- matches -> packs and block records, one WARP per stream ------------------------------------------
__device__ __forceinline__ void wsink_emit_block(TSink &s, uint64_t &out_used, const TEnv &env, uint32_t lane) {
    const uint32_t n_packs = s.n_packs_total - s.blk_pack0, n_lits = s.n_lits_total - s.blk_lit0;
    if (lane == 0) {
        EncBlock b;
        b.pack_off = env.base.n_blocks + s.blk_pack0;
        b.lit_off = env.base.n_fse + s.blk_lit0;
        b.out_off = env.base.n_lmds + out_used;
        b.src_pos = env.src_off + s.blk_src0;
        b.n_packs = n_packs; b.n_lits = n_lits; b.n_match_bytes = s.n_match_bytes;
        b.out_size = 0; b.gather = 1; b.pad = 0; b.dst_pos = 0;
        const uint32_t id = atomicAdd(env.block_counter, 1u);
        env.blocks[id] = b;
        env.block_ids[env.base.n_literals + s.n_blocks] = id;
    }
    out_used += (block_bound(n_lits, n_packs) + 15) & ~15ull;
    s.n_blocks++;
    s.blk_src0 += n_lits + s.n_match_bytes;
    s.blk_pack0 = s.n_packs_total; s.blk_lit0 = s.n_lits_total;
    s.n_match_bytes = 0; s.match_distance = 0;
}
// The conversion is sequential only at block borders, so it is split in three:
//   k_long_seg_stats   warp / segment: what the segment's matches add up to (count, literal and match bytes, all of them
//                      single packs?) 
//   k_long_blocks      warp / stream: walks the segments in order over those sums -- a segment of plain packs that stays
//                      inside the open block, or crosses into the next one because the block's 10 000 packs are full, costs
//                      a few instructions and only fixes where its packs go; anything else (L > 315, M > 2359, a block
//                      that fills up with literals) is pushed match by match right here (Buffer::push);
//   k_long_write_packs warp / segment: writes the packs of the plain segments.
struct SegAgg {
    uint32_t n;          // matches of the segment that are part of the true parse (fix list, then the speculative list from `from`)
    uint32_t sum_lit;    // literal bytes of matches 1..n-1 (the first one's depend on where the previous segment ended)
    uint32_t sum_m;      // match bytes
    uint32_t simple;     // every L (but the first) <= 315 and every M <= 2359
    uint32_t first_idx, last_end, last_dist;
    uint32_t n_fix, from;  // copied from the segment's LongSegOut: the block walk forms record addresses without another load
    uint32_t pad;
};
struct SegEntry { uint32_t pack_base, d_prev, split, first_lit; };  // split: first match of the segment that opens a new block (n: none; kNoPos: packs already written)

__device__ __forceinline__ uint4 seg_record(const uint4 *fixl, const uint4 *specl, uint32_t n_fix, uint32_t i) { return i < n_fix ? fixl[i] : specl[i - n_fix]; }

__global__ void __launch_bounds__(128)
k_long_seg_stats(const uint4 *__restrict__ spec, const uint4 *__restrict__ fix, const LongSegOut *__restrict__ seg_out, uint32_t n_rseg, SegAgg *agg,
                 uint2 *__restrict__ csum) {
    const uint32_t rs = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (rs >= n_rseg) return;
    const LongSegOut o = seg_out[rs];
    const uint4 *fixl = fix + (size_t)rs * kEmitCap, *specl = spec + (size_t)rs * kEmitCap + o.from;
    const uint32_t n = o.n_fix + (o.n_spec - o.from);
    uint32_t run_lit = 0, run_m = 0, carry_end = 0, simple = 1, first_idx = 0, last_end = 0, last_dist = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint4 r = make_uint4(0, 0, 0, 0);
        if (i < n) r = seg_record(fixl, specl, o.n_fix, i);
        const uint32_t my_end = r.x + r.y;
        uint32_t pe = __shfl_up_sync(0xFFFFFFFFu, my_end, 1);
        if (lane == 0) pe = carry_end;
        const uint32_t lit = (i < n && i != 0) ? r.x - pe : 0u;
        if (i < n && (lit > kMaxLValue || r.y > kMaxMValue)) simple = 0;
        const uint32_t cl = __reduce_add_sync(0xFFFFFFFFu, lit), cm = __reduce_add_sync(0xFFFFFFFFu, i < n ? r.y : 0u);
        if (lane == 0) csum[(size_t)rs * kSegChunks + (i0 >> 5)] = make_uint2(cl, cm);  // literal (without match 0's) and match bytes of this chunk
        run_lit += cl; run_m += cm;
        const uint32_t nn = n - i0 < 32 ? n - i0 : 32;
        carry_end = __shfl_sync(0xFFFFFFFFu, my_end, nn - 1);
        if (i0 == 0) first_idx = __shfl_sync(0xFFFFFFFFu, r.x, 0);
        last_end = carry_end; last_dist = __shfl_sync(0xFFFFFFFFu, r.z, nn - 1);
    }
    simple = __all_sync(0xFFFFFFFFu, simple != 0);
    if (lane == 0) agg[rs] = SegAgg{n, run_lit, run_m, simple, first_idx, last_end, last_dist, o.n_fix, o.from, 0};
}

__global__ void __launch_bounds__(128)
k_long_blocks(const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len, EncStream *streams, const StreamCounts *__restrict__ bases,
              const uint32_t *__restrict__ seg_list, uint32_t n_segl, const uint4 *__restrict__ spec, const uint4 *__restrict__ fix,
              const LongSegOut *__restrict__ seg_out, const SegAgg *__restrict__ agg, const uint2 *__restrict__ csum, SegEntry *entry,
              const uint4 *__restrict__ tail, uint2 *pack_scratch, uint32_t *block_ids, EncBlock *blocks, uint32_t *block_counter, uint32_t *exact_count) {
    const uint32_t slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (slot >= n_segl) return;
    const uint32_t si = seg_list[slot];
    const EncStream st = streams[si];
    TEnv env;
    env.base = bases[si]; env.blocks = blocks; env.block_ids = block_ids; env.block_counter = block_counter; env.src_off = src_off[si];
    TSink fs;
    fs.packs = pack_scratch + env.base.n_blocks;
    fs.n_packs_total = 0; fs.n_lits_total = 0; fs.blk_pack0 = 0; fs.blk_lit0 = 0; fs.n_match_bytes = 0; fs.match_distance = 0;
    fs.n_blocks = 0; fs.out_used = 0; fs.blk_src0 = 0;
    uint32_t prev_end = 0, n_exact = 0;
    uint64_t out_used = 0;  // (a 2 GiB stream's blocks pass 4 GiB of scratch)
    // Buffer::push for a list of matches, 32 per step while all of them are single packs that fit the open block
    auto run = [&](const uint4 *fixl, const uint4 *specl, uint32_t n_fix, uint32_t count) {
        for (uint32_t i0 = 0; i0 < count; i0 += 32) {
            const uint32_t n = count - i0 < 32 ? count - i0 : 32;
            uint4 r = make_uint4(0, 0, 0, 0);
            if (lane < n) r = seg_record(fixl, specl, n_fix, i0 + lane);
            const uint32_t my_end = r.x + r.y;
            uint32_t pe = __shfl_up_sync(0xFFFFFFFFu, my_end, 1);
            if (lane == 0) pe = prev_end;
            const uint32_t lit_len = lane < n ? r.x - pe : 0u;
            const bool simple = lit_len <= kMaxLValue && r.y <= kMaxMValue;
            const uint32_t sum_lit = __reduce_add_sync(0xFFFFFFFFu, lit_len);
            const bool room = fs.n_packs_total - fs.blk_pack0 + n <= kLmdsPerBlock && fs.n_lits_total - fs.blk_lit0 + sum_lit <= kLiteralsPerBlock;
            if (__all_sync(0xFFFFFFFFu, simple) && room) {
                uint32_t dp = __shfl_up_sync(0xFFFFFFFFu, r.z, 1);
                if (lane == 0) dp = fs.match_distance;
                if (lane < n) fs.packs[fs.n_packs_total + lane] = make_uint2(lit_len | (r.y << 16), r.z == dp ? 0u : r.z);
                fs.n_packs_total += n; fs.n_lits_total += sum_lit;
                fs.n_match_bytes += __reduce_add_sync(0xFFFFFFFFu, lane < n ? r.y : 0u);
                fs.match_distance = __shfl_sync(0xFFFFFFFFu, r.z, n - 1);
            } else {
                for (uint32_t i = 0; i < n; i++) {  // one match at a time, the warp in step (every lane stores the same values)
                    uint32_t l = __shfl_sync(0xFFFFFFFFu, lit_len, i), m = __shfl_sync(0xFFFFFFFFu, r.y, i);
                    const uint32_t d = __shfl_sync(0xFFFFFFFFu, r.z, i);
                    if (l <= kMaxLValue && m <= kMaxMValue && fs.n_packs_total - fs.blk_pack0 < kLmdsPerBlock &&
                        fs.n_lits_total - fs.blk_lit0 + l <= kLiteralsPerBlock) {
                        fs.n_lits_total += l;
                        tsink_push_lmd(fs, l, m, d);
                    } else {
                        while (!tsink_buffer_push(fs, l, m, d)) wsink_emit_block(fs, out_used, env, lane);
                    }
                }
            }
            prev_end = __shfl_sync(0xFFFFFFFFu, my_end, n - 1);
        }
    };
    for (uint32_t k0 = 0; k0 < st.n_rseg; k0 += 32) {
        SegAgg mine = SegAgg{0, 0, 0, 1, 0, 0, 0, 0, 0, 0};
        if (k0 + lane < st.n_rseg) mine = agg[st.rseg_base + k0 + lane];
        const uint32_t kn = st.n_rseg - k0 < 32 ? st.n_rseg - k0 : 32;
        for (uint32_t t = 0; t < kn; t++) {
            {   // ---- all segments from t on that stay inside the open block as plain packs, in one go ----
                // Lane l looks at segment l: where the previous non-empty segment ended (its first literal run), running
                // sums over the lanes, first lane that does not fit.  What follows handles that one segment in order.
                const bool in = lane >= t && lane < kn;
                const uint32_t ne = __ballot_sync(0xFFFFFFFFu, in && mine.n != 0);
                const uint32_t before = ne & ((1u << lane) - 1u);
                const int src_l = before ? 31 - __clz(before) : (int)lane;
                const uint32_t pe_s = __shfl_sync(0xFFFFFFFFu, mine.last_end, src_l), dp_s = __shfl_sync(0xFFFFFFFFu, mine.last_dist, src_l);
                const uint32_t pe = before ? pe_s : prev_end, dp = before ? dp_s : fs.match_distance;
                const bool has = in && mine.n != 0;
                const uint32_t fl = has ? mine.first_idx - pe : 0u;
                const bool okl = !has || (mine.simple && fl <= kMaxLValue);
                uint32_t cn = has ? mine.n : 0u, cl = has ? fl + mine.sum_lit : 0u, cm = has ? mine.sum_m : 0u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, cn, o), b = __shfl_up_sync(0xFFFFFFFFu, cl, o), c = __shfl_up_sync(0xFFFFFFFFu, cm, o);
                    if (lane >= (uint32_t)o) { cn += a; cl += b; cm += c; }
                }
                const uint32_t cnt0 = fs.n_packs_total - fs.blk_pack0, lits0 = fs.n_lits_total - fs.blk_lit0;
                // (sums of 32 segments stay far below 2^32: a segment has at most kEmitCap matches; its literal bytes may be large
                // when a pending match lies far back, hence the saturating test on cl through the per-lane flag below)
                const bool fit = okl && cnt0 + cn <= kLmdsPerBlock && cl <= kLiteralsPerBlock && lits0 + cl <= kLiteralsPerBlock;
                const uint32_t bad = __ballot_sync(0xFFFFFFFFu, in && !fit);
                const uint32_t t_stop = bad ? (uint32_t)__ffs(bad) - 1u : kn;  // segments [t, t_stop) are plain
                if (t_stop > t) {
                    if (has && lane < t_stop) entry[st.rseg_base + k0 + lane] = SegEntry{fs.n_packs_total + cn - mine.n, dp, mine.n, fl};
                    const uint32_t last = t_stop - 1;
                    const uint32_t tn = __shfl_sync(0xFFFFFFFFu, cn, last), tl = __shfl_sync(0xFFFFFFFFu, cl, last), tm = __shfl_sync(0xFFFFFFFFu, cm, last);
                    const uint32_t ne_run = ne & ((2u << last) - 1u);
                    if (ne_run) {
                        const int ll = 31 - __clz(ne_run);
                        fs.match_distance = __shfl_sync(0xFFFFFFFFu, mine.last_dist, ll); prev_end = __shfl_sync(0xFFFFFFFFu, mine.last_end, ll);
                    }
                    fs.n_packs_total += tn; fs.n_lits_total += tl; fs.n_match_bytes += tm;
                    t = t_stop;
                    if (t >= kn) break;
                }
            }
            const uint32_t rs = st.rseg_base + k0 + t;
            const uint32_t n = __shfl_sync(0xFFFFFFFFu, mine.n, t);
            if (n == 0) continue;
            const uint32_t sum_lit = __shfl_sync(0xFFFFFFFFu, mine.sum_lit, t), sum_m = __shfl_sync(0xFFFFFFFFu, mine.sum_m, t);
            const uint32_t simple = __shfl_sync(0xFFFFFFFFu, mine.simple, t), first_idx = __shfl_sync(0xFFFFFFFFu, mine.first_idx, t);
            const uint32_t last_end = __shfl_sync(0xFFFFFFFFu, mine.last_end, t), last_dist = __shfl_sync(0xFFFFFFFFu, mine.last_dist, t);
            const uint32_t first_lit = first_idx - prev_end;
            const uint32_t cnt = fs.n_packs_total - fs.blk_pack0, lits = fs.n_lits_total - fs.blk_lit0;
            bool done = false;
            if (simple && first_lit <= kMaxLValue) {
                if (cnt + n <= kLmdsPerBlock && lits + first_lit + sum_lit <= kLiteralsPerBlock) {
                    if (lane == 0) entry[rs] = SegEntry{fs.n_packs_total, fs.match_distance, n, first_lit};
                    fs.n_packs_total += n; fs.n_lits_total += first_lit + sum_lit; fs.n_match_bytes += sum_m;
                    done = true;
                } else if (cnt + n > kLmdsPerBlock) {
                    // the open block takes its 10 000th pack inside this segment: matches [0, j) finish it, match j opens the next one
                    const uint32_t j = kLmdsPerBlock - cnt;
                    uint2 pj = make_uint2(0, 0);  // literal bytes (without the first match's) and match bytes of matches [0, j)
                    uint32_t lit_a = 0;           // literal bytes of matches [0, j)
                    if (j != 0) {
                        const uint32_t o_n_fix = __shfl_sync(0xFFFFFFFFu, mine.n_fix, t), o_from = __shfl_sync(0xFFFFFFFFu, mine.from, t);
                        const uint4 *fixl = fix + (size_t)rs * kEmitCap, *specl = spec + (size_t)rs * kEmitCap + o_from;
                        // whole chunks from k_long_seg_stats' sums, the rest from the records of the chunk j lies in
                        uint32_t sl = 0, sm = 0;
                        const uint32_t cf = j >> 5, i = (cf << 5) + lane;
                        for (uint32_t c = lane; c < cf; c += 32) { const uint2 v = csum[(size_t)rs * kSegChunks + c]; sl += v.x; sm += v.y; }
                        uint4 r = make_uint4(0, 0, 0, 0), rp = make_uint4(0, 0, 0, 0);
                        if (i < j) r = seg_record(fixl, specl, o_n_fix, i);
                        if (lane == 0 && i != 0 && i < j) rp = seg_record(fixl, specl, o_n_fix, i - 1);
                        uint32_t pe = __shfl_up_sync(0xFFFFFFFFu, r.x + r.y, 1);
                        if (lane == 0) pe = rp.x + rp.y;
                        if (i < j && i != 0) sl += r.x - pe;
                        if (i < j) sm += r.y;
                        pj.x = __reduce_add_sync(0xFFFFFFFFu, sl); pj.y = __reduce_add_sync(0xFFFFFFFFu, sm);
                        lit_a = first_lit + pj.x;
                    }
                    const uint32_t lit_b = first_lit + sum_lit - lit_a;
                    if (lits + lit_a <= kLiteralsPerBlock && lit_b <= kLiteralsPerBlock && n - j <= kLmdsPerBlock) {
                        if (lane == 0) entry[rs] = SegEntry{fs.n_packs_total, fs.match_distance, j, first_lit};
                        fs.n_packs_total += j; fs.n_lits_total += lit_a; fs.n_match_bytes += pj.y;
                        wsink_emit_block(fs, out_used, env, lane);
                        fs.n_packs_total += n - j; fs.n_lits_total += lit_b; fs.n_match_bytes += sum_m - pj.y;
                        done = true;
                    }
                }
            }
            if (done) { fs.match_distance = last_dist; prev_end = last_end; continue; }
            const LongSegOut o = seg_out[rs];
            if (lane == 0) entry[rs] = SegEntry{0, 0, kNoPos, 0};
            run(fix + (size_t)rs * kEmitCap, spec + (size_t)rs * kEmitCap + o.from, o.n_fix, n);
            n_exact++;
        }
    }
    for (uint32_t t = 0; t < 2; t++) {
        const uint4 r = tail[2 * slot + t];
        if (r.w != 0xFFFFFFFFu) run(tail + 2 * slot + t, nullptr, 1, 1);
    }
    wsink_emit_block(fs, out_used, env, lane);  // finalize
    if (lane == 0) { streams[si].n_blocks = fs.n_blocks; if (n_exact) atomicAdd(exact_count, n_exact); }
}

__global__ void __launch_bounds__(128)
k_long_write_packs(const StreamCounts *__restrict__ bases, const LongSeg *__restrict__ rseg, uint32_t n_rseg, const uint4 *__restrict__ spec,
                   const uint4 *__restrict__ fix, const LongSegOut *__restrict__ seg_out, const SegAgg *__restrict__ agg, const SegEntry *__restrict__ entry,
                   uint2 *pack_scratch) {
    const uint32_t rs = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (rs >= n_rseg) return;
    const uint32_t n = agg[rs].n;
    if (n == 0) return;
    const SegEntry en = entry[rs];
    if (en.split == kNoPos) return;
    const LongSegOut o = seg_out[rs];
    const uint4 *fixl = fix + (size_t)rs * kEmitCap, *specl = spec + (size_t)rs * kEmitCap + o.from;
    uint2 *packs = pack_scratch + bases[rseg[rs].stream].n_blocks + en.pack_base;
    uint32_t carry_end = 0, carry_d = en.d_prev;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint4 r = make_uint4(0, 0, 0, 0);
        if (i < n) r = seg_record(fixl, specl, o.n_fix, i);
        const uint32_t my_end = r.x + r.y;
        uint32_t pe = __shfl_up_sync(0xFFFFFFFFu, my_end, 1), dp = __shfl_up_sync(0xFFFFFFFFu, r.z, 1);
        if (lane == 0) { pe = carry_end; dp = carry_d; }
        if (i == en.split) dp = 0;  // first pack of a block: nothing to repeat (Buffer::reset)
        const uint32_t lit = i == 0 ? en.first_lit : r.x - pe;
        if (i < n) packs[i] = make_uint2(lit | (r.y << 16), r.z == dp ? 0u : r.z);
        const uint32_t nn = n - i0 < 32 ? n - i0 : 32;
        carry_end = __shfl_sync(0xFFFFFFFFu, my_end, nn - 1); carry_d = __shfl_sync(0xFFFFFFFFu, r.z, nn - 1);
    }
}

// ---- the blocks of long streams are copied into the frame side by side -------------------------------------
__global__ void __launch_bounds__(256)
k_long_copy(const EncBlock *__restrict__ blocks, const uint32_t *__restrict__ n_blocks_p, const uint8_t *__restrict__ out_scratch, uint8_t *__restrict__ dst_base) {
    const uint32_t n_blocks = *n_blocks_p;
    for (uint32_t bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
        const EncBlock b = blocks[bi];
        if (b.pad != 1) continue;
        const uint8_t *s = out_scratch + b.out_off;
        uint8_t *d = dst_base + b.dst_pos;
        const uint32_t n = b.out_size;
        const uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15);
        if (n >= head + 16) {
            // the scratch side is 16-byte aligned; the frame side is reached through aligned 16-byte stores assembled from
            // two neighbouring scratch words when the two are skewed
            for (uint32_t t = threadIdx.x; t < head; t += blockDim.x) d[t] = s[t];
            const uint32_t nv = (n - head) / 16;
            if (((reinterpret_cast<uintptr_t>(s) + head) & 15) == 0) {
                const uint4 *s4 = reinterpret_cast<const uint4 *>(s + head);
                uint4 *d4 = reinterpret_cast<uint4 *>(d + head);
                for (uint32_t t = threadIdx.x; t < nv; t += blockDim.x) d4[t] = s4[t];
            } else {
                uint32_t *d32 = reinterpret_cast<uint32_t *>(d + head);
                for (uint32_t t = threadIdx.x; t < nv * 4; t += blockDim.x) d32[t] = ld4u(s + head + t * 4);
            }
            for (uint32_t t = head + nv * 16 + threadIdx.x; t < n; t += blockDim.x) d[t] = s[t];
        } else {
            for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) d[t] = s[t];
        }
    }
}
