/*
 * lzfse_oracle.c -- CPU oracle for the LZFSE hot path.  TEST INFRASTRUCTURE ONLY (see header).
 *
 * Plain-C restatement of lzfse_rust v0.2.0's `decode_bytes` / `encode_bytes`.  Every function
 * cites the reference file:line it follows (paths relative to /root/reference/src).  Nothing in
 * here is used by the shipped CUDA library.
 */
#include "lzfse_oracle.h"

#include <immintrin.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Constants: fse/constants.rs:22-69, base/magic_bytes.rs:3-7, vn/constants.rs:1-13,
 * encode/constants.rs:3-10, encode/history.rs:10-13
 * ------------------------------------------------------------------------------------------ */
#define LMDS_PER_BLOCK 10000u
#define LITERALS_PER_BLOCK 40000u
#define L_SYMBOLS 20
#define M_SYMBOLS 20
#define D_SYMBOLS 64
#define U_SYMBOLS 256
#define L_STATES 64u
#define M_STATES 64u
#define D_STATES 256u
#define U_STATES 1024u
#define MAX_L_VALUE 315u
#define MAX_M_VALUE 2359u
#define MAX_D_VALUE 262139u
#define N_WEIGHTS 360
#define V1_HEADER_SIZE 50u
#define V2_HEADER_SIZE 32u
#define V1_WEIGHT_PAYLOAD_BYTES 722u
#define V2_WEIGHT_PAYLOAD_BYTES_MAX 630u
#define MAX_L_BITS 14u
#define MAX_M_BITS 17u
#define MAX_D_BITS 23u
#define MAX_U_BITS 10u

#define BM_EOS 0x24787662u
#define BM_RAW 0x2D787662u
#define BM_VX1 0x31787662u
#define BM_VX2 0x32787662u
#define BM_VXN 0x6E787662u

#define VN_HEADER_SIZE 12u
#define VN_PAYLOAD_LIMIT 0x2000u
#define VN_MAX_D 65535u

#define GOOD_MATCH_LEN 40u
#define RAW_CUTOFF 20u
#define RAW_LIMIT 0x4000u
#define VN_CUTOFF 4096u
#define HASH_BITS 14
#define HASH_WIDTH 4
#define Q1 0x40000000u

static const uint8_t L_EXTRA_BITS[L_SYMBOLS] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 3, 5, 8};
static const uint8_t M_EXTRA_BITS[M_SYMBOLS] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 5, 8, 11};
/* D_EXTRA_BITS[i] == i / 4 (fse/constants.rs:305-311). */

/* Base values are the running sums of (1 << extra_bits) (fse/constants.rs:132-134,164-166,313-321);
 * *_BASE_FROM_VALUE / d_index (:136-157,168-303,323-353) are their inverses: the largest symbol
 * whose base is <= value.  Built once at load instead of restating ~3000 table entries. */
static uint32_t L_BASE_VALUE[L_SYMBOLS], M_BASE_VALUE[M_SYMBOLS], D_BASE_VALUE[D_SYMBOLS];
static uint8_t D_EXTRA_BITS[D_SYMBOLS];
static uint8_t L_SYM_FROM_VALUE[MAX_L_VALUE + 1], M_SYM_FROM_VALUE[MAX_M_VALUE + 1];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void init_tables_once(void) {
    uint32_t b = 0;
    for (int i = 0; i < L_SYMBOLS; i++) { L_BASE_VALUE[i] = b; b += 1u << L_EXTRA_BITS[i]; }
    b = 0;
    for (int i = 0; i < M_SYMBOLS; i++) { M_BASE_VALUE[i] = b; b += 1u << M_EXTRA_BITS[i]; }
    b = 0;
    for (int i = 0; i < D_SYMBOLS; i++) { D_EXTRA_BITS[i] = (uint8_t)(i / 4); D_BASE_VALUE[i] = b; b += 1u << D_EXTRA_BITS[i]; }
    for (uint32_t v = 0, s = 0; v <= MAX_L_VALUE; v++) { while (s + 1 < L_SYMBOLS && L_BASE_VALUE[s + 1] <= v) s++; L_SYM_FROM_VALUE[v] = (uint8_t)s; }
    for (uint32_t v = 0, s = 0; v <= MAX_M_VALUE; v++) { while (s + 1 < M_SYMBOLS && M_BASE_VALUE[s + 1] <= v) s++; M_SYM_FROM_VALUE[v] = (uint8_t)s; }
}
static void init_tables(void) { pthread_once(&g_once, init_tables_once); }

static inline uint32_t d_sym_from_value(uint32_t v) {
    /* Largest symbol with base <= v (fse/constants.rs:305-321 via d_index): four symbols per power of two,
     * base(4e + r) = ((4 + r) << e) - 4, so e = floor(log2(v + 4)) - 2 and r = ((v + 4) >> e) - 4. */
    if (v < 4) return v;
    const uint32_t e = 29u - (uint32_t)__builtin_clz(v + 4);
    return 4 * e + (((v + 4) >> e) - 4);
}

static inline uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint64_t le64(const uint8_t *p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }
static inline void st16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static inline void st32(uint8_t *p, uint32_t v) { st16(p, (uint16_t)v); st16(p + 2, (uint16_t)(v >> 16)); }
static inline void st64(uint8_t *p, uint64_t v) { st32(p, (uint32_t)v); st32(p + 4, (uint32_t)(v >> 32)); }
static inline uint64_t get_bits(uint64_t v, unsigned off, unsigned n) { return (v >> off) & ((n == 64) ? ~0ull : ((1ull << n) - 1)); }
static inline uint32_t clz32(uint32_t v) { return v ? (uint32_t)__builtin_clz(v) : 32u; }

/* ==========================================================================================
 * DECODE
 * ========================================================================================== */

/* bits/bit_reader.rs:20-71 over bits/bit_src.rs:35-46 (reads below index 0 yield 0). */
typedef struct { uint64_t accum; int64_t accum_bits; int64_t idx; const uint8_t *p; } bitreader;

static inline uint64_t br_read(const bitreader *r, int64_t idx) { return idx >= 0 ? le64(r->p + idx) : 0; }
static int br_new(bitreader *r, const uint8_t *p, size_t len, uint32_t off) {
    r->p = p; r->idx = (int64_t)len - 8; r->accum = br_read(r, r->idx); r->accum_bits = 64 - (int64_t)off;
    if (off != 0 && (r->accum >> r->accum_bits) != 0) return ORC_BAD_BITSTREAM;
    return ORC_OK;
}
static inline void br_flush(bitreader *r) {
    int64_t n_bytes = (64 - r->accum_bits) / 8;
    r->idx -= n_bytes; r->accum = br_read(r, r->idx); r->accum_bits += n_bytes * 8;
}
static inline uint64_t br_pull(bitreader *r, uint32_t n) {
    r->accum_bits -= n;
    uint64_t s = r->accum >> (r->accum_bits & 63);
    return s & ((1ull << n) - 1);
}
static int br_finalize(bitreader *r) {
    br_flush(r);
    if (r->accum_bits + r->idx * 8 < 64) return ORC_PAYLOAD_UNDERFLOW;
    return ORC_OK;
}

typedef struct { uint8_t k, v_bits; int16_t delta; uint32_t v_base; } ventry; /* fse/decoder.rs:205-212 */
typedef struct { uint8_t k, symbol; int16_t delta; } uentry;                  /* fse/decoder.rs:222-228 */

typedef struct { uint32_t num, n_payload_bytes, bits; uint16_t state[4]; } lit_param;
typedef struct { uint32_t num, n_payload_bytes, bits; uint16_t state[3]; } lmd_param;
typedef struct { lit_param literal; lmd_param lmd; uint32_t n_raw_bytes; } fse_block;

typedef struct {
    ventry v[L_STATES + M_STATES + D_STATES];
    uentry u[U_STATES];
    uint16_t weights[N_WEIGHTS];
    uint8_t literals[LITERALS_PER_BLOCK + 512];
    fse_block block;
} fse_core;

/* fse/block.rs:267-283 */
static int lmd_param_validate(const lmd_param *p) {
    uint32_t limit = 1024 + 8 + (p->num * MAX_L_BITS + p->num * MAX_M_BITS + p->num * MAX_D_BITS + 7) / 8;
    if (p->num > LMDS_PER_BLOCK || p->n_payload_bytes < 8 || p->n_payload_bytes > limit) return ORC_FSE_BAD_LMD_COUNT;
    if (p->bits > 7) return ORC_FSE_BAD_LMD_BITS;
    if (p->state[0] >= L_STATES || p->state[1] >= M_STATES || p->state[2] >= D_STATES) return ORC_FSE_BAD_LMD_STATE;
    return ORC_OK;
}
/* fse/block.rs:324-341 */
static int lit_param_validate(const lit_param *p) {
    if (p->num % 4 != 0 || p->num > LITERALS_PER_BLOCK) return ORC_FSE_BAD_LITERAL_COUNT;
    if (p->n_payload_bytes > 1024 + (p->num * MAX_U_BITS + 7) / 8) return ORC_FSE_BAD_LITERAL_COUNT;
    if (p->bits > 7) return ORC_FSE_BAD_LITERAL_BITS;
    if (p->state[0] >= U_STATES || p->state[1] >= U_STATES || p->state[2] >= U_STATES || p->state[3] >= U_STATES)
        return ORC_FSE_BAD_LMD_PAYLOAD; /* sic */
    return ORC_OK;
}
/* fse/block.rs:218-227 */
static int fse_block_validate(const fse_block *b) {
    int e;
    if ((e = lmd_param_validate(&b->lmd))) return e;
    if ((e = lit_param_validate(&b->literal))) return e;
    if (b->n_raw_bytes > b->literal.num + b->lmd.num * MAX_M_VALUE) return ORC_FSE_BAD_RAW_BYTE_COUNT;
    return ORC_OK;
}
/* fse/block.rs:80-104 */
static int fse_block_load_v1(fse_block *b, const uint8_t *s) {
    b->n_raw_bytes = le32(s + 4);
    uint32_t n_payload_bytes = le32(s + 8);
    b->literal.num = le32(s + 12);
    b->lmd.num = le32(s + 16);
    b->literal.n_payload_bytes = le32(s + 20);
    b->lmd.n_payload_bytes = le32(s + 24);
    b->literal.bits = 0u - le32(s + 28);
    for (int i = 0; i < 4; i++) b->literal.state[i] = le16(s + 32 + 2 * i);
    b->lmd.bits = 0u - le32(s + 40);
    for (int i = 0; i < 3; i++) b->lmd.state[i] = le16(s + 44 + 2 * i);
    if (n_payload_bytes < b->literal.n_payload_bytes + b->lmd.n_payload_bytes) return ORC_FSE_BAD_PAYLOAD_COUNT;
    return fse_block_validate(b);
}
/* fse/block.rs:108-136 */
static int fse_block_load_v2(fse_block *b, const uint8_t *s, uint32_t *n_weight_payload_bytes) {
    b->n_raw_bytes = le32(s + 4);
    uint64_t p = le64(s + 8);
    b->literal.num = (uint32_t)get_bits(p, 0, 20);
    b->literal.n_payload_bytes = (uint32_t)get_bits(p, 20, 20);
    b->lmd.num = (uint32_t)get_bits(p, 40, 20);
    b->literal.bits = 7 - (uint32_t)get_bits(p, 60, 3);
    p = le64(s + 16);
    for (int i = 0; i < 4; i++) b->literal.state[i] = (uint16_t)get_bits(p, 10 * i, 10);
    b->lmd.n_payload_bytes = (uint32_t)get_bits(p, 40, 20);
    b->lmd.bits = 7 - (uint32_t)get_bits(p, 60, 3);
    p = le64(s + 24);
    uint32_t header_size = (uint32_t)get_bits(p, 0, 32);
    b->lmd.state[0] = (uint16_t)get_bits(p, 32, 10);
    b->lmd.state[1] = (uint16_t)get_bits(p, 42, 10);
    b->lmd.state[2] = (uint16_t)get_bits(p, 52, 10);
    *n_weight_payload_bytes = header_size - V2_HEADER_SIZE;
    if (*n_weight_payload_bytes > V2_WEIGHT_PAYLOAD_BYTES_MAX) return ORC_FSE_BAD_WEIGHT_PAYLOAD;
    return fse_block_validate(b);
}

/* fse/weights.rs:189-201 */
static int weights_check_totals(uint16_t *w) {
    uint32_t tl = 0, tm = 0, td = 0, tu = 0;
    for (int i = 0; i < 20; i++) tl += w[i];
    for (int i = 20; i < 40; i++) tm += w[i];
    for (int i = 40; i < 104; i++) td += w[i];
    for (int i = 104; i < 360; i++) tu += w[i];
    if (tl <= L_STATES && tm <= M_STATES && td <= D_STATES && tu <= U_STATES) return ORC_OK;
    memset(w, 0, N_WEIGHTS * sizeof(uint16_t));
    return ORC_FSE_BAD_WEIGHT_PAYLOAD;
}
/* fse/weights.rs:66-80 */
static int weights_load_v1(uint16_t *w, const uint8_t *src) {
    for (int i = 0; i < N_WEIGHTS; i++) w[i] = le16(src + 2 * i);
    return weights_check_totals(w);
}
/* fse/constants.rs:115-124, fse/weight_encoder.rs:10-20 */
static const uint8_t WEIGHTS_BITS_TABLE[32] = {2, 3, 2, 5, 2, 3, 2, 8, 2, 3, 2, 5, 2, 3, 2, 14, 2, 3, 2, 5, 2, 3, 2, 8, 2, 3, 2, 5, 2, 3, 2, 14};
static const int8_t WEIGHTS_VALUE_TABLE[32] = {0, 2, 1, 4, 0, 3, 1, -1, 0, 2, 1, 5, 0, 3, 1, -1, 0, 2, 1, 6, 0, 3, 1, -1, 0, 2, 1, 7, 0, 3, 1, -1};
/* fse/weights.rs:83-105 */
static int weights_load_v2(uint16_t *w, const uint8_t *src, size_t len) {
    uint64_t accum = 0; int64_t accum_bits = 0; size_t i = 0;
    for (int n = 0; n < N_WEIGHTS; n++) {
        while (i != len && accum_bits <= 24) { accum |= (uint64_t)src[i] << accum_bits; accum_bits += 8; i++; }
        uint32_t index = (uint32_t)(accum & 0x1F);
        uint32_t bits = WEIGHTS_BITS_TABLE[index], v;
        if (bits == 8) v = 8 + (uint32_t)((accum >> 4) & 0xF);
        else if (bits == 14) v = 24 + (uint32_t)((accum >> 4) & 0x3FF);
        else v = (uint32_t)WEIGHTS_VALUE_TABLE[index];
        w[n] = (uint16_t)v;
        accum >>= bits; accum_bits -= bits;
    }
    if (accum_bits < 0) return ORC_FSE_WEIGHT_PAYLOAD_UNDERFLOW;
    if (accum_bits >= 8 || i != len) return ORC_FSE_WEIGHT_PAYLOAD_OVERFLOW;
    return weights_check_totals(w);
}

/* fse/decoder.rs:244-292 */
static void build_v_table_block(const uint16_t *weights, int n_sym, const uint8_t *bits_t, const uint32_t *base_t,
                                ventry *table, uint32_t n_states, int16_t offset) {
    uint32_t n_clz = clz32(n_states), total = 0;
    for (int i = 0; i < n_sym; i++) {
        uint32_t w = weights[i];
        if (w == 0) continue;
        uint32_t k = clz32(w) - n_clz;
        uint32_t x = ((n_states << 1) >> k) - w;
        ventry e; e.v_bits = bits_t[i]; e.v_base = base_t[i]; e.k = (uint8_t)k;
        for (uint32_t j = 0; j < x; j++) { e.delta = (int16_t)((int16_t)((((int32_t)w + (int32_t)j) << k) - (int32_t)n_states) + offset); table[total + j] = e; }
        e.k = (uint8_t)((int32_t)k - 1);
        for (uint32_t j = x; j < w; j++) { e.delta = (int16_t)((int16_t)((j - x) << (k - 1)) + offset); table[total + j] = e; }
        total += w;
    }
    for (uint32_t i = total; i < n_states; i++) { ventry e = {0, 0, (int16_t)(offset + (int16_t)i), 0}; table[i] = e; }
}
/* fse/decoder.rs:299-335 */
static void build_u_table(const uint16_t *weights, uentry *table) {
    uint32_t n_states = U_STATES, n_clz = clz32(n_states), total = 0;
    for (int i = 0; i < U_SYMBOLS; i++) {
        uint32_t w = weights[i];
        if (w == 0) continue;
        uint32_t k = clz32(w) - n_clz;
        uint32_t x = ((n_states << 1) >> k) - w;
        uentry e; e.symbol = (uint8_t)i; e.k = (uint8_t)k;
        for (uint32_t j = 0; j < x; j++) { e.delta = (int16_t)((((int32_t)w + (int32_t)j) << k) - (int32_t)n_states); table[total + j] = e; }
        e.k = (uint8_t)((int32_t)k - 1);
        for (uint32_t j = x; j < w; j++) { e.delta = (int16_t)((j - x) << (k - 1)); table[total + j] = e; }
        total += w;
    }
    for (uint32_t i = total; i < n_states; i++) { uentry e = {0, 0, (int16_t)i}; table[i] = e; }
}
/* fse/decoder.rs:21-67 */
static void decoder_init(fse_core *c) {
    build_v_table_block(c->weights, L_SYMBOLS, L_EXTRA_BITS, L_BASE_VALUE, c->v, L_STATES, 0);
    build_v_table_block(c->weights + 20, M_SYMBOLS, M_EXTRA_BITS, M_BASE_VALUE, c->v + 64, M_STATES, 64);
    build_v_table_block(c->weights + 40, D_SYMBOLS, D_EXTRA_BITS, D_BASE_VALUE, c->v + 128, D_STATES, 128);
    build_u_table(c->weights + 104, c->u);
}

typedef struct {
    const uint8_t *src; size_t src_len, pos;
    uint8_t *dst; size_t dst_cap, out;
    orc_lmd_t *trace; size_t trace_cap, trace_len;
} dctx;

#define NO_L 0xFFFFFFFFu
static inline void trace_push(dctx *d, uint32_t l, uint32_t m, uint32_t dist) {
    if (d->trace && d->trace_len < d->trace_cap) { orc_lmd_t t = {l, m, dist}; d->trace[d->trace_len] = t; }
    d->trace_len++;
}
/* lz/writer.rs:102-128 (Vec<u8> append); BufferOverflow is the C-ABI's fixed-capacity analogue. */
static inline int out_bytes(dctx *d, const uint8_t *p, size_t n) {
    if (n > d->dst_cap - d->out) return ORC_BUFFER_OVERFLOW;
    memcpy(d->dst + d->out, p, n); d->out += n;
    return ORC_OK;
}
/* Same, for sources with >= 16 bytes of slack (the FSE literal buffer): over-copy like lz/writer.rs:115-128. */
static inline int out_bytes_slack(dctx *d, const uint8_t *p, size_t n) {
    if (n > d->dst_cap - d->out) return ORC_BUFFER_OVERFLOW;
    uint8_t *q = d->dst + d->out;
    if (n <= 16 && d->dst_cap - d->out >= 16) memcpy(q, p, 16); else memcpy(q, p, n);
    d->out += n;
    return ORC_OK;
}
/* lz/writer.rs:144-180: byte i of the match equals byte i-distance. */
static inline int out_match(dctx *d, uint32_t len, uint32_t distance) {
    if ((size_t)distance > d->out || distance == 0) return ORC_BAD_D_VALUE;
    if (len > d->dst_cap - d->out) return ORC_BUFFER_OVERFLOW;
    uint8_t *q = d->dst + d->out; const uint8_t *s = q - distance;
    uint32_t i = 0;
    if (distance >= 16 && (size_t)len + 16 <= d->dst_cap - d->out) { /* 16-byte strides, over-copying into the slack like lz/writer.rs:156-177 */
        for (; i < len; i += 16) memcpy(q + i, s + i, 16);
        d->out += len;
        return ORC_OK;
    }
    if (distance >= 8) { /* 8-byte strides never read a byte this copy has not written yet (lz/object.rs:27-58) */
        for (; i + 8 <= len; i += 8) memcpy(q + i, s + i, 8);
    }
    for (; i < len; i++) q[i] = s[i];
    d->out += len;
    return ORC_OK;
}

/* fse/literals.rs:49-91 */
static int literals_load(fse_core *c, const uint8_t *p, size_t len) {
    const lit_param *prm = &c->block.literal;
    bitreader r; int e;
    if ((e = br_new(&r, p, len, prm->bits))) return e;
    uint32_t st[4] = {prm->state[0], prm->state[1], prm->state[2], prm->state[3]};
    for (uint32_t i = 0; i != prm->num; i += 4) {
        for (int s = 0; s < 4; s++) {
            uentry en = c->u[st[s]];
            st[s] = (uint32_t)((int64_t)br_pull(&r, en.k) + en.delta);
            c->literals[i + s] = en.symbol;
        }
        br_flush(&r);
    }
    if ((e = br_finalize(&r))) return e;
    if (st[0] | st[1] | st[2] | st[3]) return ORC_FSE_BAD_LMD_PAYLOAD;
    return ORC_OK;
}
/* fse/fse_core.rs:91-141 */
static int fse_decode_internal(fse_core *c, dctx *d, const uint8_t *p, size_t len) {
    const lmd_param *prm = &c->block.lmd;
    bitreader r; int e;
    if ((e = br_new(&r, p, len, prm->bits))) return e;
    uint32_t sl = prm->state[0], sm = 64u + prm->state[1], sd = 128u + prm->state[2];
    uint32_t literal_index = 0, n_match_bytes = 0, match_distance = 0;
    for (uint32_t n = prm->num; n != 0; n--) {
        ventry en = c->v[sl];
        sl = (uint32_t)((int64_t)br_pull(&r, en.k) + en.delta);
        uint32_t literal_len = en.v_base + (uint32_t)br_pull(&r, en.v_bits);
        en = c->v[sm];
        sm = (uint32_t)((int64_t)br_pull(&r, en.k) + en.delta);
        uint32_t match_len = en.v_base + (uint32_t)br_pull(&r, en.v_bits);
        en = c->v[sd];
        sd = (uint32_t)((int64_t)br_pull(&r, en.k) + en.delta);
        uint32_t dpack = en.v_base + (uint32_t)br_pull(&r, en.v_bits);
        br_flush(&r);
        if (dpack != 0) match_distance = dpack; /* lmd/lmd_type.rs:155-159 */
        const uint8_t *lit = c->literals + literal_index;
        literal_index += literal_len;
        if (literal_index > LITERALS_PER_BLOCK) return ORC_FSE_BAD_LMD_PAYLOAD;
        if ((e = out_bytes_slack(d, lit, literal_len))) return e;
        if (match_len != 0) {
            n_match_bytes += match_len;
            if ((e = out_match(d, match_len, match_distance))) return e;
            trace_push(d, literal_len, match_len, match_distance);
        } else {
            trace_push(d, literal_len, 0, 0);
        }
    }
    if ((e = br_finalize(&r))) return e;
    if (literal_index <= c->block.literal.num && n_match_bytes + literal_index == c->block.n_raw_bytes && sl == 0 &&
        sm == 64 && sd == 128)
        return ORC_OK;
    return ORC_FSE_BAD_LMD_PAYLOAD;
}

/* decode/decoder.rs:102-141 + fse/fse_core.rs:36-88 */
static int decode_fse(fse_core *c, dctx *d, int v2) {
    const uint8_t *s = d->src + d->pos; size_t rest = d->src_len - d->pos;
    uint32_t hdr = v2 ? V2_HEADER_SIZE : V1_HEADER_SIZE, nw;
    int e;
    if (rest < hdr) return ORC_PAYLOAD_UNDERFLOW; /* decode/take.rs:8-19 */
    if (v2) { if ((e = fse_block_load_v2(&c->block, s, &nw))) return e; }
    else { if ((e = fse_block_load_v1(&c->block, s))) return e; nw = V1_WEIGHT_PAYLOAD_BYTES; }
    if (rest - hdr < nw) return ORC_PAYLOAD_UNDERFLOW;
    if ((e = v2 ? weights_load_v2(c->weights, s + hdr, nw) : weights_load_v1(c->weights, s + hdr))) return e;
    decoder_init(c);
    d->pos += hdr + nw - 8; /* the literal BitSrc borrows 8 bytes of header as its pad: fse_core.rs:30-33 */
    s = d->src + d->pos; rest = d->src_len - d->pos;
    size_t nlit = (size_t)c->block.literal.n_payload_bytes + 8;
    if (rest < nlit) return ORC_PAYLOAD_UNDERFLOW;
    if ((e = literals_load(c, s, nlit))) return e;
    d->pos += nlit;
    s = d->src + d->pos; rest = d->src_len - d->pos;
    size_t nlmd = c->block.lmd.n_payload_bytes;
    if (rest < nlmd) return ORC_PAYLOAD_UNDERFLOW;
    if ((e = fse_decode_internal(c, d, s, nlmd))) return e;
    d->pos += nlmd;
    return ORC_OK;
}

/* vn/constants.rs:24-72.  Encoded compactly: classify by bit pattern instead of a 256-entry list;
 * verified against the table by tests (tests/test_oracle_vn.py::test_opcode_table). */
enum { OP_SML_L, OP_LRG_L, OP_SML_M, OP_LRG_M, OP_PRE_D, OP_SML_D, OP_MED_D, OP_LRG_D, OP_EOS, OP_UDEF, OP_NOP };
static int vn_op(uint8_t b) {
    uint32_t hi = b >> 4, lo3 = b & 7;
    if (hi == 0xE) return b == 0xE0 ? OP_LRG_L : OP_SML_L;
    if (hi == 0xF) return b == 0xF0 ? OP_LRG_M : OP_SML_M;
    if (hi == 0x7 || hi == 0xD) return OP_UDEF;
    if (hi >= 0xA && hi <= 0xB) return OP_MED_D;
    if (lo3 == 7) return OP_LRG_D;
    if (lo3 == 6) {
        if (b == 0x06) return OP_EOS;
        if (b == 0x0E || b == 0x16) return OP_NOP;
        if (b < 0x40) return OP_UDEF; /* 1E 26 2E 36 3E */
        return OP_PRE_D;
    }
    return OP_SML_D;
}
int orc_vn_op_class(uint8_t b) { return vn_op(b); } /* exported for the table test */

typedef struct { uint32_t n_raw_bytes, n_payload_bytes, match_distance; } vn_core;

/* vn/vn_core.rs:119-286: one view-limited run of atomic ops.  `v`/`vlen` is the view, *used the
 * bytes consumed from it.  Returns ORC_OK with *eos set, or an error (PayloadUnderflow when the
 * view is exhausted). */
static int vn_decode_short(vn_core *c, dctx *d, const uint8_t *v, size_t vlen, size_t *used, int *eos) {
    size_t p = 0; int e;
    *eos = 0; *used = 0;
    if (vlen < 8) return ORC_PAYLOAD_UNDERFLOW;
    for (;;) {
        const uint8_t *s = v + p; size_t rem = vlen - p; /* rem >= 8 invariant */
        uint32_t opu = le32(s);
        uint32_t L = 0, M = 0, D = 0, oplen = 0;
        int op = vn_op((uint8_t)opu);
        switch (op) {
        case OP_SML_L: L = opu & 0xF; oplen = 1; break;
        case OP_LRG_L: L = ((opu >> 8) & 0xFF) + 16; oplen = 2; break;
        case OP_SML_M: M = opu & 0xF; oplen = 1; break;
        case OP_LRG_M: M = ((opu >> 8) & 0xFF) + 16; oplen = 2; break;
        case OP_PRE_D: M = ((opu >> 3) & 7) + 3; L = (opu >> 6) & 3; oplen = 1; break;
        case OP_SML_D: D = ((opu & 7) << 8) | ((opu >> 8) & 0xFF); M = ((opu >> 3) & 7) + 3; L = (opu >> 6) & 3; oplen = 2; break;
        case OP_MED_D: M = (((opu & 7) << 2) | ((opu >> 8) & 3)) + 3; L = (opu >> 3) & 3; D = (opu >> 10) & 0x3FFF; oplen = 3; break;
        case OP_LRG_D: M = ((opu >> 3) & 7) + 3; L = (opu >> 6) & 3; D = (opu >> 8) & 0xFFFF; oplen = 3; break;
        case OP_NOP: oplen = 1; break;
        case OP_EOS:
            if (le64(s) != 0x06ull) return ORC_VN_BAD_PAYLOAD; /* vn_core.rs:179-187 */
            *used = p + 8; *eos = 1; return ORC_OK;
        default: return ORC_VN_BAD_OPCODE;
        }
        /* every op: bytes after the opcode must hold the literals plus 8 (vn_core.rs:189-286) */
        if (rem - oplen < (size_t)L + 8) return ORC_PAYLOAD_UNDERFLOW;
        if (op == OP_SML_D || op == OP_MED_D || op == OP_LRG_D) c->match_distance = D;
        if (op == OP_SML_L || op == OP_LRG_L) {
            if ((e = out_bytes(d, s + oplen, L))) return e;
            trace_push(d, L, 0, 0);
        } else if (op == OP_SML_M || op == OP_LRG_M) {
            if ((e = out_match(d, M, c->match_distance))) return e;
            trace_push(d, NO_L, M, c->match_distance);
        } else if (op != OP_NOP) {
            if ((e = out_bytes(d, s + oplen, L))) return e; /* write_quad: lz/writer.rs:131-140 */
            if ((e = out_match(d, M, c->match_distance))) return e;
            trace_push(d, L, M, c->match_distance);
        }
        p += oplen + L; *used = p;
    }
}
/* decode/decoder.rs:143-158 + vn/vn_core.rs:41-116 (decode_mark with dst_mark = u64::MAX) */
static int decode_vxn(dctx *d) {
    size_t rest = d->src_len - d->pos;
    if (rest < VN_HEADER_SIZE) return ORC_PAYLOAD_UNDERFLOW;
    vn_core c; c.n_raw_bytes = le32(d->src + d->pos + 4); c.n_payload_bytes = le32(d->src + d->pos + 8); c.match_distance = 0;
    d->pos += VN_HEADER_SIZE;
    for (;;) {
        size_t src_len = d->src_len - d->pos;
        size_t vlen = src_len < VN_PAYLOAD_LIMIT ? src_len : VN_PAYLOAD_LIMIT;
        size_t out0 = d->out, used; int eos;
        int res = vn_decode_short(&c, d, d->src + d->pos, vlen, &used, &eos);
        size_t produced = d->out - out0;
        if (used > c.n_payload_bytes) return ORC_PAYLOAD_UNDERFLOW;
        if (produced > c.n_raw_bytes) return ORC_VN_BAD_PAYLOAD;
        c.n_payload_bytes -= (uint32_t)used; c.n_raw_bytes -= (uint32_t)produced;
        int cycle = src_len > VN_PAYLOAD_LIMIT;
        d->pos += used;
        if (res == ORC_OK) { /* Ok(false): Eos reached */
            if (c.n_payload_bytes != 0) return ORC_PAYLOAD_OVERFLOW;
            if (c.n_raw_bytes != 0) return ORC_VN_BAD_PAYLOAD;
            return ORC_OK;
        }
        if (res == ORC_PAYLOAD_UNDERFLOW && cycle) continue;
        return res;
    }
}
/* decode/decoder.rs:160-174 + raw/block.rs:21-93 */
static int decode_raw(dctx *d) {
    size_t rest = d->src_len - d->pos;
    if (rest < 8) return ORC_PAYLOAD_UNDERFLOW;
    size_t n = le32(d->src + d->pos + 4);
    d->pos += 8; rest -= 8;
    size_t m = n < rest ? n : rest; int e;
    if ((e = out_bytes(d, d->src + d->pos, m))) return e;
    trace_push(d, (uint32_t)m, 0, 0);
    d->pos += m;
    if (m != n) return ORC_PAYLOAD_UNDERFLOW;
    return ORC_OK;
}

/* decode/decoder.rs:72-99 */
static int decode_execute(fse_core *c, dctx *d) {
    int e;
    for (;;) {
        if (d->src_len - d->pos < 4) return ORC_PAYLOAD_UNDERFLOW;
        uint32_t magic = le32(d->src + d->pos);
        if (magic == BM_VX1) { if ((e = decode_fse(c, d, 0))) return e; }
        else if (magic == BM_VX2) { if ((e = decode_fse(c, d, 1))) return e; }
        else if (magic == BM_VXN) { if ((e = decode_vxn(d))) return e; }
        else if (magic == BM_RAW) { if ((e = decode_raw(d))) return e; }
        else if (magic == BM_EOS) break;
        else return ORC_BAD_BLOCK;
    }
    if (d->src_len - d->pos != 4) return ORC_PAYLOAD_OVERFLOW;
    d->pos += 4;
    return ORC_OK;
}

int orc_decode_trace(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *out_len,
                     orc_lmd_t *trace, size_t trace_cap, size_t *trace_len) {
    init_tables();
    fse_core *c = (fse_core *)malloc(sizeof(fse_core));
    if (!c) return ORC_BUFFER_OVERFLOW;
    dctx d = {src, src_len, 0, dst, dst_cap, 0, trace, trace_cap, 0};
    int e = decode_execute(c, &d);
    free(c);
    if (out_len) *out_len = d.out;
    if (trace_len) *trace_len = d.trace_len;
    return e;
}
int orc_decode(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *out_len) {
    return orc_decode_trace(src, src_len, dst, dst_cap, out_len, NULL, 0, NULL);
}

/* Header-only walk.  Block byte lengths: SURVEY.md §8 a7b (decode/decoder.rs:102-174). */
int orc_probe(const uint8_t *src, size_t src_len, size_t *raw_len, size_t *n_blocks) {
    size_t pos = 0, raw = 0, nb = 0;
    for (;;) {
        if (src_len - pos < 4) return ORC_PAYLOAD_UNDERFLOW;
        uint32_t magic = le32(src + pos);
        size_t rest = src_len - pos, blen;
        if (magic == BM_EOS) break;
        if (magic == BM_RAW) { if (rest < 8) return ORC_PAYLOAD_UNDERFLOW; blen = 8 + (size_t)le32(src + pos + 4); }
        else if (magic == BM_VXN) { if (rest < 12) return ORC_PAYLOAD_UNDERFLOW; blen = 12 + (size_t)le32(src + pos + 8); }
        else if (magic == BM_VX2) {
            fse_block b; uint32_t nw; int e;
            if (rest < V2_HEADER_SIZE) return ORC_PAYLOAD_UNDERFLOW;
            if ((e = fse_block_load_v2(&b, src + pos, &nw))) return e;
            blen = (size_t)V2_HEADER_SIZE + nw + b.literal.n_payload_bytes + b.lmd.n_payload_bytes;
        } else if (magic == BM_VX1) {
            fse_block b; int e;
            if (rest < V1_HEADER_SIZE) return ORC_PAYLOAD_UNDERFLOW;
            if ((e = fse_block_load_v1(&b, src + pos))) return e;
            blen = (size_t)V1_HEADER_SIZE + V1_WEIGHT_PAYLOAD_BYTES + b.literal.n_payload_bytes + b.lmd.n_payload_bytes;
        } else return ORC_BAD_BLOCK;
        if (rest < blen) return ORC_PAYLOAD_UNDERFLOW;
        raw += le32(src + pos + 4); nb++; pos += blen;
    }
    if (src_len - pos != 4) return ORC_PAYLOAD_OVERFLOW;
    if (raw_len) *raw_len = raw;
    if (n_blocks) *n_blocks = nb;
    return ORC_OK;
}

/* ==========================================================================================
 * ENCODE
 * ========================================================================================== */

/* Fixed-capacity stand-in for the reference's growing Vec<u8>: `pos` always advances (so sizes stay
 * exact); bytes beyond `cap` are dropped and flagged, which the C-ABI reports as BufferOverflow. */
typedef struct { uint8_t *p; size_t cap, pos; int overflow; } obuf;
static inline void ob_put(obuf *o, const void *s, size_t n) {
    if (o->pos + n <= o->cap) memcpy(o->p + o->pos, s, n); else o->overflow = 1;
    o->pos += n;
}
static inline void ob_zero(obuf *o, size_t n) {
    if (o->pos + n <= o->cap) memset(o->p + o->pos, 0, n); else o->overflow = 1;
    o->pos += n;
}
static inline void ob_u32(obuf *o, uint32_t v) { uint8_t b[4]; st32(b, v); ob_put(o, b, 4); }
static inline void ob_u64(obuf *o, uint64_t v) { uint8_t b[8]; st64(b, v); ob_put(o, b, 8); }

/* bits/bit_writer.rs:16-57 */
typedef struct { uint64_t accum; int64_t accum_bits; obuf *o; } bitwriter;
static inline void bw_push(bitwriter *w, uint64_t bits, uint32_t n) { w->accum |= bits << w->accum_bits; w->accum_bits += n; }
static inline void bw_flush(bitwriter *w) {
    size_t n_bytes = (size_t)w->accum_bits / 8; uint8_t b[8]; st64(b, w->accum);
    obuf *o = w->o;
    if (o->pos + 8 <= o->cap) { memcpy(o->p + o->pos, b, 8); o->pos += n_bytes; } /* whole word, the unused tail is overwritten by what follows */
    else ob_put(o, b, n_bytes);
    w->accum = n_bytes == 8 ? 0 : w->accum >> (n_bytes * 8); w->accum_bits -= (int64_t)n_bytes * 8;
}
static inline uint32_t bw_finalize(bitwriter *w) {
    size_t n_bytes = ((size_t)w->accum_bits + 7) / 8; uint8_t b[8]; st64(b, w->accum);
    ob_put(w->o, b, n_bytes);
    w->accum_bits -= (int64_t)n_bytes * 8;
    return (uint32_t)(-w->accum_bits);
}

typedef struct { int16_t t_k, t_w; } eentry; /* fse/encoder.rs:184-188 */
/* fse/encoder.rs:219-240 */
static void build_e_table(const uint16_t *weights, int n_sym, uint32_t n_states, eentry *table) {
    uint32_t n_clz = clz32(n_states), total = 0;
    for (int i = 0; i < n_sym; i++) {
        uint32_t w = weights[i]; eentry e;
        if (w == 0) { e.t_k = (int16_t)(-(int32_t)n_states); e.t_w = 0; }
        else {
            uint32_t k = clz32(w) - n_clz;
            e.t_k = (int16_t)((int16_t)(1024 * (int32_t)k) - (int16_t)(w << k));
            e.t_w = (int16_t)((int16_t)n_states + (int16_t)total - (int16_t)w);
        }
        table[i] = e; total += w;
    }
}
/* fse/encoder.rs:190-200 */
static inline void e_encode(eentry e, bitwriter *w, uint32_t *state) {
    uint32_t s = *state;
    uint32_t n_bits = (uint32_t)((int32_t)e.t_k + (int32_t)s) >> 10;
    *state = (uint32_t)((int32_t)e.t_w + ((int32_t)s >> n_bits));
    bw_push(w, s & ((1u << n_bits) - 1), n_bits);
}

/* fse/weights.rs:218-278 */
static void normalize_m1(uint16_t *weights, size_t n, uint32_t in_total, uint32_t out_total) {
    int32_t remaining = 0; size_t max_index = 0;
    if (in_total != 0) {
        uint32_t shift = clz32(out_total), multiply = (1u << 31) / in_total, round = 1u << (shift - 1), max_weight = 0;
        remaining = (int32_t)out_total;
        for (size_t i = 0; i < n; i++) {
            if (weights[i] == 0) continue;
            uint32_t f = ((uint32_t)weights[i] * multiply + round) >> shift;
            if (f == 0) f = 1;
            weights[i] = (uint16_t)f; remaining -= (int32_t)f;
            if (f > max_weight) { max_weight = f; max_index = i; }
        }
    }
    if (-remaining < (int32_t)weights[max_index] / 4) {
        weights[max_index] = (uint16_t)((int32_t)weights[max_index] + remaining);
    } else {
        uint32_t overflow = (uint32_t)(-remaining);
        for (int shift = 3; shift >= 0; shift--)
            for (size_t i = 0; i < n; i++) {
                if (overflow == 0) break;
                if (weights[i] == 0) continue;
                uint32_t k = ((uint32_t)weights[i] - 1) >> shift; if (k > overflow) k = overflow;
                weights[i] = (uint16_t)(weights[i] - k); overflow -= k;
            }
    }
}
void orc_normalize_m1(uint16_t *weights, size_t n, uint32_t in_total, uint32_t out_total) { normalize_m1(weights, n, in_total, out_total); }

typedef struct { uint16_t l, m; uint32_t d; } lmd_pack; /* lmd/lmd_pack.rs:6-8 */

/* fse/buffer.rs:17-23 + fse/backend.rs:11-15 */
typedef struct {
    uint8_t literals[LITERALS_PER_BLOCK + 512]; uint32_t n_literals;
    lmd_pack lmds[LMDS_PER_BLOCK]; uint32_t n_lmds;
    uint32_t n_match_bytes, match_distance;
    uint16_t weights[N_WEIGHTS];
    eentry el[L_SYMBOLS], em[M_SYMBOLS], ed[D_SYMBOLS], eu[U_SYMBOLS];
} fse_backend;

static void fse_buffer_reset(fse_backend *b) { b->n_literals = 0; b->n_lmds = 0; b->n_match_bytes = 0; b->match_distance = 0; }
/* fse/buffer.rs:100-117 */
static void fse_push_l(fse_backend *b, uint16_t l) { b->match_distance = 1; lmd_pack p = {l, 0, 1}; b->lmds[b->n_lmds++] = p; }
static void fse_push_lmd(fse_backend *b, uint16_t l, uint16_t m, uint32_t d) {
    if (b->match_distance == d) d = 0; else b->match_distance = d;
    lmd_pack p = {l, m, d}; b->lmds[b->n_lmds++] = p; b->n_match_bytes += m;
}
static void fse_lit_push(fse_backend *b, const uint8_t **lit, size_t *lit_len, uint32_t n) {
    memcpy(b->literals + b->n_literals, *lit, n); b->n_literals += n; *lit += n; *lit_len -= n;
}
/* fse/buffer.rs:45-97 */
static int fse_buffer_push(fse_backend *b, const uint8_t **lit, size_t *lit_len, uint32_t *match_len, uint32_t d) {
    while (*lit_len > MAX_L_VALUE) {
        if (b->n_lmds == LMDS_PER_BLOCK) return 0;
        uint32_t limit = LITERALS_PER_BLOCK - b->n_literals;
        if (MAX_L_VALUE <= limit) { fse_lit_push(b, lit, lit_len, MAX_L_VALUE); fse_push_l(b, MAX_L_VALUE); }
        else if (limit != 0) { fse_lit_push(b, lit, lit_len, limit); fse_push_l(b, (uint16_t)limit); return 0; }
        else return 0;
    }
    if (b->n_lmds == LMDS_PER_BLOCK) return 0;
    uint32_t literal_len = (uint32_t)*lit_len;
    uint32_t limit = LITERALS_PER_BLOCK - b->n_literals;
    if (literal_len <= limit) fse_lit_push(b, lit, lit_len, literal_len);
    else if (limit != 0) { fse_lit_push(b, lit, lit_len, limit); fse_push_l(b, (uint16_t)limit); return 0; }
    else return 0;
    while (*match_len > MAX_M_VALUE) {
        fse_push_lmd(b, (uint16_t)literal_len, MAX_M_VALUE, d);
        *match_len -= MAX_M_VALUE; literal_len = 0;
        if (b->n_lmds == LMDS_PER_BLOCK) return 0;
    }
    fse_push_lmd(b, (uint16_t)literal_len, (uint16_t)*match_len, d);
    *match_len = 0;
    return 1;
}

/* fse/weights.rs:25-64 */
static void weights_load(fse_backend *b) {
    uint16_t *w = b->weights;
    memset(w, 0, sizeof(b->weights));
    if (b->n_lmds) {
        for (uint32_t i = 0; i < b->n_lmds; i++) {
            w[L_SYM_FROM_VALUE[b->lmds[i].l]]++; w[20 + M_SYM_FROM_VALUE[b->lmds[i].m]]++; w[40 + d_sym_from_value(b->lmds[i].d)]++;
        }
        normalize_m1(w, L_SYMBOLS, b->n_lmds, L_STATES);
        normalize_m1(w + 20, M_SYMBOLS, b->n_lmds, M_STATES);
        normalize_m1(w + 40, D_SYMBOLS, b->n_lmds, D_STATES);
    }
    if (b->n_literals) {
        for (uint32_t i = 0; i < b->n_literals; i++) w[104 + b->literals[i]]++;
        normalize_m1(w + 104, U_SYMBOLS, b->n_literals, U_STATES);
    }
}
/* fse/weight_encoder.rs:23-37 + fse/weights.rs:139-163 */
static uint32_t weights_store_v2(const uint16_t *w, obuf *o) {
    uint64_t accum = 0; uint32_t accum_bits = 0, n = 0;
    for (int i = 0; i < N_WEIGHTS; i++) {
        uint32_t v = w[i], u, ub;
        switch (v) {
        case 0: u = 0; ub = 2; break; case 1: u = 2; ub = 2; break; case 2: u = 1; ub = 3; break; case 3: u = 5; ub = 3; break;
        case 4: u = 3; ub = 5; break; case 5: u = 11; ub = 5; break; case 6: u = 19; ub = 5; break; case 7: u = 27; ub = 5; break;
        default: if (v < 24) { u = ((v - 8) << 4) + 7; ub = 8; } else { u = ((v - 24) << 4) + 15; ub = 14; }
        }
        accum |= (uint64_t)u << accum_bits; accum_bits += ub;
        while (accum_bits >= 8) { uint8_t by = (uint8_t)accum; ob_put(o, &by, 1); accum >>= 8; accum_bits -= 8; n++; }
    }
    if (accum_bits > 0) { uint8_t by = (uint8_t)accum; ob_put(o, &by, 1); n++; }
    return n;
}
/* fse/literals.rs:93-133 */
static void literals_store(fse_backend *b, obuf *o, lit_param *out) {
    size_t mark = o->pos;
    uint32_t n = (b->n_literals + 3) / 4 * 4;
    bitwriter w = {0, 0, o};
    uint32_t st[4] = {U_STATES, U_STATES, U_STATES, U_STATES};
    for (uint32_t i = n; i != 0; i -= 4) {
        e_encode(b->eu[b->literals[i - 1]], &w, &st[3]);
        e_encode(b->eu[b->literals[i - 2]], &w, &st[2]);
        e_encode(b->eu[b->literals[i - 3]], &w, &st[1]);
        e_encode(b->eu[b->literals[i - 4]], &w, &st[0]);
        bw_flush(&w);
    }
    out->bits = bw_finalize(&w);
    out->num = n; out->n_payload_bytes = (uint32_t)(o->pos - mark);
    for (int i = 0; i < 4; i++) out->state[i] = (uint16_t)(st[i] - U_STATES);
}
/* fse/lmds.rs:62-93 + fse/encoder.rs:32-94 */
static void lmds_store(fse_backend *b, obuf *o, lmd_param *out) {
    size_t mark = o->pos;
    ob_u64(o, 0);
    bitwriter w = {0, 0, o};
    uint32_t sl = L_STATES, sm = M_STATES, sd = D_STATES;
    for (uint32_t i = b->n_lmds; i != 0; i--) {
        lmd_pack p = b->lmds[i - 1];
        uint32_t sym = d_sym_from_value(p.d);
        bw_push(&w, p.d - D_BASE_VALUE[sym], D_EXTRA_BITS[sym]); e_encode(b->ed[sym], &w, &sd);
        sym = M_SYM_FROM_VALUE[p.m];
        bw_push(&w, p.m - M_BASE_VALUE[sym], M_EXTRA_BITS[sym]); e_encode(b->em[sym], &w, &sm);
        sym = L_SYM_FROM_VALUE[p.l];
        bw_push(&w, p.l - L_BASE_VALUE[sym], L_EXTRA_BITS[sym]); e_encode(b->el[sym], &w, &sl);
        bw_flush(&w);
    }
    out->bits = bw_finalize(&w);
    out->num = b->n_lmds; out->n_payload_bytes = (uint32_t)(o->pos - mark);
    out->state[0] = (uint16_t)(sl - L_STATES); out->state[1] = (uint16_t)(sm - M_STATES); out->state[2] = (uint16_t)(sd - D_STATES);
}
/* fse/block.rs:168-196 */
static void fse_block_store_v2(const fse_block *b, uint8_t *dst, uint32_t n_weight_payload_bytes) {
    st32(dst, BM_VX2); st32(dst + 4, b->n_raw_bytes);
    uint64_t p = (uint64_t)b->literal.num | ((uint64_t)b->literal.n_payload_bytes << 20) | ((uint64_t)b->lmd.num << 40) |
                 ((uint64_t)(7 - b->literal.bits) << 60);
    st64(dst + 8, p);
    p = (uint64_t)b->literal.state[0] | ((uint64_t)b->literal.state[1] << 10) | ((uint64_t)b->literal.state[2] << 20) |
        ((uint64_t)b->literal.state[3] << 30) | ((uint64_t)b->lmd.n_payload_bytes << 40) | ((uint64_t)(7 - b->lmd.bits) << 60);
    st64(dst + 16, p);
    p = (uint64_t)(V2_HEADER_SIZE + n_weight_payload_bytes) | ((uint64_t)b->lmd.state[0] << 32) | ((uint64_t)b->lmd.state[1] << 42) |
        ((uint64_t)b->lmd.state[2] << 52);
    st64(dst + 24, p);
}
/* fse/backend.rs:39-54 */
static void fse_emit_block_v2(fse_backend *b, obuf *o) {
    size_t mark = o->pos;
    ob_zero(o, V2_HEADER_SIZE);
    memset(b->literals + b->n_literals, b->literals[0], 4); /* fse/literals.rs:136-145 pad */
    weights_load(b);
    uint32_t nw = weights_store_v2(b->weights, o);
    build_e_table(b->weights, L_SYMBOLS, L_STATES, b->el);
    build_e_table(b->weights + 20, M_SYMBOLS, M_STATES, b->em);
    build_e_table(b->weights + 40, D_SYMBOLS, D_STATES, b->ed);
    build_e_table(b->weights + 104, U_SYMBOLS, U_STATES, b->eu);
    fse_block blk;
    literals_store(b, o, &blk.literal);
    lmds_store(b, o, &blk.lmd);
    blk.n_raw_bytes = b->n_literals + b->n_match_bytes;
    if (!o->overflow) fse_block_store_v2(&blk, o->p + mark, nw);
    fse_buffer_reset(b);
}

/* vn/backend.rs:25-31 */
typedef struct { size_t mark; uint32_t match_distance, n_literals, n_match_bytes; } vn_backend;

/* Backend dispatch (encode/backend.rs:8-34). kind: 0 = FSE, 1 = VN, 2 = capture (tests). */
typedef struct {
    int kind; obuf *o; fse_backend *fse; vn_backend vn;
    orc_lmd_t *cap; size_t cap_n, cap_len;
} backend;

static void vn_put_op(obuf *o, uint32_t opu, uint32_t oplen, const uint8_t *lit, uint32_t lit_len) {
    uint8_t b[4]; st32(b, opu); ob_put(o, b, oplen); ob_put(o, lit, lit_len);
}
/* vn/backend.rs:58-74 / 138-162 + vn/opc.rs:7-47 */
static void vn_push_literal_runs(obuf *o, const uint8_t **lit, size_t *len, size_t keep_below) {
    while (*len >= 0x10) {
        uint32_t n = *len < 0x10F ? (uint32_t)*len : 0x10F;
        vn_put_op(o, 0xE0u | ((n - 0x10) << 8), 2, *lit, n); *lit += n; *len -= n;
    }
    if (*len >= keep_below && *len > 0) { uint32_t n = (uint32_t)*len; vn_put_op(o, 0xE0u | n, 1, *lit, n); *lit += n; *len -= n; }
}
/* vn/backend.rs:76-124 + vn/opc.rs */
static void vn_push_match(backend *b, const uint8_t *lit, size_t lit_len, uint32_t match_len, uint32_t d) {
    obuf *o = b->o; vn_backend *v = &b->vn;
    v->n_literals += (uint32_t)lit_len; v->n_match_bytes += match_len;
    vn_push_literal_runs(o, &lit, &lit_len, 4);
    uint32_t L = (uint32_t)lit_len;
    uint32_t n = 0x0A - 2 * L; if (n > match_len) n = match_len;
    match_len -= n;
    if (d == v->match_distance) {
        if (L == 0) vn_put_op(o, 0xF0u | n, 1, lit, 0);                                  /* SmlM */
        else vn_put_op(o, 0x6u | ((n - 3) << 3) | (L << 6), 1, lit, L);                  /* PreD */
    } else if (d < 0x600) {
        vn_put_op(o, ((d >> 8) & 7) | ((n - 3) << 3) | (L << 6) | ((d & 0xFF) << 8), 2, lit, L); /* SmlD */
    } else if (d >= 0x4000 || match_len == 0 || n + match_len > 0x22) {
        vn_put_op(o, 0x7u | ((n - 3) << 3) | (L << 6) | (d << 8), 3, lit, L);            /* LrgD */
    } else {
        uint32_t m = n - 3;
        vn_put_op(o, ((m >> 2) & 7) | (L << 3) | (0x5u << 5) | ((m & 3) << 8) | (d << 10), 3, lit, L); /* MedD */
    }
    v->match_distance = d;
    while (match_len > 0x0F) { uint32_t lim = match_len < 0x10F ? match_len : 0x10F; vn_put_op(o, 0xF0u | ((lim - 0x10) << 8), 2, lit, 0); match_len -= lim; }
    if (match_len > 0) vn_put_op(o, 0xF0u | match_len, 1, lit, 0);
}

static void be_init(backend *b) {
    if (b->kind == 0) fse_buffer_reset(b->fse);
    else if (b->kind == 1) { b->vn.mark = b->o->pos; b->vn.match_distance = 0; b->vn.n_literals = 0; b->vn.n_match_bytes = 0; ob_zero(b->o, VN_HEADER_SIZE); }
}
static void be_push_match(backend *b, const uint8_t *lit, size_t lit_len, uint32_t match_len, uint32_t d) {
    if (b->kind == 0) { /* fse/backend.rs:76-90 */
        while (!fse_buffer_push(b->fse, &lit, &lit_len, &match_len, d)) fse_emit_block_v2(b->fse, b->o);
    } else if (b->kind == 1) vn_push_match(b, lit, lit_len, match_len, d);
    else { if (b->cap_len < b->cap_n) { orc_lmd_t t = {(uint32_t)lit_len, match_len, d}; b->cap[b->cap_len] = t; } b->cap_len++; }
}
static void be_push_literals(backend *b, const uint8_t *lit, size_t lit_len) {
    if (b->kind == 0) be_push_match(b, lit, lit_len, 0, 1); /* fse/backend.rs:66-73 */
    else if (b->kind == 1) { b->vn.n_literals += (uint32_t)lit_len; vn_push_literal_runs(b->o, &lit, &lit_len, 0); }
    else { if (b->cap_len < b->cap_n) { orc_lmd_t t = {(uint32_t)lit_len, 0, 0}; b->cap[b->cap_len] = t; } b->cap_len++; }
}
static void be_finalize(backend *b) {
    if (b->kind == 0) fse_emit_block_v2(b->fse, b->o); /* fse/backend.rs:92-95 */
    else if (b->kind == 1) {                           /* vn/backend.rs:127-135 */
        ob_u64(b->o, 0x06);
        if (!b->o->overflow) {
            uint8_t *h = b->o->p + b->vn.mark;
            st32(h, BM_VXN); st32(h + 4, b->vn.n_literals + b->vn.n_match_bytes);
            st32(h + 8, (uint32_t)(b->o->pos - b->vn.mark) - VN_HEADER_SIZE);
        }
    }
}

/* encode/history.rs:144-154, 101-131 */
/* Item {val, idx} x 4, newest first; kept as two 4-lane vectors (the layout is private to the table) so that push is two
 * shift-inserts. */
typedef struct { uint32_t idx[HASH_WIDTH], val[HASH_WIDTH]; } __attribute__((aligned(32))) history;
struct orc_encoder { history table[1 << HASH_BITS]; fse_backend fse; };

orc_encoder *orc_encoder_create(void) {
    init_tables();
    const size_t sz = (sizeof(orc_encoder) + 63) & ~(size_t)63;
    orc_encoder *e = (orc_encoder *)aligned_alloc(64, sz); /* the table's buckets are 32-byte vectors */
    if (e) memset(e, 0, sz);
    return e;
}
void orc_encoder_destroy(orc_encoder *e) { free(e); }

typedef struct { uint32_t idx, match_idx, match_len; } match_t; /* encode/match_object.rs:4-8 */

/* match_kit/match_fast.rs:22-49 */
static inline uint32_t match_inc(const uint8_t *b, size_t index, size_t match_index, size_t len, size_t max) {
    while (len + 8 <= max) {
        uint64_t x = le64(b + index + len) ^ le64(b + match_index + len);
        if (x) return (uint32_t)(len + (size_t)(__builtin_ctzll(x) / 8));
        len += 8;
    }
    while (len < max) { if (b[index + len] != b[match_index + len]) return (uint32_t)len; len++; }
    return (uint32_t)max;
}
/* match_kit/match_fast.rs:61-89 */
static inline uint32_t match_dec(const uint8_t *b, size_t index, size_t match_index, size_t max) {
    size_t len = 0;
    while (len + 8 <= max) {
        uint64_t x = le64(b + index - len - 8) ^ le64(b + match_index - len - 8);
        if (x) return (uint32_t)(len + (size_t)(__builtin_clzll(x) / 8));
        len += 8;
    }
    while (len != max) { if (b[index - len - 1] != b[match_index - len - 1]) break; len++; }
    return (uint32_t)len;
}

typedef struct {
    history *table; const uint8_t *src; size_t len; int vn;
    match_t pending; uint32_t literal_index;
    backend *be;
} frontend;

static inline uint32_t fe_hash(uint32_t val, int vn) { /* fse/object.rs:38-43, vn/object.rs:33-47 */
    if (vn) val &= 0x00FFFFFFu;
    return (val * 0x9E3779B1u) >> (32 - HASH_BITS);
}
/* encode/history.rs:24-31 */
static inline __attribute__((always_inline)) history fe_push(frontend *f, uint32_t val, uint32_t idx, const int vn) {
    history *q = &f->table[fe_hash(val, vn)], copy;
    const __m128i qi = _mm_load_si128((const __m128i *)q->idx), qv = _mm_load_si128((const __m128i *)q->val);
    _mm_store_si128((__m128i *)copy.idx, qi); _mm_store_si128((__m128i *)copy.val, qv);
    _mm_store_si128((__m128i *)q->idx, _mm_insert_epi32(_mm_slli_si128(qi, 4), (int)idx, 0));
    _mm_store_si128((__m128i *)q->val, _mm_insert_epi32(_mm_slli_si128(qv, 4), (int)val, 0));
    return copy;
}
/* encode/frontend_bytes.rs:214-268 */
static inline __attribute__((always_inline)) match_t fe_find_match(frontend *f, const history *queue, uint32_t val, uint32_t idx, const int vn) {
    match_t m = {0, 0, 0};
    uint32_t max_d = vn ? VN_MAX_D : MAX_D_VALUE;
    for (int i = 0; i < HASH_WIDTH; i++) {
        uint32_t distance = idx - queue->idx[i];
        if (distance > max_d) break;
        uint32_t x = val ^ queue->val[i], len;
        if (x == 0) len = match_inc(f->src, idx, queue->idx[i], 4, f->len - idx);
        else if (vn && (x & 0x00FFFFFFu) == 0) len = 3;
        else len = 0;
        if (len > m.match_len) { m.match_len = len; m.match_idx = queue->idx[i]; }
    }
    if (m.match_len != 0) {
        m.idx = idx;
        size_t literal_len = (size_t)idx - f->literal_index;
        size_t max = literal_len < m.match_idx ? literal_len : m.match_idx;
        uint32_t dec = match_dec(f->src, m.idx, m.match_idx, max);
        m.idx -= dec; m.match_idx -= dec; m.match_len += dec;
    }
    return m;
}
/* encode/match_object.rs:12-33 */
static int match_select(match_t *self, match_t incoming, match_t *select) {
    if (incoming.match_len == 0) return 0;
    if (incoming.match_len >= GOOD_MATCH_LEN) { *select = incoming; self->match_len = 0; return 1; }
    if (self->match_len == 0) { *self = incoming; return 0; }
    if ((int32_t)(self->idx + self->match_len - incoming.idx) <= 0) { *select = *self; *self = incoming; return 1; }
    if (incoming.match_len > self->match_len) { *select = incoming; self->match_len = 0; return 1; }
    *select = *self; self->match_len = 0; return 1;
}
/* encode/frontend_bytes.rs:287-302 */
static void fe_push_match(frontend *f, match_t m) {
    const uint8_t *lit = f->src + f->literal_index; size_t lit_len = (size_t)m.idx - f->literal_index;
    f->literal_index = m.idx + m.match_len;
    be_push_match(f->be, lit, lit_len, m.match_len, m.idx - m.match_idx);
}
/* encode/frontend_bytes.rs:121-211,271-284,305-344 (finalize = match_blocks + flush_pending +
 * flush_literals + backend.finalize).  Inputs above BLOCK_GUIDE (2 GiB, :160-182,348-375) are out
 * of scope for the batched path and rejected by the caller. */
static inline __attribute__((always_inline)) void fe_finalize_k(frontend *f, const int vn) {
    uint32_t end = (uint32_t)f->len - 3, index = 0;
    for (;;) {
        uint32_t val = le32(f->src + index);
        history queue = fe_push(f, val, index, vn);
        match_t incoming = fe_find_match(f, &queue, val, index, vn), sel;
        if (match_select(&f->pending, incoming, &sel)) {
            fe_push_match(f, sel);
            if (f->literal_index >= end) break;
            index++;
            while (index < f->literal_index) { fe_push(f, le32(f->src + index), index, vn); index++; } /* sync_history */
            if (index >= end) break;
        } else {
            index++;
            if (index == end) break;
        }
    }
    if (f->pending.match_len != 0) { fe_push_match(f, f->pending); f->pending.match_len = 0; }
    if (f->len - f->literal_index != 0) {
        be_push_literals(f->be, f->src + f->literal_index, f->len - f->literal_index);
        f->literal_index = (uint32_t)f->len;
    }
    be_finalize(f->be);
}
/* the 3-byte-key (LZVN) and 4-byte-key (FSE) front ends as two specialised copies of the loop */
static void fe_finalize(frontend *f) { if (f->vn) fe_finalize_k(f, 1); else fe_finalize_k(f, 0); }
static void fe_init(frontend *f, orc_encoder *e, const uint8_t *src, size_t len, int vn, backend *be) {
    const __m256i reset = _mm256_setr_epi32((int)(0u - Q1), (int)(0u - Q1), (int)(0u - Q1), (int)(0u - Q1), 0, 0, 0, 0);
    for (size_t i = 0; i < (1u << HASH_BITS); i++) /* encode/history.rs:72-83: idx = Q0 - Q1, val = 0 */
        _mm256_store_si256((__m256i *)&e->table[i], reset);
    f->table = e->table; f->src = src; f->len = len; f->vn = vn; f->literal_index = 0; f->be = be;
    f->pending.idx = f->pending.match_idx = f->pending.match_len = 0;
}

size_t orc_encode_bound(size_t n) { return n + n / 4 + (n / 16384 + 2) * 768 + 64; }

/* encode/frontend_bytes.rs:41-111 + raw/ops.rs:19-30 */
int orc_encode(orc_encoder *e, const uint8_t *src, size_t len, uint8_t *dst, size_t dst_cap, size_t *out_len) {
    if (len > 0x7FFFFFFFu) return ORC_BUFFER_OVERFLOW;
    obuf o = {dst, dst_cap, 0, 0};
    backend be; memset(&be, 0, sizeof(be)); be.o = &o; be.fse = &e->fse;
    frontend f;
    int raw = 0;
    if (len > VN_CUTOFF) { be.kind = 0; fe_init(&f, e, src, len, 0, &be); be_init(&be); fe_finalize(&f); }
    else if (len > RAW_CUTOFF) {
        be.kind = 1; fe_init(&f, e, src, len, 1, &be);
        size_t mark = o.pos; be_init(&be); fe_finalize(&f);
        if (len < RAW_LIMIT && len + 8 <= o.pos - mark) { o.pos = mark; o.overflow = 0; raw = 1; } /* truncate: ops/truncate.rs:12-23 */
    } else raw = 1;
    if (raw) { ob_u32(&o, BM_RAW); ob_u32(&o, (uint32_t)len); ob_put(&o, src, len); }
    ob_u32(&o, BM_EOS);
    if (out_len) *out_len = o.pos;
    return o.overflow ? ORC_BUFFER_OVERFLOW : ORC_OK;
}

int orc_frontend_lmds(orc_encoder *e, const uint8_t *src, size_t len, int vn, orc_lmd_t *lmds, size_t cap, size_t *n) {
    if (len < 4 || len > 0x7FFFFFFFu) return ORC_BUFFER_OVERFLOW;
    backend be; memset(&be, 0, sizeof(be)); be.kind = 2; be.cap = lmds; be.cap_n = cap;
    frontend f; fe_init(&f, e, src, len, vn, &be); fe_finalize(&f);
    if (n) *n = be.cap_len;
    return be.cap_len > cap ? ORC_BUFFER_OVERFLOW : ORC_OK;
}

static int encode_lmds_common(backend *be, const uint8_t *lit, size_t n_lit, const orc_lmd_t *lmds, size_t n) {
    be_init(be);
    for (size_t i = 0; i < n; i++) { /* test_utils/lmds.rs:36-61 */
        if (lmds[i].literal_len > n_lit) return ORC_BUFFER_OVERFLOW;
        if (lmds[i].match_len == 0) be_push_literals(be, lit, lmds[i].literal_len);
        else be_push_match(be, lit, lmds[i].literal_len, lmds[i].match_len, lmds[i].match_distance);
        lit += lmds[i].literal_len; n_lit -= lmds[i].literal_len;
    }
    be_finalize(be);
    return ORC_OK;
}
int orc_fse_encode_lmds(orc_encoder *e, const uint8_t *lit, size_t n_lit, const orc_lmd_t *lmds, size_t n, uint8_t *dst,
                        size_t cap, size_t *out_len) {
    obuf o = {dst, cap, 0, 0};
    backend be; memset(&be, 0, sizeof(be)); be.kind = 0; be.o = &o; be.fse = &e->fse;
    int r = encode_lmds_common(&be, lit, n_lit, lmds, n);
    if (out_len) *out_len = o.pos;
    return r ? r : (o.overflow ? ORC_BUFFER_OVERFLOW : ORC_OK);
}
int orc_vn_encode_lmds(const uint8_t *lit, size_t n_lit, const orc_lmd_t *lmds, size_t n, uint8_t *dst, size_t cap,
                       size_t *out_len) {
    init_tables();
    obuf o = {dst, cap, 0, 0};
    backend be; memset(&be, 0, sizeof(be)); be.kind = 1; be.o = &o;
    int r = encode_lmds_common(&be, lit, n_lit, lmds, n);
    if (out_len) *out_len = o.pos;
    return r ? r : (o.overflow ? ORC_BUFFER_OVERFLOW : ORC_OK);
}

/* ==========================================================================================
 * Batched CPU baseline (SURVEY.md §8d "CPU baseline timed alongside")
 * ========================================================================================== */
typedef struct {
    int encode; const uint8_t *src_base; const uint64_t *src_off, *src_len; uint8_t *dst_base;
    const uint64_t *dst_off, *dst_cap; uint64_t *out_len; int32_t *status; size_t n; atomic_size_t next;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    orc_encoder *enc = j->encode ? orc_encoder_create() : NULL;
    fse_core *core = j->encode ? NULL : (fse_core *)malloc(sizeof(fse_core));
    for (;;) {
        size_t i = atomic_fetch_add(&j->next, 1);
        if (i >= j->n) break;
        size_t out = 0; int st;
        if (j->encode) st = orc_encode(enc, j->src_base + j->src_off[i], j->src_len[i], j->dst_base + j->dst_off[i], j->dst_cap[i], &out);
        else {
            dctx d = {j->src_base + j->src_off[i], j->src_len[i], 0, j->dst_base + j->dst_off[i], j->dst_cap[i], 0, NULL, 0, 0};
            st = decode_execute(core, &d); out = d.out;
        }
        j->out_len[i] = out; j->status[i] = st;
    }
    if (enc) orc_encoder_destroy(enc);
    free(core);
    return NULL;
}
static int run_batch(batch_job *j, int n_threads) {
    init_tables();
    if (n_threads < 1) n_threads = 1;
    if (n_threads == 1) { batch_worker(j); return ORC_OK; }
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    int started = 0;
    for (int i = 0; i < n_threads; i++) if (pthread_create(&t[started], NULL, batch_worker, j) == 0) started++;
    if (started == 0) batch_worker(j);
    for (int i = 0; i < started; i++) pthread_join(t[i], NULL);
    free(t);
    return ORC_OK;
}
int orc_decode_batch(const uint8_t *src_base, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst_base,
                     const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n, int n_threads) {
    batch_job j = {0, src_base, src_off, src_len, dst_base, dst_off, dst_cap, out_len, status, n, 0};
    return run_batch(&j, n_threads);
}
int orc_encode_batch(const uint8_t *src_base, const uint64_t *src_off, const uint64_t *src_len, uint8_t *dst_base,
                     const uint64_t *dst_off, const uint64_t *dst_cap, uint64_t *out_len, int32_t *status, size_t n, int n_threads) {
    batch_job j = {1, src_base, src_off, src_len, dst_base, dst_off, dst_cap, out_len, status, n, 0};
    return run_batch(&j, n_threads);
}
