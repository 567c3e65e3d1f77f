/*
 * long_parse_model.c -- CPU model of the segmented front end that encode_long.cuh runs on the GPU for streams
 * longer than 64 KiB.  TEST INFRASTRUCTURE: it checks the ALGORITHM (per-position find words, speculative replay
 * per segment, in-order stitch) against the oracle's sequential front end (orc_frontend_lmds, i.e.
 * encode/frontend_bytes.rs:160-344) on the same bytes; the CUDA kernels are checked separately, frame for frame.
 *
 *   long_parse_model <file> <segment bytes>      prints "OK n_lmds ..." or "MISMATCH ..." (exit code 1)
 *
 * The claim being tested: what a position finds in the history does not depend on the parse, so the words can be
 * computed for all positions at once; the sequential part (backward limit, Match::select) started from a clean
 * state at a segment border falls into step with the true parse after a few matches, and from the first point
 * where both are in the SAME state (nothing pending, cursor == literal index == q) the speculative segment's
 * output is the true output.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../oracle/lzfse_oracle.h"

#define HASH_BITS 14
#define MAX_D 262139u
#define GOOD 40u
#define NONE 0xFFFFFFFFu
#define NS 256u  /* spec emits per segment that carry their state */

typedef struct { uint32_t idx, len, dist, cur_after, p_idx, p_midx, p_len; } emit_t;
typedef struct { uint32_t cur, lit, p_idx, p_midx, p_len; } state_t;

static const uint8_t *src;
static uint32_t len, end_;
static uint32_t *wdist, *wlen;

static uint32_t le32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint32_t hash(uint32_t v) { return (v * 0x9E3779B1u) >> (32 - HASH_BITS); }

static void build_words(void) { /* HistoryTable::push + find_match's candidate loop for every position, in order */
    uint32_t (*tab)[4] = malloc(sizeof(uint32_t[4]) << HASH_BITS);
    for (size_t i = 0; i < (1u << HASH_BITS); i++) for (int k = 0; k < 4; k++) tab[i][k] = NONE;
    for (uint32_t p = 0; p < end_; p++) {
        const uint32_t v = le32(src + p), h = hash(v);
        uint32_t q[4]; memcpy(q, tab[h], 16);
        tab[h][3] = q[2]; tab[h][2] = q[1]; tab[h][1] = q[0]; tab[h][0] = p;
        uint32_t best = 0, bc = 0;
        for (int k = 0; k < 4; k++) {
            if (q[k] == NONE || p - q[k] > MAX_D) break;
            if (le32(src + q[k]) != v) continue;
            uint32_t l = 4; while (l < len - p && src[p + l] == src[q[k] + l]) l++;
            if (l > best) { best = l; bc = q[k]; }
        }
        wdist[p] = best ? p - bc : 0; wlen[p] = best;
    }
    free(tab);
}

/* one position of FrontendBytes::match_any's loop; returns 1 and fills *e when a match is pushed to the back end */
static uint32_t first_cand, first_good; /* position of the first candidate a replay met, and whether its length was >= GOOD */
static uint32_t fwd_cap = 0xFFFFFFFFu; /* speculative replays give up on matches longer than this (encode_long.cuh: kSpecFwdCap) */
static int gave_up;
static int lit_limited; /* set when a backward extension stopped at the literal limit (a longer literal run could have gone on) */
static int step(state_t *s, emit_t *e) {
    const uint32_t cur = s->cur;
    if (wdist[cur] == 0) { s->cur++; return 0; }
    uint32_t i_idx = cur, i_midx = cur - wdist[cur], i_len = wlen[cur];
    if (i_len >= 1023 && i_len > fwd_cap) { if (first_cand == NONE) first_cand = cur; gave_up = 1; s->cur = NONE; return 0; }
    { uint32_t lit = cur - s->lit, lim = lit < i_midx ? lit : i_midx, dec = 0;
      while (dec < lim && src[i_idx - dec - 1] == src[i_midx - dec - 1]) dec++;
      if (dec == lit && dec < i_midx) lit_limited = 1;
      i_idx -= dec; i_midx -= dec; i_len += dec; }
    if (first_cand == NONE) { first_cand = cur; first_good = i_len >= GOOD; }
    int have = 1; uint32_t s_idx = s->p_idx, s_midx = s->p_midx, s_len = s->p_len;
    if (i_len >= GOOD) { s_idx = i_idx; s_midx = i_midx; s_len = i_len; s->p_len = 0; }
    else if (s->p_len == 0) { s->p_idx = i_idx; s->p_midx = i_midx; s->p_len = i_len; have = 0; }
    else if ((int32_t)(s->p_idx + s->p_len - i_idx) <= 0) { s->p_idx = i_idx; s->p_midx = i_midx; s->p_len = i_len; }
    else if (i_len > s->p_len) { s_idx = i_idx; s_midx = i_midx; s_len = i_len; s->p_len = 0; }
    else s->p_len = 0;
    if (!have) { s->cur++; return 0; }
    e->idx = s_idx; e->len = s_len; e->dist = s_idx - s_midx;
    s->lit = s_idx + s_len;
    if (s->lit >= end_) s->cur = end_; else s->cur = cur + 1 > s->lit ? cur + 1 : s->lit;
    e->cur_after = s->cur; e->p_idx = s->p_idx; e->p_midx = s->p_midx; e->p_len = s->p_len;
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
    fseek(f, 0, SEEK_END); len = (uint32_t)ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t *buf = malloc(len + 16); if (fread(buf, 1, len, f) != len) return 2; fclose(f);
    src = buf; end_ = len - 3;
    const uint32_t R = (uint32_t)atoi(argv[2]);
    wdist = malloc(4ull * len); wlen = malloc(4ull * len);
    build_words();
    const uint32_t n_seg = (end_ + R - 1) / R, cap = R / 4 + 8;
    emit_t *spec = malloc(sizeof(emit_t) * (size_t)cap * n_seg), *fix = malloc(sizeof(emit_t) * (size_t)cap * n_seg);
    uint32_t *n_spec = calloc(n_seg, 4), *n_fix = calloc(n_seg, 4), *from = calloc(n_seg, 4);
    state_t *exit_ = malloc(sizeof(state_t) * n_seg);
    uint8_t *lim0 = calloc(n_seg, 1), *good0 = calloc(n_seg, 1);
    uint32_t *cand0 = calloc(n_seg, 4);
    for (uint32_t k = 0; k < n_seg; k++) { /* speculative replay, every segment on its own */
        state_t s = {k * R, k * R, 0, 0, 0};
        const uint32_t se = (k + 1) * R < end_ ? (k + 1) * R : end_;
        emit_t e;
        lit_limited = 0; first_cand = NONE; gave_up = 0; fwd_cap = 0xFFFFFFFFu; /* (the give-up rule is kept for experiments: the CUDA replay now measures long matches with its whole warp instead) */
        while (s.cur < se) if (step(&s, &e)) { if (n_spec[k] == 0) lim0[k] = (uint8_t)lit_limited; if (n_spec[k] >= cap) { printf("spec overflow\n"); return 1; } spec[(size_t)k * cap + n_spec[k]++] = e; }
        if (n_spec[k] == 0) lim0[k] = (uint8_t)lit_limited;
        if (gave_up) { n_spec[k] = 0; lim0[k] = 1; }
        exit_[k] = s; cand0[k] = first_cand; good0[k] = (uint8_t)first_good;
    }
    /* stitch */
    fwd_cap = 0xFFFFFFFFu;
    state_t T = exit_[0];
    uint64_t fix_total = 0, fix_steps = 0, max_steps = 0, unsynced = 0, soft = 0;
    for (uint32_t k = 1; k < n_seg; k++) {
        const uint32_t B = k * R, se = (k + 1) * R < end_ ? (k + 1) * R : end_;
        from[k] = n_spec[k];
        if (T.cur >= se) continue;
        int synced = T.p_len == 0 && T.cur == B && !lim0[k];  /* soft sync: the same cursor, nothing pending, a literal run at least as long */
        if (synced) { from[k] = 0; const uint32_t keep = T.lit; T = exit_[k]; if (n_spec[k] == 0) T.lit = keep; soft++; continue; }
        if (T.cur == B && T.p_len != 0 && T.p_idx + T.p_len <= B && !lim0[k]) {  /* soft sync with a pending match that ends before the segment */
            soft++;
            if (cand0[k] == NONE) { T.cur = se; continue; }
            uint32_t keep = T.lit;
            if (!good0[k]) { emit_t e = {T.p_idx, T.p_len, T.p_idx - T.p_midx, 0, 0, 0, 0}; fix[(size_t)k * cap + n_fix[k]++] = e; keep = T.p_idx + T.p_len; }
            from[k] = 0; T = exit_[k]; if (n_spec[k] == 0) T.lit = keep;
            continue;
        }
        if (T.cur < cand0[k]) T.cur = cand0[k] < se ? cand0[k] : se;  /* nothing happens before the segment's first candidate */
        uint32_t j = 0; uint64_t steps = 0;
        emit_t e;
        while (!synced && T.cur < se) {
            steps++;
            if (step(&T, &e)) {
                if (n_fix[k] >= cap) { printf("fix overflow\n"); return 1; }
                fix[(size_t)k * cap + n_fix[k]++] = e;
                {
                    const uint32_t q = T.lit;
                    const emit_t *sp = spec + (size_t)k * cap;
                    while (j < n_spec[k] && j < NS && sp[j].idx + sp[j].len < q) j++;
                    if (j < n_spec[k] && j < NS && sp[j].idx + sp[j].len == q && sp[j].cur_after == T.cur && sp[j].p_len == T.p_len &&
                        (T.p_len == 0 || (sp[j].p_idx == T.p_idx && sp[j].p_midx == T.p_midx))) { synced = 1; from[k] = j + 1; }
                }
            }
        }
        fix_total += n_fix[k]; fix_steps += steps; if (steps > max_steps) max_steps = steps;
        if (synced) T = exit_[k]; else unsynced++;
    }
    /* the stitched list against the oracle's front end */
    size_t cap_l = (size_t)len / 4 + 16, n_ref = 0;
    orc_lmd_t *ref = malloc(sizeof(orc_lmd_t) * cap_l);
    orc_encoder *enc = orc_encoder_create();
    if (orc_frontend_lmds(enc, src, len, 0, ref, cap_l, &n_ref) != ORC_OK) { printf("oracle failed\n"); return 1; }
    size_t n = 0; uint32_t prev_end = 0; int bad = 0;
#define CHECK(L, M, D) do { if (n >= n_ref || ref[n].literal_len != (L) || ref[n].match_len != (M) || ref[n].match_distance != (D)) { if (!bad) printf("MISMATCH at lmd %zu\n", n); bad = 1; } n++; } while (0)
    for (uint32_t k = 0; k < n_seg && !bad; k++) {
        for (uint32_t i = 0; i < n_fix[k]; i++) { const emit_t e = fix[(size_t)k * cap + i]; CHECK(e.idx - prev_end, e.len, e.dist); prev_end = e.idx + e.len; }
        for (uint32_t i = k == 0 ? 0 : from[k]; i < n_spec[k]; i++) { const emit_t e = spec[(size_t)k * cap + i]; CHECK(e.idx - prev_end, e.len, e.dist); prev_end = e.idx + e.len; }
    }
    if (!bad && T.p_len) { CHECK(T.p_idx - prev_end, T.p_len, T.p_idx - T.p_midx); prev_end = T.p_idx + T.p_len; }
    if (!bad && len - prev_end) CHECK(len - prev_end, 0u, 0u);
    if (!bad && n != n_ref) { printf("MISMATCH count %zu vs %zu\n", n, n_ref); bad = 1; }
    if (bad) return 1;
    printf("OK n_lmds %zu segments %u soft %llu fix_emits %llu fix_steps %llu max_steps %llu unsynced %llu\n", n, n_seg, (unsigned long long)soft,
           (unsigned long long)fix_total, (unsigned long long)fix_steps, (unsigned long long)max_steps, (unsigned long long)unsynced);
    return 0;
}
