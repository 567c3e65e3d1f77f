/* textgen.c -- synthetic benchmark workloads.  Not part of the codec.
 *
 * Text: words drawn from a pool with the reference test-kit LCG x <- x*1103515245 + 12345 (mod 2^32)
 * (test_kit/src/rng.rs:15-17), emitting pool[(x >> 8) % n_words] + ' ', a newline once a line holds
 * >= 72 characters, truncated to the chunk length.  Same definition as tests/testkit.py::synth_text. */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const uint8_t *pool; const uint32_t *word_off; uint32_t n_words;
    uint8_t *dst; uint64_t chunk_len; uint64_t first, n_chunks; uint32_t seed0; int per_stream_lcg;
} job_t;

static void gen_chunk(const job_t *j, uint32_t seed, uint8_t *out, uint64_t len) {
    uint32_t x = seed; uint64_t n = 0; uint32_t line = 0;
    while (n < len) {
        x = x * 1103515245u + 12345u;
        uint32_t w = (x >> 8) % j->n_words;
        uint32_t a = j->word_off[w], b = j->word_off[w + 1], wl = b - a;
        uint64_t c = wl < len - n ? wl : len - n;
        memcpy(out + n, j->pool + a, c); n += c;
        if (n == len) break;
        line += wl + 1;
        if (line >= 72) { out[n++] = '\n'; line = 0; } else out[n++] = ' ';
    }
}
static void *worker(void *arg) {
    const job_t *j = (const job_t *)arg;
    for (uint64_t i = 0; i < j->n_chunks; i++) gen_chunk(j, j->seed0 + (uint32_t)(j->first + i), j->dst + i * j->chunk_len, j->chunk_len);
    return NULL;
}
/* n_chunks chunks of chunk_len bytes, chunk i seeded seed0 + i, written back to back into dst. */
void textgen_chunks(const uint8_t *pool, const uint32_t *word_off, uint32_t n_words, uint8_t *dst, uint64_t chunk_len, uint64_t n_chunks,
                    uint32_t seed0, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > n_chunks) n_threads = n_chunks ? (int)n_chunks : 1;
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)n_threads);
    uint64_t per = (n_chunks + (uint64_t)n_threads - 1) / (uint64_t)n_threads, done = 0;
    int started = 0;
    for (int k = 0; k < n_threads && done < n_chunks; k++) {
        uint64_t c = per < n_chunks - done ? per : n_chunks - done;
        job_t jb = {pool, word_off, n_words, dst + done * chunk_len, chunk_len, done, c, seed0, 0};
        jobs[k] = jb;
        if (pthread_create(&t[started], NULL, worker, &jobs[k]) != 0) worker(&jobs[k]); else started++;
        done += c;
    }
    for (int k = 0; k < started; k++) pthread_join(t[k], NULL);
    free(t); free(jobs);
}
