/* zbsynth.c -- synthetic corpora for benchmarks and tests.
 *
 * Built into its own tiny library (zlib_b200/libzbsynth.so), NOT into libzb200.so: the reference arm of bench.py
 * generates its input with it and must not map the product library.
 * Workload generator only; no codec arithmetic.  Every 64 KiB page is generated from
 * its own xorshift64* state derived from (seed, page index), so the output is
 * deterministic, position-addressable and can be produced by several host threads.
 *
 *   kind 0  T  text: 4096-word vocabulary (2..10 lowercase letters), word index
 *              floor(u^3 * 4096) for a Zipf-like law, separator ' ' or '\n' (1 in 12)
 *   kind 1  M  mixed: T with every odd page replaced by 16-byte records
 *              {u32 counter, u32 small(0..255), u32 0, u32 random}
 *   kind 2     raw generator output (incompressible)
 */
#include <stddef.h>
#include <stdint.h>
#include <pthread.h>
#include <string.h>
#include <unistd.h>

#define ZAPI __attribute__((visibility("default")))
#define PAGE 65536u
#define VOCAB 4096

static inline uint64_t xs64(uint64_t *s)
{
    uint64_t x = *s;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    *s = x;
    return x * 0x2545F4914F6CDD1DULL;
}

typedef struct { char w[VOCAB][10]; uint8_t len[VOCAB]; } vocab_t;

static void make_vocab(vocab_t *v, uint64_t seed)
{
    uint64_t s = 0x9E3779B97F4A7C15ULL ^ (seed * 0xD6E8FEB86659FD93ULL);
    if (s == 0) s = 1;
    for (int i = 0; i < VOCAB; i++) {
        int n = 2 + (int)(xs64(&s) % 9);
        v->len[i] = (uint8_t)n;
        for (int k = 0; k < n; k++) v->w[i][k] = (char)('a' + xs64(&s) % 26);
    }
}

static void text_page(const vocab_t *v, uint8_t *dst, size_t n, uint64_t s)
{
    size_t o = 0;
    while (o < n) {
        uint64_t r = xs64(&s);
        double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
        int idx = (int)(u * u * u * VOCAB);
        if (idx >= VOCAB) idx = VOCAB - 1;
        size_t l = v->len[idx];
        for (size_t k = 0; k < l && o < n; k++) dst[o++] = (uint8_t)v->w[idx][k];
        if (o < n) dst[o++] = ((r & 0xff) % 12 == 0) ? '\n' : ' ';
    }
}

static void record_page(uint8_t *dst, size_t n, uint64_t s, uint64_t page)
{
    uint32_t counter = (uint32_t)(page * (PAGE / 16));
    size_t o = 0;
    while (o < n) {
        uint64_t r = xs64(&s);
        uint32_t rec[4] = {counter++, (uint32_t)(r & 0xff), 0u, (uint32_t)(r >> 32)};
        size_t l = n - o < 16 ? n - o : 16;
        memcpy(dst + o, rec, l);
        o += l;
    }
}

static void noise_page(uint8_t *dst, size_t n, uint64_t s)
{
    size_t o = 0;
    while (o < n) {
        uint64_t r = xs64(&s);
        size_t l = n - o < 8 ? n - o : 8;
        memcpy(dst + o, &r, l);
        o += l;
    }
}

typedef struct { uint8_t *dst; size_t len; int kind; uint64_t seed; const vocab_t *v; int tid, nthreads; uint64_t page0; } job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    size_t pages = (j->len + PAGE - 1) / PAGE;
    for (size_t q = (size_t)j->tid; q < pages; q += (size_t)j->nthreads) {
        size_t off = q * PAGE, n = j->len - off < PAGE ? j->len - off : PAGE;
        uint64_t p = j->page0 + q;                               /* page index within the whole corpus */
        uint64_t s = 0x9E3779B97F4A7C15ULL ^ (j->seed * 0xBF58476D1CE4E5B9ULL) ^ ((p + 1) * 0x94D049BB133111EBULL);
        if (s == 0) s = 1;
        xs64(&s); xs64(&s);
        if (j->kind == 2) noise_page(j->dst + off, n, s);
        else if (j->kind == 1 && (p & 1)) record_page(j->dst + off, n, s, p);
        else text_page(j->v, j->dst + off, n, s);
    }
    return NULL;
}

/* Bytes [offset, offset + len) of the corpus (kind, seed); offset must be a multiple of 64 KiB (the page size). */
ZAPI int zbsynth_fill(void *host_dst, size_t len, int kind, uint64_t seed, uint64_t offset)
{
    if (offset % PAGE) return -1;
    vocab_t local;
    make_vocab(&local, seed);
    long nc = sysconf(_SC_NPROCESSORS_ONLN);
    int nt = nc < 1 ? 1 : nc > 32 ? 32 : (int)nc;
    if (len < 4 * PAGE) nt = 1;
    pthread_t th[32];
    job_t jobs[32];
    for (int t = 0; t < nt; t++) {
        jobs[t] = (job_t){(uint8_t *)host_dst, len, kind, seed, &local, t, nt, offset / PAGE};
        if (t > 0) pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    worker(&jobs[0]);
    for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
    return 0;
}
