/* multi_check.c -- a plain C caller of the multi-GPU entry points (include/zb200.h), no Python, no torch.
 *
 *   gcc -O2 -I include tests/c/multi_check.c -L zlib_b200 -lzb200 -lnccl -Wl,-rpath,... -o multi_check
 *   ./multi_check <libzref.so or libz.so.1> [ndev] [MiB]
 *
 * Compresses a buffer on `ndev` GPUs into ONE zlib stream (zb200_multi_deflate), decodes it with the independent zlib
 * given on the command line (the unmodified reference build oracle/_ref/libzref.so, loaded with dlopen so that its
 * symbols do not clash with the drop-in's), compares every byte, checks the sharded checksums against that library's
 * crc32 / adler32, and inflates a batch of streams on all GPUs.  Exit code 0 = everything agreed.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "zb200.h"

typedef int (*uncompress_fn)(unsigned char *, unsigned long *, const unsigned char *, unsigned long);
typedef int (*compress2_fn)(unsigned char *, unsigned long *, const unsigned char *, unsigned long, int);
typedef unsigned long (*cksum_fn)(unsigned long, const unsigned char *, unsigned int);

static uint64_t rng_state = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd(void) { uint64_t x = rng_state; x ^= x >> 12; x ^= x << 25; x ^= x >> 27; rng_state = x; return x * 0x2545F4914F6CDD1DULL; }

static void fill(unsigned char *p, size_t n)
{
    static const char *words[] = {"deflate", "inflate", "window", "hash", "chain", "block", "stream", "the", "of", "and", "b200", "zlib"};
    size_t o = 0;
    while (o < n) {
        if ((o >> 16) & 1) {                                   /* binary pages: 16-byte records */
            uint32_t rec[4] = {(uint32_t)(o / 16), (uint32_t)(rnd() & 0xff), 0u, (uint32_t)rnd()};
            size_t l = n - o < 16 ? n - o : 16;
            memcpy(p + o, rec, l); o += l;
        } else {
            const char *w = words[rnd() % 12];
            size_t l = strlen(w);
            if (l > n - o) l = n - o;
            memcpy(p + o, w, l); o += l;
            if (o < n) p[o++] = ' ';
        }
    }
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <reference zlib .so> [ndev] [MiB]\n", argv[0]); return 2; }
    void *ref = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
    if (!ref) { fprintf(stderr, "cannot load %s: %s\n", argv[1], dlerror()); return 2; }
    uncompress_fn r_uncompress = (uncompress_fn)dlsym(ref, "uncompress");
    compress2_fn r_compress2 = (compress2_fn)dlsym(ref, "compress2");
    cksum_fn r_crc32 = (cksum_fn)dlsym(ref, "crc32"), r_adler32 = (cksum_fn)dlsym(ref, "adler32");
    if (!r_uncompress || !r_compress2 || !r_crc32 || !r_adler32) { fprintf(stderr, "reference symbols missing\n"); return 2; }
    int ndev = argc > 2 ? atoi(argv[2]) : 0;
    size_t n = (size_t)(argc > 3 ? atoi(argv[3]) : 192) << 20;
    n += 12345;                                                /* not chunk aligned */
    int have = zb200_multi_devices();
    if (have <= 0) { fprintf(stderr, "no device: %s\n", zb200_last_error()); return 1; }
    if (ndev <= 0 || ndev > have) ndev = have;

    unsigned char *src = (unsigned char *)zb200_alloc_pinned(n);
    size_t cap = n + (n >> 12) + (n >> 14) + 11 + 64;
    unsigned char *z = (unsigned char *)zb200_alloc_pinned(cap);
    unsigned char *back = (unsigned char *)malloc(n);
    if (!src || !z || !back) { fprintf(stderr, "allocation failed\n"); return 1; }
    fill(src, n);

    /* ---- one stream from all GPUs ---- */
    for (int level = 1; level <= 6; level += 5) {
        size_t zl = cap;
        uint32_t crc = 0, adler = 0;
        int rc = zb200_multi_deflate(src, n, z, &zl, level, ZB200_WRAP_ZLIB, ndev, &crc, &adler);
        if (rc) { fprintf(stderr, "zb200_multi_deflate level %d -> %d: %s\n", level, rc, zb200_last_error()); return 1; }
        unsigned long bl = n;
        memset(back, 0, n);
        rc = r_uncompress(back, &bl, z, zl);
        if (rc || bl != n || memcmp(back, src, n)) { fprintf(stderr, "reference uncompress of the %d-GPU stream failed: rc %d, %lu of %zu bytes\n", ndev, rc, bl, n); return 1; }
        if (crc != (uint32_t)r_crc32(0, src, (unsigned)n) || adler != (uint32_t)r_adler32(1, src, (unsigned)n)) { fprintf(stderr, "combined checksums differ from the reference\n"); return 1; }
        printf("multi_deflate level %d on %d GPU(s): %zu -> %zu bytes, decoded bit-exact by %s\n", level, ndev, n, zl, argv[1]);
    }
    {   /* empty input and a buffer that is too small */
        size_t zl = cap;
        int rc = zb200_multi_deflate(src, 0, z, &zl, 6, ZB200_WRAP_ZLIB, ndev, NULL, NULL);
        unsigned long bl = 16;
        if (rc || r_uncompress(back, &bl, z, zl) != 0 || bl != 0) { fprintf(stderr, "empty input: rc %d\n", rc); return 1; }
        zl = 1000;
        rc = zb200_multi_deflate(src, n, z, &zl, 1, ZB200_WRAP_ZLIB, ndev, NULL, NULL);
        if (rc != -5) { fprintf(stderr, "short buffer: expected Z_BUF_ERROR, got %d\n", rc); return 1; }
    }
    /* ---- checksums: a slice per GPU, folded ---- */
    {
        uint32_t crc = 0, adler = 0;
        int rc = zb200_multi_checksum(src, n, ndev, &crc, &adler);
        if (rc || crc != (uint32_t)r_crc32(0, src, (unsigned)n) || adler != (uint32_t)r_adler32(1, src, (unsigned)n)) {
            fprintf(stderr, "zb200_multi_checksum: rc %d (%s) crc %08x adler %08x\n", rc, zb200_last_error(), crc, adler);
            return 1;
        }
        printf("multi_checksum on %d GPU(s): crc32 %08x adler32 %08x agree with the reference\n", ndev, crc, adler);
    }
    /* ---- batch inflate of reference-made streams, dealt over the GPUs ---- */
    {
        const size_t ns = 600, sz = 65536;
        uint64_t *so = (uint64_t *)malloc((ns + 1) * 8), *dof = (uint64_t *)malloc((ns + 1) * 8), *dl = (uint64_t *)malloc(ns * 8);
        int32_t *st = (int32_t *)malloc(ns * 4);
        unsigned char *arena = (unsigned char *)malloc(ns * (sz + 1024));
        size_t at = 0;
        for (size_t i = 0; i < ns; i++) {
            unsigned long l = sz + 1024;
            so[i] = at; dof[i] = i * sz;
            if (r_compress2(arena + at, &l, src + i * sz, sz, 6) != 0) { fprintf(stderr, "reference compress2 failed\n"); return 1; }
            at += l;
        }
        so[ns] = at; dof[ns] = ns * sz;
        arena[so[17] + 9] ^= 0x40;                              /* one damaged stream: its status must be the only non-zero one */
        memset(back, 0, ns * sz);
        int rc = zb200_multi_inflate_batch(arena, so, ns, back, dof, dl, st, ZB200_WRAP_ZLIB, ndev);
        if (rc) { fprintf(stderr, "zb200_multi_inflate_batch -> %d: %s\n", rc, zb200_last_error()); return 1; }
        for (size_t i = 0; i < ns; i++) {
            if (i == 17) { if (st[i] == 0) { fprintf(stderr, "damaged stream went unnoticed\n"); return 1; } continue; }
            if (st[i] != 0 || dl[i] != sz || memcmp(back + i * sz, src + i * sz, sz)) { fprintf(stderr, "stream %zu: status %d len %llu\n", i, st[i], (unsigned long long)dl[i]); return 1; }
        }
        printf("multi_inflate_batch on %d GPU(s): %zu reference streams decoded bit-exact, the damaged one reported (%d)\n", ndev, ns, st[17]);
    }
    printf("multi_check ok\n");
    return 0;
}
