/* zapi_zip.c -- a ZIP32 archive from n files compressed in one GPU batch.
 *
 * The reference writes archives member by member through zip.c: local header, deflate() in
 * 16 KiB steps, sizes and CRC patched in afterwards, central directory at close
 * (qcsrc/zip.c:902-1128, h/zip.h:157-224).  Here all members are compressed by one
 * zb200_deflate_batch() call (raw deflate, windowBits = -15 like zip.c:1005, with the CRC-32
 * of each member from the same pass) and the records are laid out once every size is known.
 * The records are the ones zip.c emits: local file header (30 bytes + name), central
 * directory header (46 bytes + name), end of central directory (22 bytes); the reference's
 * unzip.c checks that local and central headers agree (unzlocal_CheckCurrentFileCoherencyHeader,
 * unzip.c:963-1047), which they do by construction.  ZIP32 only: at most 65535 members and
 * 4 GiB - 1 for every size and offset; larger inputs are refused (Z_STREAM_ERROR).
 */
#include "../../include/zlib.h"
#include "../../include/zb200.h"
#include <stdlib.h>
#include <string.h>

#define ZAPI __attribute__((visibility("default")))

static void put16(unsigned char *p, unsigned v) { p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); }
static void put32(unsigned char *p, unsigned long v)
{
    p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); p[2] = (unsigned char)(v >> 16); p[3] = (unsigned char)(v >> 24);
}

/* Worst-case archive size for these members (every member at compressBound). */
ZAPI size_t zb200_zip_bound(const char *const *names, const uint64_t *src_off, size_t n)
{
    size_t total = 22, i;
    if (names == NULL || src_off == NULL) return 0;
    for (i = 0; i < n; i++) {
        size_t nl = strlen(names[i]);
        total += 30 + nl + 46 + nl + (size_t)compressBound((uLong)(src_off[i + 1] - src_off[i])) + 16;
    }
    return total;
}

ZAPI int zb200_zip_build(const char *const *names, const void *src, const uint64_t *src_off, size_t n, int level,
                         uint32_t dos_datetime, void *dst, size_t *dst_len)
{
    uint64_t *slot_off = NULL, *clen = NULL;
    uint32_t *crc = NULL;
    int32_t *status = NULL;
    unsigned char *arena = NULL, *out = (unsigned char *)dst, *cd;
    uint64_t pos = 0, cd_size = 0, need;
    size_t i;
    int rc = Z_OK;

    if (names == NULL || src_off == NULL || dst == NULL || dst_len == NULL || n > 65535u) return Z_STREAM_ERROR;
    if (level != Z_DEFAULT_COMPRESSION && (level < 0 || level > 9)) return Z_STREAM_ERROR;
    for (i = 0; i < n; i++) {
        if (names[i] == NULL || strlen(names[i]) > 65535u || src_off[i + 1] < src_off[i] ||
            src_off[i + 1] - src_off[i] >= 0xffffffffull) return Z_STREAM_ERROR;
    }
    slot_off = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    clen = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    crc = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    status = (int32_t *)malloc((n + 1) * sizeof(int32_t));
    if (!slot_off || !clen || !crc || !status) { rc = Z_MEM_ERROR; goto done; }
    slot_off[0] = 0;
    for (i = 0; i < n; i++) slot_off[i + 1] = slot_off[i] + compressBound((uLong)(src_off[i + 1] - src_off[i])) + 16;
    arena = (unsigned char *)zb200_alloc_pinned((size_t)slot_off[n] + 16);       /* pinned: the D2H copies run at link speed */
    if (!arena) { rc = Z_MEM_ERROR; goto done; }
    rc = zb200_deflate_batch(src, src_off, n, arena, slot_off, clen, crc, NULL, status, level, ZB200_WRAP_RAW, NULL);
    if (rc != Z_OK) goto done;
    need = 22;
    for (i = 0; i < n; i++) {
        if (status[i] != Z_OK) { rc = status[i]; goto done; }
        need += 30 + 46 + 2 * strlen(names[i]) + clen[i];
    }
    if (need >= 0xffffffffull) { rc = Z_STREAM_ERROR; goto done; }               /* ZIP32 */
    if (need > *dst_len) { *dst_len = (size_t)need; rc = Z_BUF_ERROR; goto done; }

    /* local headers + data (zip.c:969-1032), central directory built behind them (zip.c:940-967) */
    for (i = 0; i < n; i++) pos += 30 + strlen(names[i]) + clen[i];
    cd = out + pos;
    pos = 0;
    for (i = 0; i < n; i++) {
        const size_t nl = strlen(names[i]);
        const unsigned long usize = (unsigned long)(src_off[i + 1] - src_off[i]);
        unsigned char *lh = out + pos, *ch = cd + cd_size;
        put32(lh, 0x04034b50ul); put16(lh + 4, 20); put16(lh + 6, 0); put16(lh + 8, Z_DEFLATED);
        put32(lh + 10, dos_datetime); put32(lh + 14, crc[i]); put32(lh + 18, (unsigned long)clen[i]); put32(lh + 22, usize);
        put16(lh + 26, (unsigned)nl); put16(lh + 28, 0);
        memcpy(lh + 30, names[i], nl);
        memcpy(lh + 30 + nl, arena + slot_off[i], (size_t)clen[i]);
        put32(ch, 0x02014b50ul); put16(ch + 4, 0); put16(ch + 6, 20); put16(ch + 8, 0); put16(ch + 10, Z_DEFLATED);
        put32(ch + 12, dos_datetime); put32(ch + 16, crc[i]); put32(ch + 20, (unsigned long)clen[i]); put32(ch + 24, usize);
        put16(ch + 28, (unsigned)nl); put16(ch + 30, 0); put16(ch + 32, 0); put16(ch + 34, 0); put16(ch + 36, 0);
        put32(ch + 38, 0); put32(ch + 42, (unsigned long)pos);
        memcpy(ch + 46, names[i], nl);
        pos += 30 + nl + clen[i];
        cd_size += 46 + nl;
    }
    {
        unsigned char *e = cd + cd_size;                                         /* zip.c:1203-1240 */
        put32(e, 0x06054b50ul); put16(e + 4, 0); put16(e + 6, 0); put16(e + 8, (unsigned)n); put16(e + 10, (unsigned)n);
        put32(e + 12, (unsigned long)cd_size); put32(e + 16, (unsigned long)pos); put16(e + 20, 0);
    }
    *dst_len = (size_t)(pos + cd_size + 22);
done:
    if (arena) zb200_free_pinned(arena);
    free(slot_off); free(clen); free(crc); free(status);
    return rc;
}
