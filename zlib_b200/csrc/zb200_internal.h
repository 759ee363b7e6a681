/* zb200_internal.h -- C-linkage seam between the host-C zlib.h shim (zapi_*.c) and the
 * CUDA engine (zb_*.cu).  Hidden visibility: not part of the exported ABI. */
#ifndef ZB200_INTERNAL_H
#define ZB200_INTERNAL_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Forces the empty stored block at the end of a non-final shard even when the shard
 * already ends on a byte boundary (Z_SYNC_FLUSH / Z_FULL_FLUSH marker, deflate.c:808-819). */
#define ZB200I_DEFLATE_FORCE_MARK 8
/* deflateInit2's windowBits below 15 (bits 12..15 of flags, 0 = 15): no match reaches further back than the reference's
 * MAX_DIST = (1 << windowBits) - 262 (h/deflate.h:276), so a decoder that sizes its window from CINFO can follow. */
#define ZB200I_DEFLATE_WBITS(w) (((w) & 15) << 12)

/* A raw shard with pending bits at either end (deflatePrime, Z_PARTIAL_FLUSH): see zb_deflate.cu. */
int zb200i_deflate_shard_bits(const void *src, size_t src_len, const void *dict, size_t dict_len, void *dst, size_t *dst_len,
                              int level, int flags, uint32_t prime_bits, uint32_t prime_val, int partial_end,
                              uint32_t *tail, uint32_t *crc, uint32_t *adler);

/* ---- one resumable inflate stream (inflate.c state machine on the device) ---- */
typedef struct zb200i_inflater zb200i_inflater;

#define ZB200I_NEED_INPUT  100
#define ZB200I_NEED_OUTPUT 101

int  zb200i_inflate_open(zb200i_inflater **h, int wrap);
int  zb200i_inflate_reset(zb200i_inflater *h, int wrap);
void zb200i_inflate_close(zb200i_inflater *h);
int  zb200i_inflate_clone(zb200i_inflater **dst, const zb200i_inflater *src);
/* Feeds `in_len` new bytes, produces at most out_cap bytes into `out` (host memory).
 * *in_used: how many of the new bytes the stream has taken (bytes it could not decode yet
 * are kept inside the handle).  status: Z_STREAM_END, Z_NEED_DICT, Z_DATA_ERROR,
 * ZB200I_NEED_INPUT, ZB200I_NEED_OUTPUT.  *check = running Adler-32 (zlib) of the output,
 * or the DICTID when status is Z_NEED_DICT. */
int  zb200i_inflate_run(zb200i_inflater *h, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                        size_t *in_used, size_t *out_len, int *status, int *msg, uint32_t *check);
int  zb200i_inflate_set_dict(zb200i_inflater *h, const uint8_t *dict, size_t n);
/* Drops buffered input and restarts block decoding (after inflateSync found a marker). */
int  zb200i_inflate_resync(zb200i_inflater *h);
/* The same when the marker ends `drop` bytes into the input the stream already holds: those bytes go, the rest stays. */
int  zb200i_inflate_resync_keep(zb200i_inflater *h, size_t drop);
/* inflatePrime: bits (<= 16) of value go ahead of the next input byte; only with no input pending. */
int  zb200i_inflate_prime(zb200i_inflater *h, int bits, int value);
int  zb200i_inflate_mode(const zb200i_inflater *h);         /* InfMode of the device state */
size_t zb200i_inflate_pending_input(const zb200i_inflater *h);
const uint8_t *zb200i_inflate_pending_bytes(const zb200i_inflater *h);

/* One call holding the whole stream and the whole output buffer: 0 = decoded by the segment-parallel decoder
 * (wrap 0 raw / 1 zlib / 2 gzip / 3 either; *check = Adler-32 of the output, CRC-32 for gzip), 1 = not taken, use the streaming decoder. */
int  zb200i_inflate_try_parallel(const uint8_t *in, size_t in_len, uint8_t *out, size_t cap, int wrap,
                                 size_t *in_used, size_t *out_len, uint32_t *check);

const char *zb200i_inflate_msg(int msg);

#ifdef __cplusplus
}
#endif
#endif
