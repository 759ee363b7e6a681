/* zapi_oneshot.c -- compress2 / compress / uncompress (qcsrc/compress.c:22-79, qcsrc/uncompr.c:26-61).
 *
 * The reference implements these as deflateInit + deflate(Z_FINISH) + deflateEnd on a
 * stack z_stream.  Here they go straight to the engine's whole-buffer entry points, which
 * is the same work without the streaming shim's staging copy.  Return codes follow the
 * reference: Z_BUF_ERROR when dest is too small, Z_STREAM_ERROR for a bad level,
 * Z_DATA_ERROR for damaged or truncated input, Z_MEM_ERROR when device memory runs out.
 */
#include "../../include/zlib.h"
#include "../../include/zb200.h"

#define ZAPI __attribute__((visibility("default")))

ZAPI int compress2(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen, int level)
{
    size_t out_len;
    int rc;
    if (dest == Z_NULL || destLen == Z_NULL || (source == Z_NULL && sourceLen != 0)) return Z_STREAM_ERROR;
    if (level != Z_DEFAULT_COMPRESSION && (level < 0 || level > 9)) return Z_STREAM_ERROR;   /* deflate.c:265-269 */
    out_len = (size_t)*destLen;
    rc = zb200_deflate(source, (size_t)sourceLen, dest, &out_len, level, ZB200_WRAP_ZLIB, NULL);
    if (rc == Z_OK) *destLen = (uLong)out_len;
    return rc;
}

ZAPI int compress(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen)
{
    return compress2(dest, destLen, source, sourceLen, Z_DEFAULT_COMPRESSION);
}

ZAPI int uncompress(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen)
{
    uint64_t src_off[2], dst_off[2], out_len = 0;
    int32_t status = Z_STREAM_ERROR;
    int rc;
    if (dest == Z_NULL || destLen == Z_NULL || source == Z_NULL) return Z_STREAM_ERROR;
    src_off[0] = 0; src_off[1] = sourceLen;
    dst_off[0] = 0; dst_off[1] = *destLen;
    rc = zb200_inflate_batch(source, src_off, 1, dest, dst_off, &out_len, &status, ZB200_WRAP_ZLIB, NULL);
    if (rc != Z_OK) return rc;
    if (status == Z_OK) *destLen = (uLong)out_len;
    return status;
}
