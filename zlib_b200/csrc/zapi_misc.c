/* zapi_misc.c -- the small fixed-answer entry points of zlib.h (host C).
 * Reference: qcsrc/zutil.c:14-36,133 (z_errmsg, zlibVersion, zlibCompileFlags, zError),
 * qcsrc/crc32.c:205 (get_crc_table), qcsrc/compress.c:75 (compressBound). */
#include "../../include/zlib.h"

#include <stdlib.h>
#define ZAPI __attribute__((visibility("default")))


/* zutil.c:14-24, indexed by 2 - code.  Exported under the reference's name because its own
 * gzio.c reaches for it through the ERR_MSG macro of zutil.h. */
ZAPI const char *const z_errmsg[10] = {
    "need dictionary", "stream end", "", "file error", "stream error",
    "data error", "insufficient memory", "buffer error", "incompatible version", ""};
#define errmsg z_errmsg

/* zutil.c:300-318 */
ZAPI voidpf zcalloc(voidpf opaque, unsigned items, unsigned size) { (void)opaque; return malloc((size_t)items * size); }
ZAPI void zcfree(voidpf opaque, voidpf ptr) { (void)opaque; free(ptr); }

ZAPI const char *zlibVersion(void) { return ZLIB_VERSION; }

ZAPI const char *zError(int err)
{
    int i = 2 - err;
    if (i < 0 || i > 9) i = 9;
    return errmsg[i];
}

/* zutil.c:32-113: size codes for uInt, uLong, voidpf, z_off_t (2 bits each).  No other
 * flag applies to this build (no DEBUG, no ASMV, tables are built at run time on the
 * host for get_crc_table only, gzprintf lives in the reference's own gzio.c). */
ZAPI uLong zlibCompileFlags(void)
{
    uLong f = 0;
    f += sizeof(uInt) == 4 ? 1 : 0;
    f += (sizeof(uLong) == 8 ? 2UL : 1UL) << 2;
    f += (sizeof(voidpf) == 8 ? 2UL : 1UL) << 4;
    f += (sizeof(z_off_t) == 8 ? 2UL : 1UL) << 6;
    return f;
}

/* compress.c:75-79 */
ZAPI uLong compressBound(uLong sourceLen) { return sourceLen + (sourceLen >> 12) + (sourceLen >> 14) + 11; }

/* crc32.c:205: table 0 of the byte-wise CRC, 256 x unsigned long; zip.c/unzip.c key the
 * PKWARE cipher with it (zip.c:886, unzip.c:1176).  Constants, not data arithmetic. */
ZAPI const uLongf *get_crc_table(void)
{
    static uLong table[256];
    static volatile int ready;
    if (!ready) {
        unsigned n, k;
        for (n = 0; n < 256; n++) {
            unsigned long c = n;
            for (k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320UL : c >> 1;
            table[n] = c;
        }
        __sync_synchronize();
        ready = 1;
    }
    return table;
}
