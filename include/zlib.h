/* zlib.h -- the zlib 1.2.3 C ABI, as served by the zb200 B200-native engine.
 *
 * This header is the DROP-IN BOUNDARY.  It declares, with the reference's
 * names, argument meaning, struct layout and error codes, the entry points of
 * the reference's h/zlib.h (ChrisHird/ZLIB, zlib 1.2.3).  Each group cites the
 * reference prototype it replaces.  Programs written against the reference
 * header (qcsrc/compress.c, uncompr.c, zip.c, unzip.c, gzio.c, example.c)
 * compile unchanged against this one.
 *
 * Behind these symbols every byte of codec / checksum arithmetic runs in
 * hand-written sm_100a CUDA kernels (zlib_b200/csrc/).  There is no CPU
 * implementation: without a usable B200 the calls fail (Z_STREAM_ERROR for
 * stream calls; checksums abort with a message), they never fall back.
 *
 * Buffers may be ordinary host memory, CUDA pinned memory, or device memory
 * (detected per call); see include/zb200.h for the batch / device-resident
 * extensions.
 */
#ifndef ZB200_ZLIB_H
#define ZB200_ZLIB_H

#include "zconf.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ZLIB_VERSION "1.2.3"     /* h/zlib.h:40 -- version[0] must stay '1' */
#define ZLIB_VERNUM  0x1230      /* h/zlib.h:41 */

typedef voidpf (*alloc_func)(voidpf opaque, uInt items, uInt size);   /* h/zlib.h:77 */
typedef void   (*free_func)(voidpf opaque, voidpf address);           /* h/zlib.h:78 */

struct internal_state;

/* h/zlib.h:82-101.  sizeof == 112; offsets 0,8,16,24,32,40,48,56,64,72,80,88,96,104. */
typedef struct z_stream_s {
    Bytef    *next_in;
    uInt      avail_in;
    uLong     total_in;

    Bytef    *next_out;
    uInt      avail_out;
    uLong     total_out;

    char     *msg;
    struct internal_state *state;

    alloc_func zalloc;
    free_func  zfree;
    voidpf     opaque;

    int       data_type;
    uLong     adler;
    uLong     reserved;
} z_stream;

typedef z_stream *z_streamp;

/* h/zlib.h:108-124 */
typedef struct gz_header_s {
    int     text;
    uLong   time;
    int     xflags;
    int     os;
    Bytef  *extra;
    uInt    extra_len;
    uInt    extra_max;
    Bytef  *name;
    uInt    name_max;
    Bytef  *comment;
    uInt    comm_max;
    int     hcrc;
    int     done;
} gz_header;

typedef gz_header *gz_headerp;

/* flush values, h/zlib.h:162-167 */
#define Z_NO_FLUSH      0
#define Z_PARTIAL_FLUSH 1
#define Z_SYNC_FLUSH    2
#define Z_FULL_FLUSH    3
#define Z_FINISH        4
#define Z_BLOCK         5

/* return codes, h/zlib.h:170-178 */
#define Z_OK            0
#define Z_STREAM_END    1
#define Z_NEED_DICT     2
#define Z_ERRNO        (-1)
#define Z_STREAM_ERROR (-2)
#define Z_DATA_ERROR   (-3)
#define Z_MEM_ERROR    (-4)
#define Z_BUF_ERROR    (-5)
#define Z_VERSION_ERROR (-6)

/* levels, h/zlib.h:183-186 */
#define Z_NO_COMPRESSION         0
#define Z_BEST_SPEED             1
#define Z_BEST_COMPRESSION       9
#define Z_DEFAULT_COMPRESSION  (-1)

/* strategies, h/zlib.h:189-193 */
#define Z_FILTERED            1
#define Z_HUFFMAN_ONLY        2
#define Z_RLE                 3
#define Z_FIXED               4
#define Z_DEFAULT_STRATEGY    0

/* data_type, h/zlib.h:196-199 */
#define Z_BINARY   0
#define Z_TEXT     1
#define Z_ASCII    Z_TEXT
#define Z_UNKNOWN  2

#define Z_DEFLATED   8           /* h/zlib.h:202 */
#define Z_NULL  0                /* h/zlib.h:205 */

#define zlib_version zlibVersion()

/* ---- utility (h/zlib.h:212, 1046, 1336, 1347-1351) ---- */
ZEXTERN const char   *zlibVersion(void);
ZEXTERN uLong         zlibCompileFlags(void);
ZEXTERN const char   *zError(int);
ZEXTERN const uLongf *get_crc_table(void);
ZEXTERN int           inflateSyncPoint(z_streamp);

/* ---- streaming compressor (h/zlib.h:242-349, 457-634); qcsrc/deflate.c ---- */
ZEXTERN int   deflate(z_streamp strm, int flush);
ZEXTERN int   deflateEnd(z_streamp strm);
ZEXTERN int   deflateSetDictionary(z_streamp strm, const Bytef *dictionary, uInt dictLength);
ZEXTERN int   deflateCopy(z_streamp dest, z_streamp source);
ZEXTERN int   deflateReset(z_streamp strm);
ZEXTERN int   deflateParams(z_streamp strm, int level, int strategy);
ZEXTERN int   deflateTune(z_streamp strm, int good_length, int max_lazy, int nice_length, int max_chain);
ZEXTERN uLong deflateBound(z_streamp strm, uLong sourceLen);
ZEXTERN int   deflatePrime(z_streamp strm, int bits, int value);
ZEXTERN int   deflateSetHeader(z_streamp strm, gz_headerp head);

/* ---- streaming decompressor (h/zlib.h:374-454, 636-800); qcsrc/inflate.c ---- */
ZEXTERN int   inflate(z_streamp strm, int flush);
ZEXTERN int   inflateEnd(z_streamp strm);
ZEXTERN int   inflateSetDictionary(z_streamp strm, const Bytef *dictionary, uInt dictLength);
ZEXTERN int   inflateSync(z_streamp strm);
ZEXTERN int   inflateCopy(z_streamp dest, z_streamp source);
ZEXTERN int   inflateReset(z_streamp strm);
ZEXTERN int   inflatePrime(z_streamp strm, int bits, int value);
ZEXTERN int   inflateGetHeader(z_streamp strm, gz_headerp head);

/* ---- one-shot helpers (h/zlib.h:1059-1115); qcsrc/compress.c, uncompr.c ---- */
ZEXTERN int   compress(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen);
ZEXTERN int   compress2(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen, int level);
ZEXTERN uLong compressBound(uLong sourceLen);
ZEXTERN int   uncompress(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen);

/* ---- checksums (h/zlib.h:1260-1309); qcsrc/adler32.c, crc32.c ---- */
ZEXTERN uLong adler32(uLong adler, const Bytef *buf, uInt len);
ZEXTERN uLong adler32_combine(uLong adler1, uLong adler2, z_off_t len2);
ZEXTERN uLong crc32(uLong crc, const Bytef *buf, uInt len);
ZEXTERN uLong crc32_combine(uLong crc1, uLong crc2, z_off_t len2);

/* ---- versioned initialisers behind the macros (h/zlib.h:1317-1342) ---- */
ZEXTERN int deflateInit_(z_streamp strm, int level, const char *version, int stream_size);
ZEXTERN int inflateInit_(z_streamp strm, const char *version, int stream_size);
ZEXTERN int deflateInit2_(z_streamp strm, int level, int method, int windowBits, int memLevel,
                          int strategy, const char *version, int stream_size);
ZEXTERN int inflateInit2_(z_streamp strm, int windowBits, const char *version, int stream_size);

#define deflateInit(strm, level) \
        deflateInit_((strm), (level), ZLIB_VERSION, (int)sizeof(z_stream))
#define inflateInit(strm) \
        inflateInit_((strm), ZLIB_VERSION, (int)sizeof(z_stream))
#define deflateInit2(strm, level, method, windowBits, memLevel, strategy) \
        deflateInit2_((strm), (level), (method), (windowBits), (memLevel), \
                      (strategy), ZLIB_VERSION, (int)sizeof(z_stream))
#define inflateInit2(strm, windowBits) \
        inflateInit2_((strm), (windowBits), ZLIB_VERSION, (int)sizeof(z_stream))

/* ---- outside the hot path ----
 * inflateBack* (qcsrc/infback.c) and gz* (qcsrc/gzio.c) are callers/variants of
 * the path, not part of it.  The gz* prototypes are kept so that the reference's
 * own gzio.c and example.c compile and link against this library unchanged. */
typedef voidp gzFile;
typedef unsigned (*in_func)(void *, unsigned char **);
typedef int (*out_func)(void *, unsigned char *, unsigned);

ZEXTERN gzFile gzopen(const char *path, const char *mode);
ZEXTERN gzFile gzdopen(int fd, const char *mode);
ZEXTERN int    gzsetparams(gzFile file, int level, int strategy);
ZEXTERN int    gzread(gzFile file, voidp buf, unsigned len);
ZEXTERN int    gzwrite(gzFile file, voidpc buf, unsigned len);
ZEXTERN int    gzprintf(gzFile file, const char *format, ...);
ZEXTERN int    gzputs(gzFile file, const char *s);
ZEXTERN char  *gzgets(gzFile file, char *buf, int len);
ZEXTERN int    gzputc(gzFile file, int c);
ZEXTERN int    gzgetc(gzFile file);
ZEXTERN int    gzungetc(int c, gzFile file);
ZEXTERN int    gzflush(gzFile file, int flush);
ZEXTERN z_off_t gzseek(gzFile file, z_off_t offset, int whence);
ZEXTERN int    gzrewind(gzFile file);
ZEXTERN z_off_t gztell(gzFile file);
ZEXTERN int    gzeof(gzFile file);
ZEXTERN int    gzdirect(gzFile file);
ZEXTERN int    gzclose(gzFile file);
ZEXTERN const char *gzerror(gzFile file, int *errnum);
ZEXTERN void   gzclearerr(gzFile file);

#ifdef __cplusplus
}
#endif

#endif /* ZB200_ZLIB_H */
