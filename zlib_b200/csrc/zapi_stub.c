/* TEMPORARY: entry points not implemented yet fail loudly (never fall back). */
#include "../../include/zlib.h"
#include "../../include/zb200.h"
#define ZAPI __attribute__((visibility("default")))
#define STUB(name, proto) ZAPI int name proto { return Z_STREAM_ERROR; }
STUB(inflateSyncPoint, (z_streamp s)) STUB(deflate, (z_streamp s, int f)) STUB(deflateEnd, (z_streamp s))
STUB(deflateSetDictionary, (z_streamp s, const Bytef *d, uInt n)) STUB(deflateCopy, (z_streamp d, z_streamp s))
STUB(deflateReset, (z_streamp s)) STUB(deflateParams, (z_streamp s, int l, int t))
STUB(deflateTune, (z_streamp s, int a, int b, int c, int d)) STUB(deflatePrime, (z_streamp s, int b, int v))
STUB(deflateSetHeader, (z_streamp s, gz_headerp h)) STUB(inflate, (z_streamp s, int f)) STUB(inflateEnd, (z_streamp s))
STUB(inflateSetDictionary, (z_streamp s, const Bytef *d, uInt n)) STUB(inflateSync, (z_streamp s))
STUB(inflateCopy, (z_streamp d, z_streamp s)) STUB(inflateReset, (z_streamp s)) STUB(inflatePrime, (z_streamp s, int b, int v))
STUB(inflateGetHeader, (z_streamp s, gz_headerp h))
STUB(deflateInit_, (z_streamp s, int l, const char *v, int sz)) STUB(inflateInit_, (z_streamp s, const char *v, int sz))
STUB(deflateInit2_, (z_streamp s, int l, int m, int w, int ml, int st, const char *v, int sz))
STUB(inflateInit2_, (z_streamp s, int w, const char *v, int sz))
ZAPI uLong deflateBound(z_streamp s, uLong n) { (void)s; return compressBound(n); }
STUB(zb200_checksum_batch, (const void *b, const uint64_t *o, size_t n, uint32_t *c, uint32_t *a, void *s))
STUB(zb200_deflate_batch, (const void *s, const uint64_t *so, size_t n, void *d, const uint64_t *dof, uint64_t *dl, uint32_t *c, uint32_t *a, int32_t *stt, int l, int w, void *st))
