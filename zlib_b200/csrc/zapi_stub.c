/* TEMPORARY: entry points not implemented yet fail loudly (never fall back). */
#include "../../include/zlib.h"
#include "../../include/zb200.h"
#define ZAPI __attribute__((visibility("default")))
#define STUB(name, proto) ZAPI int name proto { return Z_STREAM_ERROR; }
STUB(zb200_checksum_batch, (const void *b, const uint64_t *o, size_t n, uint32_t *c, uint32_t *a, void *s))
STUB(zb200_deflate_batch, (const void *s, const uint64_t *so, size_t n, void *d, const uint64_t *dof, uint64_t *dl, uint32_t *c, uint32_t *a, int32_t *stt, int l, int w, void *st))
