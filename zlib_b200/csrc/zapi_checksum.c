/* zapi_checksum.c -- zlib.h checksum entry points over the zb200 engine (host C).
 *
 * crc32 / adler32 hand every data byte to the GPU (zb200_checksum, K5/K6); what
 * stays on the host is scalar bookkeeping on three 32-bit words: folding the
 * caller's running value into the GPU's result, and the two _combine functions,
 * which the reference also computes on scalars (qcsrc/crc32.c:370-423,
 * qcsrc/adler32.c:128-149).
 */
#include "../../include/zlib.h"
#include "../../include/zb200.h"
#include <stdio.h>
#include <stdlib.h>

#define ZAPI __attribute__((visibility("default")))

#define BASE 65521UL

static void die(const char *what)
{
    fprintf(stderr, "zb200: %s failed: %s (no CPU path exists)\n", what, zb200_last_error());
    abort();
}

/* a(x)*b(x) mod P; reflected bit order, bit 31 = x^0 */
static unsigned gf2_mul(unsigned a, unsigned b)
{
    unsigned r = 0;
    int i;
    for (i = 31; i >= 0; i--) {
        if ((b >> i) & 1u) r ^= a;
        a = (a & 1u) ? (a >> 1) ^ 0xEDB88320u : a >> 1;
    }
    return r;
}

/* qcsrc/crc32.c:370: CRC of A||B from CRC(A), CRC(B), len(B).  The pre/post inversions
 * cancel, leaving crc1 * x^(8*len2) mod P xor crc2. */
ZAPI uLong crc32_combine(uLong crc1, uLong crc2, z_off_t len2)
{
    unsigned p = 0x00800000u, acc = 0x80000000u;
    unsigned long n;
    if (len2 <= 0) return crc1;            /* crc32.c:383; a negative length never terminates there */
    for (n = (unsigned long)len2; n; n >>= 1) {
        if (n & 1) acc = gf2_mul(acc, p);
        p = gf2_mul(p, p);
    }
    return (uLong)(gf2_mul((unsigned)crc1, acc) ^ (unsigned)crc2);
}

/* qcsrc/adler32.c:128-149, including the '>' comparisons of the final folds. */
ZAPI uLong adler32_combine(uLong adler1, uLong adler2, z_off_t len2)
{
    unsigned long sum1, sum2;
    unsigned rem = (unsigned)(len2 % (z_off_t)BASE);
    sum1 = adler1 & 0xffff;
    sum2 = (rem * sum1) % BASE;
    sum1 += (adler2 & 0xffff) + BASE - 1;
    sum2 += ((adler1 >> 16) & 0xffff) + ((adler2 >> 16) & 0xffff) + BASE - rem;
    if (sum1 > BASE) sum1 -= BASE;
    if (sum1 > BASE) sum1 -= BASE;
    if (sum2 > (BASE << 1)) sum2 -= (BASE << 1);
    if (sum2 > BASE) sum2 -= BASE;
    return sum1 | (sum2 << 16);
}

/* qcsrc/crc32.c:219 */
ZAPI uLong crc32(uLong crc, const Bytef *buf, uInt len)
{
    uint32_t c0 = 0;
    if (buf == Z_NULL) return 0UL;
    if (len == 0) return crc & 0xffffffffUL;
    if (zb200_checksum(buf, len, &c0, NULL, NULL) != Z_OK) die("crc32");
    if ((crc & 0xffffffffUL) == 0) return c0;
    return crc32_combine(crc & 0xffffffffUL, c0, (z_off_t)len);
}

/* qcsrc/adler32.c:57.  The GPU returns adler32(1, buf); the running value is folded in
 * exactly as the reference's arithmetic does, including its single-subtraction path for
 * len == 1 (adler32.c:70-78) which matters for non-canonical inputs. */
ZAPI uLong adler32(uLong adler, const Bytef *buf, uInt len)
{
    uint32_t g = 1;
    unsigned long s1 = adler & 0xffff, s2 = (adler >> 16) & 0xffff, a, b;
    if (len == 1 && buf != Z_NULL) {
        if (zb200_checksum(buf, 1, NULL, &g, NULL) != Z_OK) die("adler32");
        s1 += (g & 0xffff) - 1;            /* the byte value */
        if (s1 >= BASE) s1 -= BASE;
        s2 += s1;
        if (s2 >= BASE) s2 -= BASE;
        return s1 | (s2 << 16);
    }
    if (buf == Z_NULL) return 1UL;
    if (len != 0 && zb200_checksum(buf, len, NULL, &g, NULL) != Z_OK) die("adler32");
    a = ((g & 0xffff) + BASE - 1) % BASE;                   /* sum of bytes */
    b = ((g >> 16) + BASE - (unsigned long)len % BASE) % BASE;   /* position-weighted sum */
    s2 = (s2 + ((unsigned long)len % BASE) * s1 + b) % BASE;
    s1 = (s1 + a) % BASE;
    return s1 | (s2 << 16);
}
