/* zconf.h -- type configuration for the zb200 drop-in (LP64 Linux hosts only).
 *
 * Binary contract = the reference's h/zconf.h:260-302 on an LP64 build:
 *   Byte 8 bit, uInt 32 bit, uLong 64 bit, voidpf = void*, z_off_t = long.
 * The reference's 16-bit / far-pointer / K&R configuration space is not
 * reproduced: this library only exists where a B200 does.
 */
#ifndef ZB200_ZCONF_H
#define ZB200_ZCONF_H

#include <stddef.h>
#include <sys/types.h>

#define MAX_MEM_LEVEL 9          /* h/zconf.h:134-140 */
#define MAX_WBITS     15         /* h/zconf.h:147-149: 32 KiB window */

#ifndef OF
#  define OF(args) args
#endif
#define ZEXTERN extern
#define ZEXPORT
#define ZEXPORTVA
#define FAR
#ifndef z_const
#  define z_const const
#endif

typedef unsigned char  Byte;
typedef unsigned int   uInt;
typedef unsigned long  uLong;

typedef Byte   Bytef;
typedef char   charf;
typedef int    intf;
typedef uInt   uIntf;
typedef uLong  uLongf;

typedef void const *voidpc;
typedef void       *voidpf;
typedef void       *voidp;

#ifndef z_off_t
#  define z_off_t long
#endif
#ifndef SEEK_SET
#  define SEEK_SET 0
#  define SEEK_CUR 1
#  define SEEK_END 2
#endif

#endif /* ZB200_ZCONF_H */
