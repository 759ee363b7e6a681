// zb_inflate.cuh -- state shared between the inflate kernels and the host shim.
#pragma once
#include "zb_common.cuh"

namespace zb {

// Where a stream stands (the reference's inflate_mode, h/inflate.h:20-51, collapsed to the
// points at which this decoder can pause).
enum InfMode { kModeHead = 0, kModeBlock, kModeStored, kModeCodes, kModeCopy, kModeTrailer, kModeDone, kModeBad, kModeDict };

// strm->msg texts of the reference (inflate.c, inffast.c), by index.
enum InfMsg { kMsgNone = 0, kMsgHeader, kMsgMethod, kMsgWindow, kMsgNeedDict, kMsgBlockType, kMsgStored,
              kMsgTooMany, kMsgCodeLens, kMsgRepeat, kMsgLitSet, kMsgDistSet, kMsgBadLit, kMsgBadDist,
              kMsgFar, kMsgCheck, kMsgLength, kMsgFlags, kMsgHcrc };

constexpr int kWrapAuto = 3;             // windowBits + 32: zlib or gzip, detected from the header (inflate.c:596)
constexpr uint32_t kFlagGzipTrailer = 1;   // InfState::flags: crc / isize hold a gzip trailer that is still to be checked

// Per-call outcome codes of the streaming kernel (beyond zlib's own codes).
enum { kNeedInput = 100, kNeedOutput = 101 };

struct InfState {
    int32_t  mode, last, msg, wrap;
    uint32_t bit_off;                  // bits of the first input byte that are already consumed
    uint32_t stored_left, copy_len, copy_dist;
    uint32_t s1, s2;                   // running Adler-32 of the output
    uint32_t crc, isize;               // running CRC-32 / length (gzip)
    int32_t  nlen, ndist;              // sizes of the code in force (for table rebuild on resume)
    uint64_t hist;                     // valid bytes of history in front of the output pointer
    uint64_t total_out;
    uint32_t dict_id, flags;
    uint8_t  lens[320];
};

struct InfCallResult {                 // written by the streaming kernel
    uint64_t out_len, in_used;
    int32_t  status, msg;
};

}  // namespace zb
