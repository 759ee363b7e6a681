// zb_zip.cu -- a ZIP32 archive from n files compressed in one GPU batch, assembled on the device.
//
// The reference writes archives member by member through zip.c: local header, deflate() in 16 KiB steps, sizes and
// CRC patched in afterwards, central directory at close (qcsrc/zip.c:902-1128, h/zip.h:157-224).  Here all members
// are compressed by one zb200_deflate_batch() call (raw deflate, windowBits = -15 like zip.c:1005, with the CRC-32 of
// each member from the same pass) into device slots; once every size is known the records are laid out by prefix sum
// and one copy kernel moves headers and member data to their final places, so the archive leaves the GPU as ONE
// contiguous D2H copy.  The records are the ones zip.c emits: local file header (30 bytes + name), central directory
// header (46 bytes + name), end of central directory (22 bytes); the reference's unzip.c checks that local and
// central headers agree (unzlocal_CheckCurrentFileCoherencyHeader, unzip.c:963-1047), which they do by construction.
//
// A *segment* is the run of [local header | name | data] records of some members; segments built on different GPUs
// concatenate (BASELINE config 5: files are dealt to the ranks, the root appends the central directory).
// ZIP32 only: at most 65535 members and 4 GiB - 1 for every size and offset; larger inputs are refused.
#include "zb_common.cuh"
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace zb {

struct Piece { const uint8_t* src; uint8_t* dst; uint64_t len; };
constexpr uint32_t kPieceBytes = 256u << 10;
constexpr int kGatherThreads = 256;

// One CTA per piece: dst-aligned 4-byte stores fed by two aligned loads and a funnel shift (source and destination
// are misaligned against each other in general), four words per thread in flight; ragged ends byte by byte.
__global__ void __launch_bounds__(kGatherThreads) k_zip_gather(const Piece* __restrict__ pieces)
{
    const Piece pc = pieces[blockIdx.x];
    const uint8_t* s = pc.src;
    uint8_t* d = pc.dst;
    uint64_t n = pc.len;
    const uint32_t head = (uint32_t)min((uint64_t)((4 - ((uintptr_t)d & 3)) & 3), n);
    if (threadIdx.x < head) d[threadIdx.x] = s[threadIdx.x];
    s += head; d += head; n -= head;
    const uint64_t nw = n >> 2;
    const uint32_t sh = (uint32_t)((uintptr_t)s & 3) * 8;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>((uintptr_t)s & ~(uintptr_t)3);
    uint32_t* dw = reinterpret_cast<uint32_t*>(d);
    uint64_t i = threadIdx.x;
    for (; i + 3 * kGatherThreads < nw; i += 4 * kGatherThreads) {
        uint32_t a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { a[k] = sw[i + k * kGatherThreads]; b[k] = sh ? sw[i + k * kGatherThreads + 1] : 0u; }
#pragma unroll
        for (int k = 0; k < 4; k++) dw[i + k * kGatherThreads] = __funnelshift_r(a[k], b[k], sh);
    }
    for (; i < nw; i += kGatherThreads) dw[i] = __funnelshift_r(sw[i], sh ? sw[i + 1] : 0u, sh);
    const uint32_t tail = (uint32_t)(n & 3);
    if (threadIdx.x < tail) d[nw * 4 + threadIdx.x] = s[nw * 4 + threadIdx.x];
}

static void put16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void put32(uint8_t* p, uint64_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

}  // namespace zb

using namespace zb;

extern "C" unsigned long compressBound(unsigned long sourceLen);   // zapi_oneshot.c

// Worst-case archive size for these members (every member at compressBound).
ZB_API size_t zb200_zip_bound(const char* const* names, const uint64_t* src_off, size_t n)
{
    size_t total = 22;
    if (names == nullptr || src_off == nullptr) return 0;
    for (size_t i = 0; i < n; i++) {
        const size_t nl = strlen(names[i]);
        total += 30 + nl + 46 + nl + (size_t)compressBound((unsigned long)(src_off[i + 1] - src_off[i])) + 16;
    }
    return total;
}

ZB_API int zb200_zip_segment(const char* const* names, const void* src, const uint64_t* src_off, size_t n, int level,
                             uint32_t dos_datetime, void* dst, size_t* dst_len, zb200_zip_member* members)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (names == nullptr || src_off == nullptr || dst == nullptr || dst_len == nullptr || members == nullptr || n > 65535u ||
        (level != -1 && (level < 0 || level > 9))) {
        set_error("zb200_zip_segment: bad argument");
        return ZB_STREAM_ERROR;
    }
    for (size_t i = 0; i < n; i++) {
        if (names[i] == nullptr || strlen(names[i]) > 65535u || src_off[i + 1] < src_off[i] || src_off[i + 1] - src_off[i] >= 0xffffffffull) {
            set_error("zb200_zip_segment: member %zu is outside ZIP32", i);
            return ZB_STREAM_ERROR;
        }
    }
    if (n == 0) { *dst_len = 0; return 0; }
    std::vector<uint64_t> slot(n + 1), clen(n);
    std::vector<uint32_t> crc(n);
    std::vector<int32_t> status(n);
    slot[0] = 0;
    for (size_t i = 0; i < n; i++) slot[i + 1] = slot[i] + compressBound((unsigned long)(src_off[i + 1] - src_off[i])) + 16;

    Ctx* c = ctx_acquire_own();
    if (!c) return ZB_MEM_ERROR;
    cudaStream_t s = c->own_stream;
    do {
        if ((rc = c->out.ensure(slot[n] + 64)) != 0) break;
        uint8_t* d_arena = c->out.as<uint8_t>();
        cudaStreamSynchronize(s);                               // the arena may still be read by the context's previous borrower
        if ((rc = zb200_deflate_batch(src, src_off, n, d_arena, slot.data(), clen.data(), crc.data(), nullptr, status.data(), level,
                                      ZB200_WRAP_RAW, nullptr)) != 0) break;
        // ---- layout by prefix sum; local headers (zip.c:969-1032) into one blob ----
        uint64_t pos = 0, blob_len = 0;
        size_t npieces = 0;
        for (size_t i = 0; i < n; i++) {
            if (status[i] != 0) { rc = status[i]; set_error("zb200_zip_segment: member %zu failed (%d)", i, rc); break; }
            const size_t nl = strlen(names[i]);
            members[i].local_off = pos; members[i].comp_len = clen[i]; members[i].raw_len = src_off[i + 1] - src_off[i];
            members[i].crc32 = crc[i]; members[i].reserved = 0;
            pos += 30 + nl + clen[i];
            blob_len += 30 + nl;
            npieces += 1 + (size_t)((clen[i] + kPieceBytes - 1) / kPieceBytes);
        }
        if (rc) break;
        if (pos >= 0xffffffffull) { set_error("zb200_zip_segment: segment of %llu bytes is outside ZIP32", (unsigned long long)pos); rc = ZB_STREAM_ERROR; break; }
        if (pos > *dst_len) { *dst_len = (size_t)pos; set_error("zb200_zip_segment: output buffer too small"); rc = ZB_BUF_ERROR; break; }
        const bool dst_on_host = classify(dst) != kDevice;
        uint8_t* d_arc = (uint8_t*)dst;
        if (dst_on_host) {
            if ((rc = c->in.ensure(pos + 64)) != 0) break;
            d_arc = c->in.as<uint8_t>();
        }
        const size_t tab_at = (size_t)((blob_len + 15) & ~(uint64_t)15), up_bytes = tab_at + npieces * sizeof(Piece);
        if ((rc = c->ensure_pinned(up_bytes)) != 0) break;
        if ((rc = c->ws[0].ensure(up_bytes + 64)) != 0) break;
        uint8_t* h_up = (uint8_t*)c->pinned;
        uint8_t* d_up = c->ws[0].as<uint8_t>();
        Piece* pieces = reinterpret_cast<Piece*>(h_up + tab_at);
        uint64_t bpos = 0;
        size_t k = 0;
        for (size_t i = 0; i < n; i++) {
            const size_t nl = strlen(names[i]);
            uint8_t* lh = h_up + bpos;
            put32(lh, 0x04034b50ul); put16(lh + 4, 20); put16(lh + 6, 0); put16(lh + 8, 8 /* Z_DEFLATED */);
            put32(lh + 10, dos_datetime); put32(lh + 14, crc[i]); put32(lh + 18, clen[i]); put32(lh + 22, members[i].raw_len);
            put16(lh + 26, (uint32_t)nl); put16(lh + 28, 0);
            memcpy(lh + 30, names[i], nl);
            pieces[k++] = Piece{d_up + bpos, d_arc + members[i].local_off, 30 + nl};
            for (uint64_t o = 0; o < clen[i]; o += kPieceBytes)
                pieces[k++] = Piece{d_arena + slot[i] + o, d_arc + members[i].local_off + 30 + nl + o, std::min<uint64_t>(kPieceBytes, clen[i] - o)};
            bpos += 30 + nl;
        }
        cudaError_t e = cudaMemcpyAsync(d_up, h_up, up_bytes, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { set_error("zip table upload failed: %s", cudaGetErrorString(e)); rc = ZB_STREAM_ERROR; break; }
        ZB_LAUNCH(k_zip_gather, (unsigned)k, kGatherThreads, 0, s, reinterpret_cast<const Piece*>(d_up + tab_at));
        e = cudaGetLastError();
        const bool drain_threads = dst_on_host && pos >= HostStager::kMinBytes && classify(dst) == kHostPageable;
        if (e == cudaSuccess && dst_on_host && !drain_threads) e = cudaMemcpyAsync(dst, d_arc, pos, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("zip assembly failed: %s", cudaGetErrorString(e)); rc = ZB_STREAM_ERROR; break; }
        if (drain_threads) {                                    // malloc'ed archive buffer: host threads pull it through pinned slots
            HostDrainer drainer;
            if ((rc = drainer.drain(dst, d_arc, pos)) != 0) break;
        }
        *dst_len = (size_t)pos;
    } while (0);
    ctx_release(c, s);
    return rc;
}

// Central directory (zip.c:940-967) and end record (zip.c:1203-1240) for members whose local headers sit at
// members[i].local_off of the finished archive; the directory itself starts at cd_offset.  Host memory, no GPU work.
ZB_API int zb200_zip_directory(const char* const* names, const zb200_zip_member* members, size_t n, uint32_t dos_datetime,
                               uint64_t cd_offset, void* dst, size_t* dst_len)
{
    if (names == nullptr || members == nullptr || dst == nullptr || dst_len == nullptr || n > 65535u) return ZB_STREAM_ERROR;
    uint64_t need = 22;
    for (size_t i = 0; i < n; i++) need += 46 + strlen(names[i]);
    if (cd_offset + need >= 0xffffffffull) { set_error("zb200_zip_directory: archive is outside ZIP32"); return ZB_STREAM_ERROR; }
    if (need > *dst_len) { *dst_len = (size_t)need; return ZB_BUF_ERROR; }
    uint8_t* ch = (uint8_t*)dst;
    for (size_t i = 0; i < n; i++) {
        const size_t nl = strlen(names[i]);
        put32(ch, 0x02014b50ul); put16(ch + 4, 0); put16(ch + 6, 20); put16(ch + 8, 0); put16(ch + 10, 8);
        put32(ch + 12, dos_datetime); put32(ch + 16, members[i].crc32); put32(ch + 20, members[i].comp_len); put32(ch + 24, members[i].raw_len);
        put16(ch + 28, (uint32_t)nl); put16(ch + 30, 0); put16(ch + 32, 0); put16(ch + 34, 0); put16(ch + 36, 0);
        put32(ch + 38, 0); put32(ch + 42, members[i].local_off);
        memcpy(ch + 46, names[i], nl);
        ch += 46 + nl;
    }
    const uint64_t cd_size = need - 22;
    put32(ch, 0x06054b50ul); put16(ch + 4, 0); put16(ch + 6, 0); put16(ch + 8, (uint32_t)n); put16(ch + 10, (uint32_t)n);
    put32(ch + 12, cd_size); put32(ch + 16, cd_offset); put16(ch + 20, 0);
    *dst_len = (size_t)need;
    return 0;
}

ZB_API int zb200_zip_build(const char* const* names, const void* src, const uint64_t* src_off, size_t n, int level,
                           uint32_t dos_datetime, void* dst, size_t* dst_len)
{
    if (dst == nullptr || dst_len == nullptr || n > 65535u) return ZB_STREAM_ERROR;
    std::vector<zb200_zip_member> members(n ? n : 1);
    size_t cd = 22;
    for (size_t i = 0; i < n; i++) cd += names && names[i] ? 46 + strlen(names[i]) : 0;
    if (*dst_len < cd) { *dst_len = cd; return ZB_BUF_ERROR; }
    size_t seg = *dst_len - cd;                                 // what is left for the members once the directory has its room
    int rc = zb200_zip_segment(names, src, src_off, n, level, dos_datetime, dst, &seg, members.data());
    if (rc == ZB_BUF_ERROR) *dst_len = seg + cd;
    if (rc) return rc;
    if (classify(dst) == kDevice) { set_error("zb200_zip_build: the archive buffer must be host memory"); return ZB_STREAM_ERROR; }
    size_t dl = cd;
    if ((rc = zb200_zip_directory(names, members.data(), n, dos_datetime, seg, (uint8_t*)dst + seg, &dl)) != 0) return rc;
    *dst_len = seg + dl;
    return 0;
}
