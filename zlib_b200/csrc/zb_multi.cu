// zb_multi.cu -- the sharded paths of SURVEY.md 8(e) behind the C ABI: one process, one host thread per GPU.
//
// zlib.h's calls see one host buffer; a box has eight B200s.  These entry points take the same host buffers and spread
// the work over `ndev` devices with no Python or torch in the way:
//
//   zb200_multi_deflate        the input is cut into chunk-aligned pieces dealt round robin (global piece g = round * ndev
//                              + device); every device compresses its piece with the 32 KiB in front of it as dictionary
//                              (zb200_deflate_shard) and ends on a byte boundary, so pieces concatenate (deflate.c:808-819).
//                              After every round ONE NCCL all-gather over NVLink of {compressed bytes, input bytes, crc32,
//                              adler32} per device gives every piece its place in the stream; its bytes then go D2H
//                              straight to that place while the next round is compressed.  Checksums are folded in stream
//                              order with crc32_combine / adler32_combine (crc32.c:370, adler32.c:128) -- the combine is not
//                              commutative, so no reduction collective can do it.
//   zb200_multi_checksum       a contiguous slice per device, one all-gather of {crc32, adler32, len}, the same fold.
//   zb200_multi_inflate_batch  streams dealt in contiguous ranges of nearly equal compressed size; nothing is exchanged
//                              (status and length of every stream land in the caller's arrays).
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): a process that never calls these does not need it, and a process
// that already carries an NCCL (torch) keeps its own.  Threads and communicators are created once and reused.
#include "zb_common.cuh"
#include "zb200_internal.h"

#include <dlfcn.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

extern "C" unsigned long crc32_combine(unsigned long crc1, unsigned long crc2, long len2);   // zapi_checksum.c

namespace zb {

// ---- the few NCCL entry points, resolved at run time ----
typedef void* nccl_comm_t;
struct NcclApi {
    int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
constexpr int kNcclUint64 = 5;                                  // ncclDataType_t (nccl.h)

static NcclApi& nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.CommInitAll = (int (*)(nccl_comm_t*, int, const int*))dlsym(h, "ncclCommInitAll");
        api.AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllGather");
        api.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
        api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        api.ok = api.CommInitAll && api.AllGather && api.CommDestroy && api.GetErrorString;
    });
    return api;
}

// ---- one worker thread per device, alive for the life of the process ----
struct Team {
    int ndev = 0;
    std::vector<nccl_comm_t> comms;
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<int(int)> job;                                // the work of one call, run by every worker with its rank
    uint64_t epoch = 0;
    int pending = 0, rc = 0;
    bool quit = false;
    // per-device scratch
    std::vector<cudaStream_t> copy_stream;
    std::vector<uint64_t*> d_send, d_recv;                      // all-gather buffers (device)
    std::vector<uint64_t*> h_recv;                              // pinned copies of the gathered records
    std::vector<uint8_t*> d_out[2];                             // two output buffers per device (rounds alternate)
    std::vector<size_t> d_out_cap[2];
    std::vector<cudaEvent_t> ev[2];                             // D2H of the buffer finished
};

static std::mutex g_team_mu;
static Team* g_team = nullptr;
static std::mutex g_call_mu;                                    // one multi-GPU call at a time

static void team_worker(Team* t, int r)
{
    zb200_init(r);                                              // binds this thread to device r
    uint64_t seen = 0;
    for (;;) {
        std::function<int(int)> job;
        {
            std::unique_lock<std::mutex> lk(t->mu);
            t->cv_go.wait(lk, [&] { return t->quit || t->epoch != seen; });
            if (t->quit) return;
            seen = t->epoch;
            job = t->job;
        }
        const int rc = job(r);
        {
            std::lock_guard<std::mutex> lk(t->mu);
            if (rc && !t->rc) t->rc = rc;
            if (--t->pending == 0) t->cv_done.notify_all();
        }
    }
}

static int team_run(Team* t, std::function<int(int)> job)
{
    std::unique_lock<std::mutex> lk(t->mu);
    t->job = std::move(job);
    t->pending = t->ndev; t->rc = 0;
    t->epoch++;
    t->cv_go.notify_all();
    t->cv_done.wait(lk, [&] { return t->pending == 0; });
    return t->rc;
}

static Team* team_get(int ndev)
{
    std::lock_guard<std::mutex> lk(g_team_mu);
    if (g_team && g_team->ndev == ndev) return g_team;
    if (g_team) {                                               // a different device count: rebuild
        { std::lock_guard<std::mutex> l2(g_team->mu); g_team->quit = true; }
        g_team->cv_go.notify_all();
        for (auto& th : g_team->threads) th.join();
        NcclApi& api = nccl_api();
        for (auto c : g_team->comms) if (c) api.CommDestroy(c);
        delete g_team;                                          // device scratch is left to the driver at exit
        g_team = nullptr;
    }
    NcclApi& api = nccl_api();
    if (!api.ok) { set_error("zb200_multi: libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing"); return nullptr; }
    Team* t = new Team();
    t->ndev = ndev;
    t->comms.assign(ndev, nullptr);
    std::vector<int> devs(ndev);
    for (int i = 0; i < ndev; i++) devs[i] = i;
    const int nrc = api.CommInitAll(t->comms.data(), ndev, devs.data());
    if (nrc != 0) { set_error("ncclCommInitAll failed: %s", api.GetErrorString(nrc)); delete t; return nullptr; }
    t->copy_stream.assign(ndev, nullptr);
    t->d_send.assign(ndev, nullptr); t->d_recv.assign(ndev, nullptr); t->h_recv.assign(ndev, nullptr);
    for (int k = 0; k < 2; k++) { t->d_out[k].assign(ndev, nullptr); t->d_out_cap[k].assign(ndev, 0); t->ev[k].assign(ndev, nullptr); }
    for (int r = 0; r < ndev; r++) t->threads.emplace_back(team_worker, t, r);
    // per-device scratch, allocated by the worker that owns the device
    const int rc = team_run(t, [t, ndev](int r) -> int {
        bool ok = cudaStreamCreateWithFlags(&t->copy_stream[r], cudaStreamNonBlocking) == cudaSuccess &&
                  cudaMalloc(&t->d_send[r], 64) == cudaSuccess && cudaMalloc(&t->d_recv[r], (size_t)ndev * 64) == cudaSuccess &&
                  cudaMallocHost(&t->h_recv[r], (size_t)ndev * 64) == cudaSuccess &&
                  cudaEventCreateWithFlags(&t->ev[0][r], cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&t->ev[1][r], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { cudaGetLastError(); set_error("zb200_multi: scratch allocation failed on device %d", r); return ZB_MEM_ERROR; }
        return 0;
    });
    if (rc) return nullptr;                                     // (the team is leaked on this path; the process is in trouble anyway)
    g_team = t;
    return t;
}

// One all-gather of four 64-bit words per device; the gathered table arrives in h_recv[r] (ndev x 4 words).
static int gather4(Team* t, int r, const uint64_t rec[4], cudaStream_t s)
{
    NcclApi& api = nccl_api();
    if (cudaMemcpyAsync(t->d_send[r], rec, 32, cudaMemcpyHostToDevice, s) != cudaSuccess) { set_error("zb200_multi: record upload failed"); return ZB_STREAM_ERROR; }
    const int nrc = api.AllGather(t->d_send[r], t->d_recv[r], 4, kNcclUint64, t->comms[r], s);
    if (nrc != 0) { set_error("ncclAllGather failed: %s", api.GetErrorString(nrc)); return ZB_STREAM_ERROR; }
    if (cudaMemcpyAsync(t->h_recv[r], t->d_recv[r], (size_t)t->ndev * 32, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { set_error("zb200_multi: record readback failed: %s", cudaGetErrorString(cudaGetLastError())); return ZB_STREAM_ERROR; }
    return 0;
}

static uint32_t adler_join(uint32_t a1, uint32_t a2, uint64_t len2)   // exact Adler-32 of A||B
{
    const uint64_t rem = len2 % kAdlerBase;
    const uint64_t s1 = a1 & 0xffffu, s2 = a1 >> 16, t1 = a2 & 0xffffu, t2 = a2 >> 16;
    const uint64_t r1 = (s1 + t1 + kAdlerBase - 1) % kAdlerBase;
    const uint64_t r2 = (s2 + t2 + rem * s1 + kAdlerBase - rem) % kAdlerBase;
    return (uint32_t)(r1 | (r2 << 16));
}

static int clamp_devices(int ndev)
{
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess) { cudaGetLastError(); have = 0; }
    if (have <= 0) { set_error("zb200: no CUDA device available -- this library has no CPU path"); return -1; }
    if (ndev <= 0 || ndev > have) ndev = have;
    return ndev;
}

}  // namespace zb

using namespace zb;

ZB_API int zb200_multi_deflate(const void* src, size_t src_len, void* dst, size_t* dst_len, int level, int wrap, int ndev,
                               uint32_t* crc_out, uint32_t* adler_out)
{
    if (!dst || !dst_len || (src_len && !src) || level < -1 || level > 9 || wrap < 0 || wrap > 2) { set_error("zb200_multi_deflate: bad argument"); return ZB_STREAM_ERROR; }
    if ((ndev = clamp_devices(ndev)) < 0) return ZB_STREAM_ERROR;
    std::lock_guard<std::mutex> call(g_call_mu);
    Team* t = team_get(ndev);
    if (!t) return ZB_STREAM_ERROR;
    const int lvl = level < 0 ? 6 : level;
    // ---- pieces: chunk aligned, at most 256 MiB, a whole number of rounds ----
    const uint64_t chunk = ZB200_CHUNK, nchunks = (src_len + chunk - 1) / chunk;
    uint64_t rounds = std::max<uint64_t>(1, (src_len + (uint64_t)ndev * (256u << 20) - 1) / ((uint64_t)ndev * (256u << 20)));
    if (rounds < 4 && nchunks >= (uint64_t)ndev * 4 * 64) rounds = 4;               // enough rounds to hide the D2H behind the kernels
    const uint64_t npieces = rounds * ndev;
    const uint64_t per = std::max<uint64_t>(1, (nchunks + npieces - 1) / npieces) * chunk;
    uint8_t header[10]; size_t hl = 0;
    if (wrap == ZB200_WRAP_ZLIB) {
        const uint32_t fl = lvl < 2 ? 0 : lvl < 6 ? 1 : lvl == 6 ? 2 : 3;           // deflate.c:628-636
        uint32_t h = (0x78u << 8) | (fl << 6);
        h += 31 - h % 31;
        header[0] = (uint8_t)(h >> 8); header[1] = (uint8_t)h; hl = 2;
    } else if (wrap == ZB200_WRAP_GZIP) {
        const uint8_t g[10] = {31, 139, 8, 0, 0, 0, 0, 0, (uint8_t)(lvl == 9 ? 2 : lvl < 2 ? 4 : 0), 3};   // deflate.c:590-593
        memcpy(header, g, 10); hl = 10;
    }
    const size_t cap = *dst_len;
    struct Shared { std::atomic<int> overflow{0}; uint64_t total = 0; uint32_t crc = 0, adler = 1; } sh;
    const uint8_t* in = (const uint8_t*)src;
    uint8_t* out = (uint8_t*)dst;
    const size_t piece_cap = (size_t)(per + (per >> 12) + (per >> 14) + 11 + 64);

    const int rc = team_run(t, [&, t](int r) -> int {
        cudaStream_t cs = t->copy_stream[r];
        uint64_t pos = hl;                                      // every worker folds the same records, so all agree on `pos`
        uint32_t crc = 0, adler = 1;
        void* job[2] = {nullptr, nullptr};
        // Piece j + 1 is enqueued (zb200_deflate_shard_begin) before piece j's length is read back, so the GPU never waits
        // for the host; the D2H of piece j runs on the copy stream behind the kernels of piece j + 1.
        auto range = [&](uint64_t j, uint64_t& a, uint64_t& b) {
            const uint64_t g = j * ndev + r;
            a = std::min<uint64_t>(src_len, g * per); b = std::min<uint64_t>(src_len, (g + 1) * per);
        };
        auto begin = [&](uint64_t j) -> int {
            const int k = (int)(j & 1);
            uint64_t a, b;
            range(j, a, b);
            const bool ends_stream = b == src_len && a < src_len;     // the piece that ends the stream (later ones, if any, are empty)
            const bool empty_stream_owner = src_len == 0 && j == 0 && r == 0;
            if (b <= a && !empty_stream_owner) return 0;
            if (t->d_out_cap[k][r] < piece_cap) {
                if (t->d_out[k][r]) cudaFree(t->d_out[k][r]);
                t->d_out[k][r] = nullptr; t->d_out_cap[k][r] = 0;
                if (cudaMalloc(&t->d_out[k][r], piece_cap) != cudaSuccess) { cudaGetLastError(); set_error("zb200_multi_deflate: device allocation of %zu bytes failed", piece_cap); return ZB_MEM_ERROR; }
                t->d_out_cap[k][r] = piece_cap;
            }
            if (j >= 2 && cudaEventSynchronize(t->ev[k][r]) != cudaSuccess) { set_error("zb200_multi_deflate: D2H failed"); return ZB_STREAM_ERROR; }   // the buffer's previous contents are out
            const size_t dl = (size_t)std::min<uint64_t>(a, 32768);
            const int flags = ZB200_DEFLATE_NO_HEADER | ZB200_DEFLATE_NO_TRAILER | ((ends_stream || empty_stream_owner) ? 0 : ZB200_DEFLATE_NOT_LAST);
            return zb200_deflate_shard_begin(&job[k], in + a, (size_t)(b - a), dl ? in + a - dl : nullptr, dl, t->d_out[k][r], piece_cap, level,
                                             ZB200_WRAP_RAW, flags, nullptr);
        };
        int rc0 = begin(0);
        for (uint64_t j = 0; j < rounds && !rc0; j++) {
            const int k = (int)(j & 1);
            if (j + 1 < rounds && (rc0 = begin(j + 1)) != 0) break;
            uint64_t a, b;
            range(j, a, b);
            uint64_t rec[4] = {0, b - a, 0, 1};
            if (job[k]) {
                size_t got = 0;
                uint32_t c32 = 0, a32 = 1;
                rc0 = zb200_deflate_shard_end(job[k], &got, &c32, &a32);
                job[k] = nullptr;
                if (rc0) break;
                rec[0] = got; rec[2] = c32; rec[3] = a32;
            }
            if ((rc0 = gather4(t, r, rec, cs)) != 0) break;
            const uint64_t* all = t->h_recv[r];
            uint64_t mine_at = 0;
            for (int q = 0; q < ndev; q++) {
                const uint64_t cl = all[4 * q], m = all[4 * q + 1];
                if (q == r) mine_at = pos;
                pos += cl;
                if (m) { crc = (uint32_t)crc32_combine(crc, (uint32_t)all[4 * q + 2], (long)m); adler = adler_join(adler, (uint32_t)all[4 * q + 3], m); }
            }
            if (rec[0]) {
                if (mine_at + rec[0] > cap) sh.overflow.store(1);
                else if (cudaMemcpyAsync(out + mine_at, t->d_out[k][r], rec[0], cudaMemcpyDeviceToHost, cs) != cudaSuccess) { set_error("zb200_multi_deflate: D2H failed: %s", cudaGetErrorString(cudaGetLastError())); rc0 = ZB_STREAM_ERROR; break; }
            }
            cudaEventRecord(t->ev[k][r], cs);
        }
        for (int k = 0; k < 2; k++)                              // after a failure: nothing may stay enqueued against the scratch
            if (job[k]) { size_t g = 0; zb200_deflate_shard_end(job[k], &g, nullptr, nullptr); job[k] = nullptr; }
        if (rc0) { cudaStreamSynchronize(cs); return rc0; }
        if (cudaStreamSynchronize(cs) != cudaSuccess) { set_error("zb200_multi_deflate: D2H failed: %s", cudaGetErrorString(cudaGetLastError())); return ZB_STREAM_ERROR; }
        if (r == 0) { sh.total = pos; sh.crc = crc; sh.adler = adler; }
        return 0;
    });
    if (rc) return rc;
    const size_t tl = wrap == ZB200_WRAP_ZLIB ? 4 : wrap == ZB200_WRAP_GZIP ? 8 : 0;
    *dst_len = (size_t)sh.total + tl;
    if (crc_out) *crc_out = sh.crc;
    if (adler_out) *adler_out = sh.adler;
    if (sh.overflow.load() || sh.total + tl > cap) { set_error("output buffer too small: need %llu, have %zu", (unsigned long long)(sh.total + tl), cap); return ZB_BUF_ERROR; }
    memcpy(out, header, hl);
    uint8_t* tr = out + sh.total;
    if (wrap == ZB200_WRAP_ZLIB) { tr[0] = sh.adler >> 24; tr[1] = sh.adler >> 16; tr[2] = sh.adler >> 8; tr[3] = (uint8_t)sh.adler; }
    else if (wrap == ZB200_WRAP_GZIP) {
        for (int i = 0; i < 4; i++) tr[i] = (uint8_t)(sh.crc >> (8 * i));
        for (int i = 0; i < 4; i++) tr[4 + i] = (uint8_t)((uint64_t)src_len >> (8 * i));
    }
    return 0;
}

ZB_API int zb200_multi_checksum(const void* buf, size_t len, int ndev, uint32_t* crc_out, uint32_t* adler_out)
{
    if (len && !buf) { set_error("zb200_multi_checksum: bad argument"); return ZB_STREAM_ERROR; }
    if ((ndev = clamp_devices(ndev)) < 0) return ZB_STREAM_ERROR;
    std::lock_guard<std::mutex> call(g_call_mu);
    Team* t = team_get(ndev);
    if (!t) return ZB_STREAM_ERROR;
    const uint64_t per = ((len + ndev - 1) / ndev + 4095) & ~(uint64_t)4095;
    struct { uint32_t crc = 0, adler = 1; } sh;
    const int rc = team_run(t, [&, t](int r) -> int {
        const uint64_t a = std::min<uint64_t>(len, (uint64_t)r * per), b = std::min<uint64_t>(len, (uint64_t)(r + 1) * per);
        uint32_t c32 = 0, a32 = 1;
        if (b > a) { const int rc1 = zb200_checksum((const uint8_t*)buf + a, (size_t)(b - a), &c32, &a32, nullptr); if (rc1) return rc1; }
        const uint64_t rec[4] = {c32, a32, b - a, 0};
        const int rc2 = gather4(t, r, rec, t->copy_stream[r]);
        if (rc2) return rc2;
        if (r == 0) {
            uint32_t crc = 0, adler = 1;
            for (int q = 0; q < ndev; q++) {
                const uint64_t* e = t->h_recv[r] + 4 * q;
                if (e[2]) { crc = (uint32_t)crc32_combine(crc, (uint32_t)e[0], (long)e[2]); adler = adler_join(adler, (uint32_t)e[1], e[2]); }
            }
            sh.crc = crc; sh.adler = adler;
        }
        return 0;
    });
    if (rc) return rc;
    if (crc_out) *crc_out = sh.crc;
    if (adler_out) *adler_out = sh.adler;
    return 0;
}

ZB_API int zb200_multi_inflate_batch(const void* src, const uint64_t* src_off, size_t n, void* dst, const uint64_t* dst_off,
                                     uint64_t* dst_len, int32_t* status, int wrap, int ndev)
{
    if (n && (!src || !src_off || !dst || !dst_off || !dst_len || !status)) { set_error("zb200_multi_inflate_batch: bad argument"); return ZB_STREAM_ERROR; }
    if ((ndev = clamp_devices(ndev)) < 0) return ZB_STREAM_ERROR;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> call(g_call_mu);
    Team* t = team_get(ndev);
    if (!t) return ZB_STREAM_ERROR;
    // contiguous ranges with nearly equal compressed bytes (streams stay in order, so the output arena needs no shuffle)
    std::vector<size_t> cut(ndev + 1, n);
    cut[0] = 0;
    const uint64_t total = src_off[n] - src_off[0];
    size_t i = 0;
    for (int r = 0; r < ndev; r++) {
        const uint64_t want = src_off[0] + total * (uint64_t)(r + 1) / ndev;
        while (i < n && (r == ndev - 1 || (src_off[i] + src_off[i + 1]) / 2 <= want)) i++;
        cut[r + 1] = i;
    }
    return team_run(t, [&](int r) -> int {
        const size_t a = cut[r], b = cut[r + 1];
        if (b <= a) return 0;
        // the sub-batch sees its own part of the arenas, with offset tables that start at zero
        std::vector<uint64_t> so(b - a + 1), dof(b - a + 1);
        for (size_t k = 0; k <= b - a; k++) { so[k] = src_off[a + k] - src_off[a]; dof[k] = dst_off[a + k] - dst_off[a]; }
        return zb200_inflate_batch((const uint8_t*)src + src_off[a], so.data(), b - a, (uint8_t*)dst + dst_off[a], dof.data(), dst_len + a,
                                   status + a, wrap, nullptr);
    });
}

ZB_API int zb200_multi_devices(void) { const int n = clamp_devices(0); return n < 0 ? 0 : n; }
