// zb_common.cuh -- internal declarations shared by the zb200 CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <atomic>

#include "../../include/zb200.h"

#define ZB_API extern "C" __attribute__((visibility("default")))
#define ZB_STR1(x) #x
#define ZB_STR2(x) ZB_STR1(x)

// zlib return codes (include/zlib.h)
enum { ZB_OK = 0, ZB_STREAM_END = 1, ZB_NEED_DICT = 2, ZB_STREAM_ERROR = -2, ZB_DATA_ERROR = -3,
       ZB_MEM_ERROR = -4, ZB_BUF_ERROR = -5 };

namespace zb {

int device_sms();                          // SM count of the calling thread's device (148 on B200: 2 dies x 74)
constexpr uint32_t kCrcPoly = 0xEDB88320u;
constexpr uint32_t kAdlerBase = 65521u;

extern std::atomic<uint64_t> g_launches;
void set_error(const char* fmt, ...);

// Growable device buffer owned by a context.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);              // 0 or ZB_MEM_ERROR
    void release();
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// One execution context: a stream, scratch memory and pinned staging.  Contexts are
// pooled; a z_stream or an extension call borrows one for the duration of a call.
struct Ctx {
    cudaStream_t own_stream = nullptr;
    cudaEvent_t  idle = nullptr;           // recorded when the context is released
    DevBuf in, out, ws[12];
    DevBuf small;                          // few-KB result words
    DevBuf flags;                          // 256 bytes, zero whenever no kernel of this context is running (self-resetting counters)
    int ensure_flags();
    void* pinned = nullptr; size_t pinned_cap = 0;
    int ensure_pinned(size_t bytes);
    cudaStream_t aux[2] = {nullptr, nullptr};   // copy-in / copy-out streams of the slab pipeline
    cudaEvent_t* evs = nullptr; int nev = 0;
    int ensure_aux(int nevents);               // both streams and at least `nevents` events
    int device = 0;                        // the device this context's streams and buffers live on
    Ctx* next = nullptr;
};

int  ensure_init();                        // 0 or negative zlib code
Ctx* ctx_acquire(cudaStream_t use);        // waits (on `use`) for the context's previous work
Ctx* ctx_acquire_own();                    // for work on the context's own stream
void ctx_release(Ctx* c, cudaStream_t used);
// NULL means CUDA's legacy default stream (stream 0), which orders with torch's default stream.
inline cudaStream_t pick_stream(Ctx*, void* user) { return (cudaStream_t)user; }

enum MemKind { kHostPageable = 0, kHostPinned = 1, kDevice = 2 };
MemKind classify(const void* p);

// Brings `len` bytes at `src` (any memory kind) to the device; returns a device pointer that
// is either `src` itself or c->in.  Asynchronous on `s` for pinned/device sources.
const uint8_t* to_device(Ctx* c, const void* src, size_t len, cudaStream_t s, int* err);

// Pageable host memory -> device faster than the driver's own staging copy (one host thread, ~10 GB/s): a few host
// threads copy pieces into pinned slots and send each on its own stream, so the caller's malloc'ed buffer moves at several
// times that.  The consumer stream waits per byte range; the host blocks only until the pieces of that range are issued.
struct HostStager {
    static constexpr size_t kPiece = 4u << 20;
    static constexpr int kThreads = 8;
    static constexpr size_t kMinBytes = 64u << 20;             // below this the driver's path is as good
    int start(const void* src, void* d_dst, size_t n, cudaStream_t after);
    int wait_range(size_t off, size_t len, cudaStream_t consumer);
    void finish();
    ~HostStager() { finish(); }
    struct Impl;
    Impl* impl = nullptr;
};

// The other direction: device -> pageable host memory.  drain() returns when the bytes are in the caller's buffer; inside,
// each of a few host threads pulls its pieces into a pinned slot and copies them on.  The source must be complete.
struct HostDrainer {
    int drain(void* dst, const void* d_src, size_t n);
    void finish();
    ~HostDrainer() { finish(); }
    uint8_t* slots = nullptr;
    cudaStream_t st[HostStager::kThreads] = {};
};

#define ZB_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            zb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return ZB_STREAM_ERROR;                                                     \
        }                                                                               \
    } while (0)

// Per-kernel timing for bench.py's roofline line: when enabled, every launch is bracketed by CUDA
// events on the launching stream (zb200_profile / zb200_profile_report in zb_runtime.cu).
extern bool g_profile;
void profile_mark(const char* name, cudaStream_t s, bool begin);

#define ZB_LAUNCH(kernel, grid, block, smem, stream, ...)                               \
    do {                                                                                \
        if (zb::g_profile) zb::profile_mark(#kernel, (stream), true);                   \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                     \
        if (zb::g_profile) zb::profile_mark(#kernel, (stream), false);                  \
        zb::g_launches.fetch_add(1, std::memory_order_relaxed);                         \
    } while (0)

#define ZB_CHECK_LAUNCH()                                                               \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            zb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return ZB_STREAM_ERROR;                                                     \
        }                                                                               \
    } while (0)

// ---- checksum engine (zb_checksum.cu) ----
int checksum_setup();                      // uploads tables; called from ensure_init
// crc32(0,..)/adler32(1,..) of d_buf[0..len) -> d_out2[0..1]; async on s.
int checksum_launch(Ctx* c, const uint8_t* d_buf, size_t len, uint32_t* d_out2, cudaStream_t s);
// n buffers in one launch: buffer i = d_base[d_off[i] .. +len) with len = d_lens[i] or d_off[i+1]-d_off[i].
// Either writes crc/adler per buffer, or (d_expect != nullptr) checks {crc32, length} pairs and flags d_ok[i].
int checksum_batch_launch(const uint8_t* d_base, const uint64_t* d_off, const uint64_t* d_lens, size_t n, uint32_t* d_crc,
                          uint32_t* d_adler, const uint32_t* d_expect, int32_t* d_ok, cudaStream_t s);
// job mode of the deflate batch: per-chunk checksums joined per job (tables of zb_deflate.cuh)
struct ChunkDesc; struct JobDesc;
int checksum_jobs_launch(Ctx* c, const uint8_t* d_base, const ChunkDesc* d_cd, uint32_t nchunks, const JobDesc* d_jobs,
                         uint32_t njobs, uint32_t* d_crc, uint32_t* d_adler, cudaStream_t s);
const unsigned long* host_crc_table();

}  // namespace zb
