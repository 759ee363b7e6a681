// zb_runtime.cu -- process-wide state of the zb200 engine: device selection, a pool of
// execution contexts (stream + scratch + pinned staging), memory classification and
// the small C-ABI helpers of include/zb200.h.  No codec arithmetic lives here.
#include "zb_common.cuh"

#include <mutex>
#include <thread>
#include <atomic>
#include <memory>
#include <vector>
#include <string>
#include <map>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace zb {

std::atomic<uint64_t> g_launches{0};
bool g_profile = false;

struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
void profile_mark(const char* name, cudaStream_t s, bool begin)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (begin) {
        ProfRec r{name, nullptr, nullptr};
        cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, s);
        g_prof.push_back(r);
    } else if (!g_prof.empty()) {
        cudaEventRecord(g_prof.back().e1, s);
    }
}

static thread_local char t_err[512];
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    if (getenv("ZB200_TRACE")) fprintf(stderr, "[zb200] %s\n", t_err);
}

// Per-device state.  A thread's calls run on its *bound* device (zb200_init(d) binds the calling thread; threads that
// never called it use the process default = the first device bound, else CUDA's current device).  Every device has its
// own context pool, table uploads and kernel attributes, so one process can drive all eight GPUs of a box (one host
// thread per device, see zb_multi.cu) as well as the one-process-per-GPU layout.
constexpr int kMaxDevices = 32;
struct DevState {
    std::mutex mu;
    std::atomic<int> state{0};             // 0 = untouched, 1 = ready, -1 = failed
    Ctx* free_list = nullptr;
    int sms = 148;
};
static DevState g_dev[kMaxDevices];
static std::atomic<int> g_default_device{-1};
static thread_local int t_device = -1;

int deflate_setup();                       // zb_deflate.cu: kernel attributes (per device)
int checksum_attr_setup();                 // zb_checksum.cu

int DevBuf::ensure(size_t bytes)
{
    if (bytes <= cap) return 0;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + (bytes >> 3) + 256;      // slack so slowly growing inputs do not realloc every call
    if (cudaMalloc(&p, want) != cudaSuccess) {
        cudaGetLastError();
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            set_error("device allocation of %zu bytes failed", bytes);
            return ZB_MEM_ERROR;
        }
        want = bytes;
    }
    cap = want;
    return 0;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

int Ctx::ensure_pinned(size_t bytes)
{
    if (bytes <= pinned_cap) return 0;
    if (pinned) cudaFreeHost(pinned);
    pinned = nullptr; pinned_cap = 0;
    if (cudaMallocHost(&pinned, bytes) != cudaSuccess) {
        cudaGetLastError();
        set_error("pinned allocation of %zu bytes failed", bytes);
        return ZB_MEM_ERROR;
    }
    pinned_cap = bytes;
    return 0;
}

int Ctx::ensure_flags()
{
    if (flags.p) return 0;
    int rc = flags.ensure(256);
    if (rc) return rc;
    if (cudaMemset(flags.p, 0, 256) != cudaSuccess) { cudaGetLastError(); set_error("flag words could not be cleared"); return ZB_MEM_ERROR; }
    return 0;
}

int Ctx::ensure_aux(int nevents)
{
    for (int i = 0; i < 2; i++) {
        if (!aux[i] && cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError(); aux[i] = nullptr;
            set_error("copy stream creation failed");
            return ZB_MEM_ERROR;
        }
    }
    if (nevents > nev) {
        const int want = nevents + 16;
        cudaEvent_t* ne = (cudaEvent_t*)calloc((size_t)want, sizeof(cudaEvent_t));
        if (!ne) { set_error("out of host memory"); return ZB_MEM_ERROR; }
        for (int i = 0; i < nev; i++) ne[i] = evs[i];
        for (int i = nev; i < want; i++) {
            if (cudaEventCreateWithFlags(&ne[i], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                for (int k = nev; k < i; k++) cudaEventDestroy(ne[k]);
                free(ne);
                set_error("event creation failed");
                return ZB_MEM_ERROR;
            }
        }
        free(evs);
        evs = ne; nev = want;
    }
    return 0;
}

static int current_device_index()
{
    int dev = t_device;
    if (dev < 0) dev = g_default_device.load(std::memory_order_acquire);
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return dev;
}

int device_sms() { const int d = current_device_index(); return (d >= 0 && d < kMaxDevices) ? g_dev[d].sms : 148; }

int ensure_init()
{
    const int dev = current_device_index();
    if (dev < 0 || dev >= kMaxDevices) { set_error("zb200: device index %d out of range", dev); return ZB_STREAM_ERROR; }
    DevState& D = g_dev[dev];
    if (D.state.load(std::memory_order_acquire) == 1) {
        cudaSetDevice(dev);                                     // the calling thread may be new, or last used another device
        return 0;
    }
    std::lock_guard<std::mutex> lk(D.mu);
    if (D.state.load(std::memory_order_relaxed) == 1) { cudaSetDevice(dev); return 0; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("zb200: no CUDA device available -- this library has no CPU path");
        D.state.store(-1, std::memory_order_release);
        return ZB_STREAM_ERROR;
    }
    if (dev >= n) { set_error("zb200: device %d does not exist (%d devices)", dev, n); D.state.store(-1, std::memory_order_release); return ZB_STREAM_ERROR; }
    if (cudaSetDevice(dev) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", dev); D.state.store(-1, std::memory_order_release); return ZB_STREAM_ERROR; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { set_error("cudaGetDeviceProperties failed"); D.state.store(-1, std::memory_order_release); return ZB_STREAM_ERROR; }
    if (prop.major != 10) {
        set_error("zb200: device %d is sm_%d%d; this build contains sm_100a code only", dev, prop.major, prop.minor);
        D.state.store(-1, std::memory_order_release);
        return ZB_STREAM_ERROR;
    }
    D.sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    if (checksum_setup() != 0 || checksum_attr_setup() != 0 || deflate_setup() != 0) { D.state.store(-1, std::memory_order_release); return ZB_STREAM_ERROR; }
    int expected = -1;
    g_default_device.compare_exchange_strong(expected, dev, std::memory_order_acq_rel);
    D.state.store(1, std::memory_order_release);
    return 0;
}

Ctx* ctx_acquire(cudaStream_t use)
{
    const int dev = current_device_index();
    DevState& D = g_dev[dev];
    Ctx* c = nullptr;
    {
        std::lock_guard<std::mutex> lk(D.mu);
        if (D.free_list) { c = D.free_list; D.free_list = c->next; c->next = nullptr; }
    }
    if (!c) {
        c = new Ctx();
        c->device = dev;
        if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->idle, cudaEventDisableTiming) != cudaSuccess) {
            set_error("stream/event creation failed");
            delete c;
            return nullptr;
        }
        cudaEventRecord(c->idle, c->own_stream);
    }
    cudaStreamWaitEvent(use, c->idle, 0);    // scratch may still be in use by the previous borrower's stream
    return c;
}

Ctx* ctx_acquire_own()
{
    Ctx* c = ctx_acquire(nullptr);
    if (c) cudaStreamWaitEvent(c->own_stream, c->idle, 0);
    return c;
}

void ctx_release(Ctx* c, cudaStream_t used)
{
    cudaEventRecord(c->idle, used);
    DevState& D = g_dev[c->device];
    std::lock_guard<std::mutex> lk(D.mu);
    c->next = D.free_list;
    D.free_list = c;
}

MemKind classify(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return kHostPageable; }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return kDevice;
    if (a.type == cudaMemoryTypeHost) return kHostPinned;
    return kHostPageable;
}

const uint8_t* to_device(Ctx* c, const void* src, size_t len, cudaStream_t s, int* err)
{
    *err = 0;
    if (len == 0) {                         // kernels never dereference an empty buffer, but want a valid pointer
        if ((*err = c->in.ensure(16)) != 0) return nullptr;
        return c->in.as<uint8_t>();
    }
    MemKind k = classify(src);
    if (k == kDevice) return static_cast<const uint8_t*>(src);
    if ((*err = c->in.ensure(len + 64)) != 0) return nullptr;
    if (k == kHostPageable && len >= HostStager::kMinBytes) {   // a long malloc'ed buffer: host threads through pinned slots
        HostStager stager;
        if ((*err = stager.start(src, c->in.p, len, s)) != 0) return nullptr;
        if ((*err = stager.wait_range(0, len, s)) != 0) return nullptr;
        return c->in.as<uint8_t>();
    }
    cudaError_t e = cudaMemcpyAsync(c->in.p, src, len, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { set_error("H2D copy of %zu bytes failed: %s", len, cudaGetErrorString(e)); *err = ZB_STREAM_ERROR; return nullptr; }
    return c->in.as<uint8_t>();
}

// ---- HostStager ----
static std::mutex g_slot_mu;
static std::vector<void*> g_slot_pool;                          // pinned slot sets (kThreads * kPiece each), reused across calls

struct HostStager::Impl {
    const uint8_t* src = nullptr;
    uint8_t* d_dst = nullptr;
    size_t n = 0, npieces = 0;
    int device = 0;
    uint8_t* slots = nullptr;
    cudaStream_t st[HostStager::kThreads] = {};
    cudaEvent_t go = nullptr;
    std::vector<cudaEvent_t> ev;
    std::unique_ptr<std::atomic<int>[]> issued;
    std::atomic<int> failed{0};
    std::vector<std::thread> th;

    void run(int t)
    {
        cudaSetDevice(device);
        uint8_t* slot = slots + (size_t)t * HostStager::kPiece;
        if (cudaStreamWaitEvent(st[t], go, 0) != cudaSuccess) failed = 1;
        for (size_t p = (size_t)t; p < npieces; p += HostStager::kThreads) {
            const size_t off = p * HostStager::kPiece, len = std::min(HostStager::kPiece, n - off);
            if (!failed) {
                if (p >= (size_t)HostStager::kThreads && cudaEventSynchronize(ev[p - HostStager::kThreads]) != cudaSuccess) failed = 1;   // the slot's last send
                memcpy(slot, src + off, len);
                if (cudaMemcpyAsync(d_dst + off, slot, len, cudaMemcpyHostToDevice, st[t]) != cudaSuccess) failed = 1;
                if (cudaEventRecord(ev[p], st[t]) != cudaSuccess) failed = 1;
            }
            issued[p].store(1, std::memory_order_release);
        }
    }
};

int HostStager::start(const void* src, void* d_dst, size_t n, cudaStream_t after)
{
    impl = new Impl();
    Impl& m = *impl;
    m.src = (const uint8_t*)src; m.d_dst = (uint8_t*)d_dst; m.n = n;
    m.npieces = (n + kPiece - 1) / kPiece;
    cudaGetDevice(&m.device);
    {
        std::lock_guard<std::mutex> lk(g_slot_mu);
        if (!g_slot_pool.empty()) { m.slots = (uint8_t*)g_slot_pool.back(); g_slot_pool.pop_back(); }
    }
    if (!m.slots && cudaMallocHost((void**)&m.slots, (size_t)kThreads * kPiece) != cudaSuccess) {
        cudaGetLastError(); m.slots = nullptr;
        delete impl; impl = nullptr;
        set_error("pinned staging slots could not be allocated");
        return ZB_MEM_ERROR;
    }
    bool ok = cudaEventCreateWithFlags(&m.go, cudaEventDisableTiming) == cudaSuccess && cudaEventRecord(m.go, after) == cudaSuccess;
    for (int t = 0; t < kThreads && ok; t++) ok = cudaStreamCreateWithFlags(&m.st[t], cudaStreamNonBlocking) == cudaSuccess;
    m.ev.assign(m.npieces, nullptr);
    for (size_t p = 0; p < m.npieces && ok; p++) ok = cudaEventCreateWithFlags(&m.ev[p], cudaEventDisableTiming) == cudaSuccess;
    m.issued.reset(new std::atomic<int>[m.npieces]);
    for (size_t p = 0; p < m.npieces; p++) m.issued[p].store(0, std::memory_order_relaxed);
    if (!ok) { cudaGetLastError(); finish(); set_error("staging streams or events could not be created"); return ZB_MEM_ERROR; }
    for (int t = 0; t < kThreads; t++) m.th.emplace_back([this, t]() { impl->run(t); });
    return 0;
}

int HostStager::wait_range(size_t off, size_t len, cudaStream_t consumer)
{
    if (!impl || len == 0) return 0;
    Impl& m = *impl;
    for (size_t p = off / kPiece; p <= (off + len - 1) / kPiece && p < m.npieces; p++) {
        while (!m.issued[p].load(std::memory_order_acquire)) std::this_thread::yield();
        if (m.failed) { set_error("host staging failed"); return ZB_STREAM_ERROR; }
        if (cudaStreamWaitEvent(consumer, m.ev[p], 0) != cudaSuccess) { set_error("host staging wait failed"); return ZB_STREAM_ERROR; }
    }
    return 0;
}

void HostStager::finish()
{
    if (!impl) return;
    Impl& m = *impl;
    for (auto& t : m.th) if (t.joinable()) t.join();
    for (int t = 0; t < kThreads; t++) if (m.st[t]) { cudaStreamSynchronize(m.st[t]); cudaStreamDestroy(m.st[t]); }
    for (cudaEvent_t e : m.ev) if (e) cudaEventDestroy(e);
    if (m.go) cudaEventDestroy(m.go);
    if (m.slots) { std::lock_guard<std::mutex> lk(g_slot_mu); g_slot_pool.push_back(m.slots); }
    delete impl;
    impl = nullptr;
}

// ---- HostDrainer ----
int HostDrainer::drain(void* dst, const void* d_src, size_t n)
{
    constexpr size_t kPiece = HostStager::kPiece;
    constexpr int kThreads = HostStager::kThreads;
    if (n == 0) return 0;
    if (n < 2 * kPiece) {
        if (cudaMemcpy(dst, d_src, n, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError())); return ZB_STREAM_ERROR; }
        return 0;
    }
    if (!slots) {
        {
            std::lock_guard<std::mutex> lk(g_slot_mu);
            if (!g_slot_pool.empty()) { slots = (uint8_t*)g_slot_pool.back(); g_slot_pool.pop_back(); }
        }
        if (!slots && cudaMallocHost((void**)&slots, (size_t)kThreads * kPiece) != cudaSuccess) {
            cudaGetLastError(); slots = nullptr;
            set_error("pinned staging slots could not be allocated");
            return ZB_MEM_ERROR;
        }
        for (int t = 0; t < kThreads; t++)
            if (cudaStreamCreateWithFlags(&st[t], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); st[t] = nullptr; set_error("staging stream creation failed"); return ZB_MEM_ERROR; }
    }
    int device = 0;
    cudaGetDevice(&device);
    const size_t npieces = (n + kPiece - 1) / kPiece;
    std::atomic<int> failed{0};
    std::vector<std::thread> th;
    for (int t = 0; t < kThreads; t++) {
        th.emplace_back([&, t]() {
            cudaSetDevice(device);
            uint8_t* slot = slots + (size_t)t * kPiece;
            for (size_t p = (size_t)t; p < npieces; p += kThreads) {
                const size_t off = p * kPiece, len = std::min(kPiece, n - off);
                if (cudaMemcpyAsync(slot, (const uint8_t*)d_src + off, len, cudaMemcpyDeviceToHost, st[t]) != cudaSuccess ||
                    cudaStreamSynchronize(st[t]) != cudaSuccess) { failed = 1; return; }
                memcpy((uint8_t*)dst + off, slot, len);
            }
        });
    }
    for (auto& t : th) t.join();
    if (failed) { set_error("D2H staging failed: %s", cudaGetErrorString(cudaGetLastError())); return ZB_STREAM_ERROR; }
    return 0;
}

void HostDrainer::finish()
{
    for (int t = 0; t < HostStager::kThreads; t++) if (st[t]) { cudaStreamDestroy(st[t]); st[t] = nullptr; }
    if (slots) { std::lock_guard<std::mutex> lk(g_slot_mu); g_slot_pool.push_back(slots); slots = nullptr; }
}

}  // namespace zb

using namespace zb;

ZB_API int zb200_init(int device)
{
    if (device >= kMaxDevices) { set_error("zb200_init: device %d out of range", device); return ZB_STREAM_ERROR; }
    if (device >= 0) {
        t_device = device;                                      // binds the calling thread
        DevState& D = g_dev[device];
        int failed = -1;
        D.state.compare_exchange_strong(failed, 0);             // a failed attempt may be retried
    } else {
        const int d = current_device_index();
        if (d >= 0 && d < kMaxDevices) { int failed = -1; g_dev[d].state.compare_exchange_strong(failed, 0); }
    }
    return ensure_init();
}

ZB_API int zb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

ZB_API const char* zb200_last_error(void) { return t_err; }

ZB_API const char* zb200_build_info(void)
{
    return "zb200 " __DATE__ " sm_100a nvcc " ZB_STR2(__CUDACC_VER_MAJOR__) "." ZB_STR2(__CUDACC_VER_MINOR__);
}

ZB_API void* zb200_alloc_pinned(size_t bytes)
{
    if (ensure_init()) return nullptr;
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); set_error("pinned allocation failed"); return nullptr; }
    return p;
}
ZB_API void zb200_free_pinned(void* p) { if (p) cudaFreeHost(p); }

ZB_API void* zb200_alloc_device(size_t bytes)
{
    if (ensure_init()) return nullptr;
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); set_error("device allocation failed"); return nullptr; }
    return p;
}
ZB_API void zb200_free_device(void* p) { if (p) cudaFree(p); }

ZB_API int zb200_copy(void* dst, const void* src, size_t bytes, void* stream)
{
    int rc = ensure_init();
    if (rc) return rc;
    ZB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    ZB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

ZB_API int zb200_copy_async(void* dst, const void* src, size_t bytes, void* stream)
{
    int rc = ensure_init();
    if (rc) return rc;
    ZB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return 0;
}

ZB_API int zb200_host_register(void* host_ptr, size_t bytes)
{
    int rc = ensure_init();
    if (rc) return rc;
    ZB_CUDA(cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));
    return 0;
}

ZB_API int zb200_host_unregister(void* host_ptr)
{
    int rc = ensure_init();
    if (rc) return rc;
    ZB_CUDA(cudaHostUnregister(host_ptr));
    return 0;
}

ZB_API int zb200_ipc_export(void* dev_ptr, unsigned char handle[64])
{
    int rc = ensure_init();
    if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    ZB_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, 64);
    return 0;
}

ZB_API int zb200_ipc_open(const unsigned char handle[64], void** peer_ptr)
{
    int rc = ensure_init();
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    ZB_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

ZB_API int zb200_ipc_close(void* peer_ptr)
{
    int rc = ensure_init();
    if (rc) return rc;
    ZB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}

ZB_API int zb200_sync(void* stream)
{
    int rc = ensure_init();
    if (rc) return rc;
    ZB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

ZB_API uint64_t zb200_kernel_launches(void) { return g_launches.load(); }

ZB_API void zb200_profile(int enable)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
    g_profile = enable != 0;
}

// "kernel=total_ms:launches;..." for everything launched since zb200_profile(1); synchronises the device.
ZB_API int zb200_profile_report(char* out, size_t cap)
{
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<double, int>> acc;
    std::vector<std::string> order;
    for (auto& r : g_prof) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); continue; }
        if (!acc.count(r.name)) order.push_back(r.name);
        acc[r.name].first += ms; acc[r.name].second += 1;
    }
    std::string s;
    for (auto& k : order) {
        char b[160];
        snprintf(b, sizeof(b), "%s=%.6f:%d;", k.c_str(), acc[k].first, acc[k].second);
        s += b;
    }
    if (cap) { snprintf(out, cap, "%s", s.c_str()); }
    return (int)s.size();
}
