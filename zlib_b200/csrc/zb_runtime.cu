// zb_runtime.cu -- process-wide state of the zb200 engine: device selection, a pool of
// execution contexts (stream + scratch + pinned staging), memory classification and
// the small C-ABI helpers of include/zb200.h.  No codec arithmetic lives here.
#include "zb_common.cuh"

#include <mutex>
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

static std::mutex g_mu;
static int g_state = 0;                    // 0 = untouched, 1 = ready, -1 = failed
static int g_device = -1;
static Ctx* g_free = nullptr;

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

int ensure_init()
{
    if (g_state == 1) {
        if (g_device >= 0) cudaSetDevice(g_device);   // calling thread may be new
        return 0;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_state == 1) return 0;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("zb200: no CUDA device available -- this library has no CPU path");
        g_state = -1;
        return ZB_STREAM_ERROR;
    }
    int dev = 0;
    if (g_device >= 0) dev = g_device; else cudaGetDevice(&dev);
    if (cudaSetDevice(dev) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", dev); g_state = -1; return ZB_STREAM_ERROR; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { set_error("cudaGetDeviceProperties failed"); g_state = -1; return ZB_STREAM_ERROR; }
    if (prop.major != 10) {
        set_error("zb200: device %d is sm_%d%d; this build contains sm_100a code only", dev, prop.major, prop.minor);
        g_state = -1;
        return ZB_STREAM_ERROR;
    }
    g_device = dev;
    if (checksum_setup() != 0) { g_state = -1; return ZB_STREAM_ERROR; }
    g_state = 1;
    return 0;
}

Ctx* ctx_acquire(cudaStream_t use)
{
    Ctx* c = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_free) { c = g_free; g_free = c->next; c->next = nullptr; }
    }
    if (!c) {
        c = new Ctx();
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
    std::lock_guard<std::mutex> lk(g_mu);
    c->next = g_free;
    g_free = c;
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
    cudaError_t e = cudaMemcpyAsync(c->in.p, src, len, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { set_error("H2D copy of %zu bytes failed: %s", len, cudaGetErrorString(e)); *err = ZB_STREAM_ERROR; return nullptr; }
    return c->in.as<uint8_t>();
}

}  // namespace zb

using namespace zb;

ZB_API int zb200_init(int device)
{
    if (device >= 0) {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_state == 1 && g_device != device) { set_error("zb200 already initialised on device %d", g_device); return ZB_STREAM_ERROR; }
        g_device = device;
        if (g_state == -1) g_state = 0;
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
