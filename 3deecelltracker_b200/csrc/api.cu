// Error reporting, ABI version and launch accounting for libct3d.
#include "common.cuh"
#include <cstring>
#include <mutex>
#include <vector>

namespace ct {
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
thread_local int g_reserved_sms = 0;       // per host thread: a setting is scoped to the launches of the thread that made it

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// pinned staging ring for small uploads (see common.cuh)
// ---------------------------------------------------------------------------------------------
namespace {
struct StageRing {
    static constexpr int SLOTS = 64;
    static constexpr size_t SLOT_BYTES = 8192;
    char* base = nullptr;
    cudaEvent_t ev[SLOTS] = {};
    bool used[SLOTS] = {};
    int next = 0;
    int device = -1;
};
thread_local StageRing t_ring;
}  // namespace

int stage_h2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
    StageRing& r = t_ring;
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
    if (bytes > StageRing::SLOT_BYTES)                                   // rare (very large batches): the pageable path
        return check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync(descriptors)");
    if (r.base == nullptr || r.device != dev) {                          // first use on this thread / device
        if (r.base == nullptr &&
            check_cuda(cudaHostAlloc(reinterpret_cast<void**>(&r.base), StageRing::SLOTS * StageRing::SLOT_BYTES, cudaHostAllocDefault),
                       "cudaHostAlloc(staging ring)")) return 1;
        for (int i = 0; i < StageRing::SLOTS; ++i) {
            if (r.ev[i]) { cudaEventSynchronize(r.ev[i]); cudaEventDestroy(r.ev[i]); r.ev[i] = nullptr; }
            if (check_cuda(cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming), "cudaEventCreate")) return 1;
            r.used[i] = false;
        }
        r.device = dev;
    }
    const int i = r.next;
    r.next = (r.next + 1) % StageRing::SLOTS;
    if (r.used[i] && check_cuda(cudaEventSynchronize(r.ev[i]), "cudaEventSynchronize(staging slot)")) return 1;   // 64 uploads ago
    char* slot = r.base + (size_t)i * StageRing::SLOT_BYTES;
    memcpy(slot, src, bytes);
    if (check_cuda(cudaMemcpyAsync(dst, slot, bytes, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync(descriptors)")) return 1;
    if (check_cuda(cudaEventRecord(r.ev[i], s), "cudaEventRecord")) return 1;
    r.used[i] = true;
    return 0;
}
}  // namespace ct

// ---------------------------------------------------------------------------------------------
// opt-in device-side profiler: CUDA events around tagged launches on the launching stream
// ---------------------------------------------------------------------------------------------
namespace ct {
struct ProfRec { int tag; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;
static std::mutex g_prof_mu;                 // launches may come from several host threads (one stream each)

static cudaEvent_t prof_event() {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(int tag, cudaStream_t s) : tag_(tag), s_(s), on_(g_prof_on) {
    if (!on_) return;
    a_ = prof_event();
    cudaEventRecord(a_, s_);
}
ProfScope::~ProfScope() {
    if (!on_) return;
    cudaEvent_t b = prof_event();
    cudaEventRecord(b, s_);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof.push_back({tag_, a_, b});
}
}  // namespace ct

extern "C" {
int ct_profile_enable(int on) {
    ct::g_prof_on = on != 0;
    return 0;
}

int ct_profile_read(int tag, double* total_ms, unsigned long long* count, int reset) {
    double ms = 0.0;
    unsigned long long n = 0;
    std::lock_guard<std::mutex> lock(ct::g_prof_mu);
    for (auto& r : ct::g_prof) {
        if (r.tag != tag) continue;
        if (cudaEventSynchronize(r.b) != cudaSuccess) { ct::set_error("ct_profile_read: event sync failed"); return 1; }
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        ms += t;
        ++n;
    }
    if (total_ms) *total_ms = ms;
    if (count) *count = n;
    if (reset) {
        std::vector<ct::ProfRec> keep;
        for (auto& r : ct::g_prof) {
            if (r.tag == tag) { ct::g_prof_pool.push_back(r.a); ct::g_prof_pool.push_back(r.b); }
            else keep.push_back(r);
        }
        ct::g_prof.swap(keep);
    }
    return 0;
}

int ct_set_reserved_sms(int n) {
    if (n < 0) n = 0;
    if (n > 64) n = 64;
    const int old = ct::g_reserved_sms;
    ct::g_reserved_sms = n;
    return old;
}

int ct_abi_version(void) { return CT3D_ABI_VERSION; }
const char* ct_last_error(void) { return ct::g_err; }
unsigned long long ct_launch_count(void) { return ct::g_launches.load(); }
}
