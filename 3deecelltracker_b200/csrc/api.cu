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
