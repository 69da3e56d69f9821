// Error reporting, ABI version and launch accounting for libct3d.
#include "common.cuh"
#include <cstring>

namespace ct {
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace ct

extern "C" {
int ct_abi_version(void) { return CT3D_ABI_VERSION; }
const char* ct_last_error(void) { return ct::g_err; }
unsigned long long ct_launch_count(void) { return ct::g_launches.load(); }
}
