// Library-level entry points: version, thread-local error string, launch counter.
#include "common.cuh"

namespace ups {
static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch() { ++g_launches; }
int after_launch(const char* what) {
    ++g_launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return UPS_E_CUDA;
    }
    return UPS_OK;
}
}  // namespace ups

#ifndef UPS_SRC_HASH
#define UPS_SRC_HASH "unknown"
#endif
extern "C" const char* ups_version(void) { return "ups_b200 0.2.0 (sm_100a, src=" UPS_SRC_HASH ")"; }
extern "C" const char* ups_last_error_string(void) { return ups::g_err; }
extern "C" long long ups_launch_count(void) { return ups::g_launches; }
extern "C" void ups_launch_count_reset(void) { ups::g_launches = 0; }
