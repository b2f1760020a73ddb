// core.cu -- error plumbing, launch counter, device queries.
#include <stdarg.h>

#include "common.cuh"

namespace dpm {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

char *err_buf() { return g_err; }

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches += n; }

int device_sm_count() {
    static thread_local int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;  // B200
        cached = n;
    }
    return cached;
}

}  // namespace dpm

extern "C" int dpm_version(void) { return 100; }
extern "C" const char *dpm_last_error(void) { return dpm::err_buf(); }
extern "C" long long dpm_launch_count(void) { return dpm::g_launches; }
extern "C" void dpm_launch_count_reset(void) { dpm::g_launches = 0; }
