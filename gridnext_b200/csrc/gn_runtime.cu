// Library-wide state: version, thread-local error string, cached device properties.
#include "gn_common.cuh"
#include <stdarg.h>
#include <mutex>

static thread_local char g_err[512] = "";

void gn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

GN_API const char* gn_last_error(void) { return g_err; }

GN_API int gn_version(void) { return 100; }  // 0.1.0

int gn_num_sms() {
    static int sms = 0;
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) { sms = 148; return; }
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    });
    return sms;
}

GN_API int gn_device_sm_count(void) { return gn_num_sms(); }
