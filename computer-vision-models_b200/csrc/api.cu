// Error plumbing and device queries shared by all entry points of libcvmhot.so.
#include <stdarg.h>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
}

void cvm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cvm_num_sms() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
        cached = n;
        return n;
    }
    (void)cudaGetLastError();
    return CVM_NUM_SMS_FALLBACK;  // no device visible (e.g. workspace-size queries on a CPU box): B200 value
}

extern "C" const char* cvm_last_error(void) { return g_err; }

extern "C" int cvm_version(void) { return 100; }
