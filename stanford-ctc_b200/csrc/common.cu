// common.cu -- error string, launch counter, device queries.
#include "common.cuh"
#include <stdarg.h>

namespace ctcb {
thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}
}  // namespace ctcb

extern "C" int ctcb_version(void) { return 100; }
extern "C" const char *ctcb_last_error(void) { return ctcb::g_last_error; }
extern "C" uint64_t ctcb_launch_count(void) { return ctcb::g_launch_count.load(); }
