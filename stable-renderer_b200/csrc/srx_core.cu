// srx_core.cu — error plumbing, version, device queries.
#include "srx_common.cuh"

static thread_local char g_err[512] = "";

int srx_set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int srx_sm_count_cached() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

extern "C" {
int srx_version(void) { return SRX_VERSION; }
const char *srx_last_error(void) { return g_err; }
int srx_device_sm_count(void) { return srx_sm_count_cached(); }
}
