// Error reporting, version and device queries of the C ABI (include/marconet_b200.h).
#include <cstdlib>
#include <stdarg.h>
#include <string.h>
#include "mn_common.cuh"

static thread_local char g_err[512] = "";

void mn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int mn_num_sms() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

static int g_max_ctas = 0;
int mn_max_ctas() { return g_max_ctas; }
extern "C" int mn_set_max_ctas(int n) {
    const int old = g_max_ctas;
    g_max_ctas = n > 0 ? n : 0;
    return old;
}

static int g_pdl = -1;
int mn_pdl_enabled() {
    if (g_pdl < 0) { const char* e = getenv("MN_PDL"); g_pdl = (e && e[0] == '0') ? 0 : 1; }
    return g_pdl;
}
extern "C" int mn_set_pdl(int on) {
    const int old = mn_pdl_enabled();
    g_pdl = on ? 1 : 0;
    return old;
}

extern "C" const char* mn_last_error(void) { return g_err; }
extern "C" int mn_version(void) { return 100; }
extern "C" int mn_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10;
}
