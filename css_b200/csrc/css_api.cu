// Library-level entry points: version, thread-local error text, device properties.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "css_common.cuh"

static thread_local char g_err[512] = "";

void css_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int css_cached_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static int g_pdl = -1;
bool css_pdl_enabled() {
    if (g_pdl < 0) {
        const char* v = getenv("CSS_B200_PDL");
        g_pdl = (v && v[0] == '1') ? 1 : 0;      // measured on B200 (r02i): no gain inside CUDA graphs (0.478 vs 0.475 ms/step), so opt-in
    }
    return g_pdl != 0;
}
extern "C" int css_set_pdl(int on) {
    g_pdl = on ? 1 : 0;
    return 0;
}

static unsigned long long g_launches = 0;
void css_count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
extern "C" unsigned long long css_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int css_version(void) { return CSS_B200_VERSION; }
extern "C" const char* css_last_error(void) { return g_err; }
extern "C" int css_sm_count(void) {
    int dev = 0, n = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
        css_set_error("css_sm_count: %s", cudaGetErrorString(e));
        return -(int)e;
    }
    return n;
}
