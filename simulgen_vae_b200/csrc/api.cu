// Error reporting and device probing for the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace sg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

}  // namespace sg

extern "C" {

const char* sg_last_error(void) { return sg::g_err; }

int sg_version(void) { return 1; }

int sg_device_supported(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    return prop.major == 10 ? 1 : 0;
}

}  // extern "C"
