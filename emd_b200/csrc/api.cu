// C-ABI plumbing shared by all stages: error text, version, device probe.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void emd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* emd_last_error_string() { return g_err; }

extern "C" int emd_abi_version() { return 1; }

// 0 when the current device can run the sm_100a kernels in this library.
extern "C" int emd_device_check() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        emd_set_error("emd_device_check: %s", cudaGetErrorString(e));
        return EMD_ERR_CUDA;
    }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        emd_set_error("emd_device_check: device is sm_%d%d; this library is built for sm_100a only", major, minor);
        return EMD_ERR_UNSUPPORTED;
    }
    return EMD_OK;
}
