// multibox_b200 -- C-ABI plumbing: version, error string, device queries.
#include <cstdarg>
#include <cstdio>

#include "mbx_common.cuh"

namespace mbx {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return static_cast<int>(e);
}

struct DevInfo {
    int dev = -1, sms = 0, smem = 0;
};
static thread_local DevInfo g_dev;

static void refresh_dev() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    if (dev == g_dev.dev) return;
    g_dev.dev = dev;
    cudaDeviceGetAttribute(&g_dev.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_dev.smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
}

int sm_count() {
    refresh_dev();
    return g_dev.sms > 0 ? g_dev.sms : 148;
}
int max_smem_optin() {
    refresh_dev();
    return g_dev.smem > 0 ? g_dev.smem : 232448;
}

}  // namespace mbx

extern "C" int mbx_version(void) { return MBX_VERSION; }
extern "C" const char *mbx_last_error(void) { return mbx::g_err; }
extern "C" int mbx_device_info(int *sm_count, int *max_smem_per_block) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        mbx::set_error("mbx_device_info: no CUDA device (%s)", cudaGetErrorString(e));
        return e != cudaSuccess ? static_cast<int>(e) : MBX_E_ARG;
    }
    if (sm_count) *sm_count = mbx::sm_count();
    if (max_smem_per_block) *max_smem_per_block = mbx::max_smem_optin();
    return 0;
}
