// t3d_api.cu -- library-level entry points of libt3d_sm100.so (error string,
// version, device info, launch counter).  See include/t3d.h.
#include "t3d_common.cuh"

#include <atomic>
#include <mutex>

namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::once_flag g_dev_once;
int g_sm_count = 0, g_cc_major = 0, g_cc_minor = 0;

void init_device_props() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&g_cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
}
}  // namespace

void t3d_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void t3d_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int t3d_sm_count() {
    std::call_once(g_dev_once, init_device_props);
    return g_sm_count > 0 ? g_sm_count : 148;
}

extern "C" {

int t3d_version(void) { return T3D_ABI_VERSION; }

const char* t3d_last_error(void) { return g_err; }

int t3d_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    std::call_once(g_dev_once, init_device_props);
    if (g_sm_count <= 0) {
        t3d_set_error("no CUDA device visible");
        return T3D_ERR_DEVICE;
    }
    if (sm_count) *sm_count = g_sm_count;
    if (cc_major) *cc_major = g_cc_major;
    if (cc_minor) *cc_minor = g_cc_minor;
    return T3D_OK;
}

uint64_t t3d_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
