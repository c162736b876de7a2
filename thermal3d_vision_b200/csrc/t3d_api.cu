// t3d_api.cu -- library-level entry points of libt3d_sm100.so (error string,
// version, device info, launch counter).  See include/t3d.h.
#include "t3d_common.cuh"

#include <atomic>
#include <mutex>
#include <vector>
#include <string.h>

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

// ---------------------------------------------------------------- per-kernel event timing
namespace {
struct Prof {
    std::mutex mu;
    bool active = false;
    char pattern[96] = "";
    std::vector<cudaEvent_t> ev;   // start/stop pairs
    std::vector<const char*> names; // kernel name of each pair (string literals)
    int used = 0, cap = 0;
    int stride = 1, seen = 0;       // only every stride-th matching launch is bracketed
} g_prof;
}  // namespace

// `pattern`: one substring, or several separated by '|'
static bool prof_matches(const char* name, const char* pattern) {
    if (pattern[0] == 0) return true;               // empty pattern: every kernel
    const char* p = pattern;
    for (;;) {
        const char* bar = strchr(p, '|');
        const size_t len = bar ? (size_t)(bar - p) : strlen(p);
        if (len > 0 && len < 64) {
            char sub[64];
            memcpy(sub, p, len); sub[len] = 0;
            if (strstr(name, sub)) return true;
        }
        if (!bar) return false;
        p = bar + 1;
    }
}

bool t3d_prof_before(const char* name, cudaStream_t st) {
    if (!g_prof.active) return false;
    std::lock_guard<std::mutex> lk(g_prof.mu);
    if (!g_prof.active || g_prof.used >= g_prof.cap || !prof_matches(name, g_prof.pattern)) return false;
    if ((g_prof.seen++ % g_prof.stride) != 0) return false;
    cudaEventRecord(g_prof.ev[2 * g_prof.used], st);
    g_prof.names[g_prof.used] = name;
    return true;
}

void t3d_prof_after(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    if (!g_prof.active || g_prof.used >= g_prof.cap) return;
    cudaEventRecord(g_prof.ev[2 * g_prof.used + 1], st);
    ++g_prof.used;
}

void t3d_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int t3d_sm_count() {
    std::call_once(g_dev_once, init_device_props);
    return g_sm_count > 0 ? g_sm_count : 148;
}

int t3d_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev % kT3dMaxDevices;
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

int t3d_profile_begin(const char* kernel_name_substr, int max_launches) {
    T3D_REQUIRE(kernel_name_substr && max_launches > 0 && max_launches <= (1 << 20), "bad arguments");
    std::lock_guard<std::mutex> lk(g_prof.mu);
    for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
    g_prof.ev.assign((size_t)2 * max_launches, nullptr);
    for (auto& e : g_prof.ev) T3D_CUDA(cudaEventCreate(&e));
    g_prof.names.assign((size_t)max_launches, nullptr);
    strncpy(g_prof.pattern, kernel_name_substr, sizeof(g_prof.pattern) - 1);
    g_prof.pattern[sizeof(g_prof.pattern) - 1] = 0;
    g_prof.used = 0; g_prof.cap = max_launches; g_prof.active = true;
    g_prof.stride = 1; g_prof.seen = 0;
    return T3D_OK;
}

int t3d_profile_stride(int every_nth) {
    T3D_REQUIRE(every_nth >= 1, "bad stride");
    std::lock_guard<std::mutex> lk(g_prof.mu);
    g_prof.stride = every_nth;
    return T3D_OK;
}

int t3d_profile_timeline(char* names, int name_stride, double* start_ms, double* stop_ms, int cap, int* launches) {
    T3D_REQUIRE(names && name_stride > 1 && start_ms && stop_ms && cap > 0 && launches, "bad arguments");
    std::lock_guard<std::mutex> lk(g_prof.mu);
    g_prof.active = false;
    const int n = g_prof.used < cap ? g_prof.used : cap;
    for (int i = 0; i < n; ++i) {
        T3D_CUDA(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
        float a = 0.f, b = 0.f;
        T3D_CUDA(cudaEventElapsedTime(&a, g_prof.ev[0], g_prof.ev[2 * i]));
        T3D_CUDA(cudaEventElapsedTime(&b, g_prof.ev[0], g_prof.ev[2 * i + 1]));
        start_ms[i] = a; stop_ms[i] = b;
        strncpy(names + (size_t)i * name_stride, g_prof.names[i] ? g_prof.names[i] : "", name_stride - 1);
        names[(size_t)i * name_stride + name_stride - 1] = 0;
    }
    *launches = n;
    return T3D_OK;
}

int t3d_profile_end(double* total_ms, int* launches) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    g_prof.active = false;
    double tot = 0.0;
    for (int i = 0; i < g_prof.used; ++i) {
        T3D_CUDA(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
        float ms = 0.f;
        T3D_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = g_prof.used;
    for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
    g_prof.ev.clear(); g_prof.used = 0; g_prof.cap = 0;
    return T3D_OK;
}

}  // extern "C"
