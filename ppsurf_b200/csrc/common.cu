#include <stdarg.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace pps {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::mutex g_flag_mutex;
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

static bool g_profile = false;
static std::vector<cudaEvent_t> g_events;  // start/end pairs
static size_t g_used = 0;
static cudaEvent_t next_event() {
    if (g_used == g_events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        g_events.push_back(e);
    }
    return g_events[g_used++];
}
void profile_begin(cudaStream_t st) {
    if (g_profile) cudaEventRecord(next_event(), st);
}
void profile_end(cudaStream_t st) {
    if (g_profile) cudaEventRecord(next_event(), st);
}

bool first_use_on_device(unsigned char* flags) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return true;  // unknown device: (re)configure, harmless
    std::lock_guard<std::mutex> lock(g_flag_mutex);
    if (flags[dev]) return false;
    flags[dev] = 1;
    return true;
}

// per-device pool of timing-disabled events for the host-buffer entry points (created once: the C ABI promises no
// allocation per call)
static std::mutex g_event_mutex;
static cudaEvent_t g_dev_events[64][8];
static int g_dev_event_count[64] = {0};
cudaEvent_t* device_events(int count) {
    int dev = 0;
    if (count > 8 || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        set_error("device_events: bad device or count");
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(g_event_mutex);
    while (g_dev_event_count[dev] < count) {
        cudaError_t e = cudaEventCreateWithFlags(&g_dev_events[dev][g_dev_event_count[dev]], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            set_error("device_events: %s", cudaGetErrorString(e));
            return nullptr;
        }
        ++g_dev_event_count[dev];
    }
    return g_dev_events[dev];
}
}  // namespace pps

#ifndef PPS_SOURCE_HASH
#define PPS_SOURCE_HASH 0ull
#endif

extern "C" {
unsigned long long pps_source_hash(void) { return PPS_SOURCE_HASH; }
unsigned long long pps_launch_count(void) { return pps::g_launches; }
void pps_profile_enable(int on) {
    pps::g_profile = on != 0;
    pps::g_used = 0;
}
int pps_profile_read(double* total_ms, long long* brackets) {
    double sum = 0.0;
    for (size_t i = 0; i + 1 < pps::g_used; i += 2) {
        if (cudaEventSynchronize(pps::g_events[i + 1]) != cudaSuccess) return PPS_ERR_CUDA;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pps::g_events[i], pps::g_events[i + 1]) != cudaSuccess) return PPS_ERR_CUDA;
        sum += ms;
    }
    if (total_ms) *total_ms = sum;
    if (brackets) *brackets = (long long)(pps::g_used / 2);
    pps::g_used = 0;
    return PPS_OK;
}
const char* pps_last_error(void) { return pps::g_err; }
int pps_version(void) { return 200; }
int pps_compiled_arch(void) { return 100; }
int pps_check_device(void) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        pps::set_error("no CUDA device: ppsurf_b200 has no CPU fallback");
        cudaGetLastError();
        return PPS_ERR_NO_DEVICE;
    }
    if (prop.major != 10) {
        pps::set_error("device %s is sm_%d%d; the kernels are built for sm_100a only", prop.name, prop.major, prop.minor);
        return PPS_ERR_NO_DEVICE;
    }
    return PPS_OK;
}
}
