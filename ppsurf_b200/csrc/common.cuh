// Shared helpers for the ppsurf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ppsurf_b200.h"

namespace pps {

void set_error(const char* fmt, ...);
void count_launch();
// optional CUDA-event bracket around the dominant kernel (bench.py's roofline); no-ops unless pps_profile_enable(1)
void profile_begin(cudaStream_t st);
void profile_end(cudaStream_t st);
// `count` (<= 8) timing-disabled events of the current device, created on first use and kept for the process
cudaEvent_t* device_events(int count);

#define PPS_CHECK_ARG(cond, ...)                \
    do {                                        \
        if (!(cond)) {                          \
            ::pps::set_error(__VA_ARGS__);      \
            return PPS_ERR_INVALID;             \
        }                                       \
    } while (0)

#define PPS_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            ::pps::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return PPS_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

// every kernel launch is followed by PPS_LAUNCH_CHECK(): it also feeds pps_launch_count()
#define PPS_LAUNCH_CHECK()              \
    do {                                \
        ::pps::count_launch();          \
        PPS_CUDA(cudaGetLastError());   \
    } while (0)

#define PPS_TRY(call)            \
    do {                         \
        int s__ = (call);        \
        if (s__ != 0) return s__; \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// bump allocator over a caller-owned workspace
struct Arena {
    char* base;
    size_t size;
    size_t off;
    Arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* r = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= size; }
};

constexpr int kNumSMs = 148;  // B200

// true the first time it is called for `flags` on the CURRENT device (function attributes such as the dynamic shared-memory limit are
// per device: a process driving several GPUs must set them on each).  flags: a zero-initialised static array of kMaxDevices bytes.
constexpr int kMaxDevices = 64;
bool first_use_on_device(unsigned char* flags);

}  // namespace pps
