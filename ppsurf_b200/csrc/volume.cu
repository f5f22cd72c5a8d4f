// Region-growing bookkeeping of the occupancy volume on the device.
//
// Replaces the mask handling of _create_volume (source/poco_utils.py:178-254): the per-point Python dilation loop
// (181-196), np.argwhere over the (res+2)^3 mask (210-213), the scatter `volume[mask] = z` (229) and the sign-change
// frontier (232-244).  Voxels are addressed by their C-order linear index; lists come out in ascending index order
// (= np.argwhere order) through cub::DeviceSelect.  A voxel decoded in an earlier sweep is not decoded again: the decode
// is deterministic per query, so the reference's re-evaluation writes the same value.
#include <cub/cub.cuh>

#include "common.cuh"

namespace pps {

// every seed stamps the clipped box [p-d, p+d]^3 (poco_utils.py:185-191)
__global__ void volume_stamp_kernel(const int32_t* __restrict__ seeds, long long n, int r, int d, uint8_t* __restrict__ mask) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int side = 2 * d + 1;
    const long long per = (long long)side * side;
    if (e >= n * per) return;
    const long long s = e / per;
    const int o = int(e % per);
    const int v = seeds[s];
    const int iz = v % r, iy = (v / r) % r, ix = v / (r * r);
    const int x = ix + o / side - d, y = iy + o % side - d;
    if (x < 0 || x >= r || y < 0 || y >= r) return;
    const int z0 = max(iz - d, 0), z1 = min(iz + d, r - 1);
    uint8_t* row = mask + ((long long)x * r + y) * r;
    for (int z = z0; z <= z1; ++z) row[z] = 1;
}

// seeds are taken off `to_see`; a seed with value <= 0 stamps mask_a, one with value >= 0 stamps mask_b (poco_utils.py:232-239)
__global__ void volume_stamp_signed_kernel(const int32_t* __restrict__ seeds, long long n, int r, int d,
                                           const float* __restrict__ volume, uint8_t* __restrict__ to_see,
                                           uint8_t* __restrict__ mask_a, uint8_t* __restrict__ mask_b) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int side = 2 * d + 1;
    const long long per = (long long)side * side;
    if (e >= n * per) return;
    const long long s = e / per;
    const int o = int(e % per);
    const int v = seeds[s];
    const float val = volume[v];
    if (o == 0) to_see[v] = 0;
    const bool a = val <= 0.f, b = val >= 0.f;  // NaN: neither
    if (!a && !b) return;
    const int iz = v % r, iy = (v / r) % r, ix = v / (r * r);
    const int x = ix + o / side - d, y = iy + o % side - d;
    if (x < 0 || x >= r || y < 0 || y >= r) return;
    const int z0 = max(iz - d, 0), z1 = min(iz + d, r - 1);
    const long long base = ((long long)x * r + y) * r;
    for (int z = z0; z <= z1; ++z) {
        if (a) mask_a[base + z] = 1;
        if (b) mask_b[base + z] = 1;
    }
}

struct PendingFlag {  // in the mask and not decoded yet
    const uint8_t* mask;
    const float* volume;
    __device__ bool operator()(int i) const { return mask[i] && isnan(volume[i]); }
};
struct FrontierFlag {  // (mask_neg & volume >= 0 & to_see) | (mask_pos & volume <= 0 & to_see)   (poco_utils.py:243)
    const uint8_t* mask_a;
    const uint8_t* mask_b;
    const uint8_t* to_see;
    const float* volume;
    __device__ bool operator()(int i) const {
        const float v = volume[i];
        return to_see[i] && ((mask_a[i] && v >= 0.f) || (mask_b[i] && v <= 0.f));
    }
};

__global__ void volume_queries_kernel(const int32_t* __restrict__ ids, long long n, int r, float step, float bmin_pad,
                                      float* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int v = ids[e];
    const int iz = v % r, iy = (v / r) % r, ix = v / (r * r);
    out[3 * e + 0] = __fadd_rn(__fmul_rn(float(ix), step), bmin_pad);  // same two fp32 roundings as poco_utils.py:213
    out[3 * e + 1] = __fadd_rn(__fmul_rn(float(iy), step), bmin_pad);
    out[3 * e + 2] = __fadd_rn(__fmul_rn(float(iz), step), bmin_pad);
}

__global__ void volume_scatter_kernel(const int32_t* __restrict__ ids, const float* __restrict__ vals, long long n,
                                      float* __restrict__ volume) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) volume[ids[e]] = vals[e];
}

__global__ void volume_init_kernel(float* volume, uint8_t* to_see, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    volume[e] = __int_as_float(0x7fc00000);
    to_see[e] = 1;
}

__global__ void volume_border_kernel(float* volume, int r, int padding, float out_value) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)r * r * r) return;
    const int iz = int(e % r), iy = int((e / r) % r), ix = int(e / ((long long)r * r));
    if (ix < padding || ix >= r - padding || iy < padding || iy >= r - padding || iz < padding || iz >= r - padding)
        volume[e] = out_value;
}

struct RegionLayout {
    size_t mask_a, mask_b, temp, temp_bytes, total;
};
static RegionLayout region_layout(int r) {
    RegionLayout l;
    const size_t total = (size_t)r * r * r;
    size_t off = 0;
    l.mask_a = off;
    off = align_up(off + total, 256);
    l.mask_b = off;
    off = align_up(off + total, 256);
    size_t need = 0;
    cub::CountingInputIterator<int> it(0);
    cub::DeviceSelect::If(nullptr, need, it, (int32_t*)nullptr, (long long*)nullptr, (int)total,
                          FrontierFlag{nullptr, nullptr, nullptr, nullptr});
    l.temp = off;
    l.temp_bytes = need + 256;
    off = align_up(off + l.temp_bytes, 256);
    l.total = off;
    return l;
}

}  // namespace pps

using namespace pps;

extern "C" {

size_t pps_region_workspace_bytes(int r) { return r > 0 ? region_layout(r).total : 0; }

int pps_region_init(int r, float* volume, uint8_t* to_see, void* stream) {
    PPS_CHECK_ARG(volume && to_see && r > 0 && r <= 1024, "pps_region_init: bad arguments");
    const long long total = (long long)r * r * r;
    volume_init_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(volume, to_see, total);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_region_pending(const int32_t* seeds, int64_t n_seeds, int r, int dilation, const float* volume, void* workspace,
                       size_t workspace_bytes, int32_t* out_ids, long long* out_count, void* stream) {
    PPS_CHECK_ARG(seeds && volume && workspace && out_ids && out_count && r > 0 && r <= 1024 && dilation >= 0 && n_seeds >= 0,
                  "pps_region_pending: bad arguments");
    RegionLayout l = region_layout(r);
    if (workspace_bytes < l.total) {
        set_error("pps_region_pending: workspace %zu < required %zu", workspace_bytes, l.total);
        return PPS_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(workspace);
    uint8_t* mask = reinterpret_cast<uint8_t*>(base + l.mask_a);
    const long long total = (long long)r * r * r;
    PPS_CUDA(cudaMemsetAsync(mask, 0, total, st));
    if (n_seeds > 0) {
        const long long work = n_seeds * (2 * dilation + 1) * (2 * dilation + 1);
        volume_stamp_kernel<<<(unsigned)ceil_div(work, 256), 256, 0, st>>>(seeds, n_seeds, r, dilation, mask);
        PPS_LAUNCH_CHECK();
    }
    cub::CountingInputIterator<int> it(0);
    size_t temp = l.temp_bytes;
    PPS_CUDA(cub::DeviceSelect::If(base + l.temp, temp, it, out_ids, out_count, (int)total, PendingFlag{mask, volume}, st));
    count_launch();
    return PPS_OK;
}

int pps_region_frontier(const int32_t* seeds, int64_t n_seeds, int r, int dilation, const float* volume, uint8_t* to_see,
                        void* workspace, size_t workspace_bytes, int32_t* out_ids, long long* out_count, void* stream) {
    PPS_CHECK_ARG(seeds && volume && to_see && workspace && out_ids && out_count && r > 0 && r <= 1024 && dilation >= 0 &&
                      n_seeds >= 0,
                  "pps_region_frontier: bad arguments");
    PPS_CHECK_ARG(out_ids != seeds, "pps_region_frontier: the new frontier must not overwrite the seed list it is derived from");
    RegionLayout l = region_layout(r);
    if (workspace_bytes < l.total) {
        set_error("pps_region_frontier: workspace %zu < required %zu", workspace_bytes, l.total);
        return PPS_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(workspace);
    uint8_t* mask_a = reinterpret_cast<uint8_t*>(base + l.mask_a);
    uint8_t* mask_b = reinterpret_cast<uint8_t*>(base + l.mask_b);
    const long long total = (long long)r * r * r;
    PPS_CUDA(cudaMemsetAsync(mask_a, 0, total, st));
    PPS_CUDA(cudaMemsetAsync(mask_b, 0, total, st));
    if (n_seeds > 0) {
        const long long work = n_seeds * (2 * dilation + 1) * (2 * dilation + 1);
        volume_stamp_signed_kernel<<<(unsigned)ceil_div(work, 256), 256, 0, st>>>(seeds, n_seeds, r, dilation, volume, to_see, mask_a,
                                                                                  mask_b);
        PPS_LAUNCH_CHECK();
    }
    cub::CountingInputIterator<int> it(0);
    size_t temp = l.temp_bytes;
    PPS_CUDA(cub::DeviceSelect::If(base + l.temp, temp, it, out_ids, out_count, (int)total,
                                   FrontierFlag{mask_a, mask_b, to_see, volume}, st));
    count_launch();
    return PPS_OK;
}

int pps_region_queries(const int32_t* ids, int64_t n, int r, float step, float bmin_pad, float* out, void* stream) {
    PPS_CHECK_ARG(ids && out && n >= 0 && r > 0, "pps_region_queries: bad arguments");
    if (n == 0) return PPS_OK;
    volume_queries_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ids, n, r, step, bmin_pad, out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_region_scatter(const int32_t* ids, const float* values, int64_t n, float* volume, void* stream) {
    PPS_CHECK_ARG(ids && values && volume && n >= 0, "pps_region_scatter: bad arguments");
    if (n == 0) return PPS_OK;
    volume_scatter_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ids, values, n, volume);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_region_finish(float* volume, int r, int padding, float out_value, void* stream) {
    PPS_CHECK_ARG(volume && r > 0 && padding >= 0 && 2 * padding <= r, "pps_region_finish: bad arguments");
    if (padding == 0) return PPS_OK;
    const long long total = (long long)r * r * r;
    volume_border_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(volume, r, padding, out_value);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
}
