// Device marching cubes and bisection refinement of the mesh vertices: the steps that follow the occupancy
// volume in export_mesh_and_refine_vertices_region_growing_v3 (source/poco_utils.py:87-168).  The reference calls
// skimage.measure.marching_cubes (Cython, CPU) on a host copy of the volume and runs ten decode sweeps with a host round trip of
// all vertices per sweep; here the volume never leaves the device, vertices are created once per crossed GRID EDGE (shared by the
// up to four cells around it: the merge_vertices pass of the reference's mesh cleaning is implicit and exact) and the refinement
// state (bracketing endpoints, their values, the current vertex) stays on the device for all sweeps.
//
// Cells with a NaN corner (voxels the region growing never decoded) produce no triangles: the reference's marching cubes yields
// NaN vertices there, which its clean_simple_inplace -> remove_infinite_values drops with their faces.
// Case table: ppsurf_b200/mc_tables.py (generated; bit i of the case = corner i below the level = inside).
#include <cub/cub.cuh>

#include "common.cuh"

namespace pps {

__constant__ int8_t c_edge_corner[12][2];
__constant__ int8_t c_edge_axis[12];
__constant__ int8_t c_edge_origin[12][3];

struct McLayout {
    size_t tri_count, tri_off, edge_flag, edge_off, temp, temp_bytes, total;
};
static McLayout mc_layout(int r) {
    McLayout l;
    const size_t cells = (size_t)(r - 1) * (r - 1) * (r - 1), edges = 3 * (size_t)r * r * r;
    size_t off = 0;
    l.tri_count = off;
    off = align_up(off + (cells + 1) * 4, 256);
    l.tri_off = off;
    off = align_up(off + (cells + 1) * 4, 256);
    l.edge_flag = off;
    off = align_up(off + (edges + 1) * 4, 256);
    l.edge_off = off;
    off = align_up(off + (edges + 1) * 4, 256);
    size_t t1 = 0, t2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t1, (int*)nullptr, (int*)nullptr, (int)(cells + 1));
    cub::DeviceScan::ExclusiveSum(nullptr, t2, (int*)nullptr, (int*)nullptr, (int)(edges + 1));
    l.temp = off;
    l.temp_bytes = t1 > t2 ? t1 : t2;
    off = align_up(off + l.temp_bytes, 256);
    l.total = off;
    return l;
}

__device__ __forceinline__ int mc_case(const float* __restrict__ vol, int r, int x, int y, int z, float level, float (&v)[8]) {
    int c = 0;
    bool nan = false;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = vol[((size_t)(x + (i & 1)) * r + (y + ((i >> 1) & 1))) * r + (z + ((i >> 2) & 1))];
        nan |= isnan(v[i]);
        c |= (v[i] < level) ? (1 << i) : 0;
    }
    return nan ? 0 : c;
}

// pass 1: triangles per cell, crossed grid edges
__global__ void mc_count_kernel(const float* __restrict__ vol, int r, float level, const int8_t* __restrict__ table, int width,
                                int* __restrict__ tri_count, int* __restrict__ edge_flag) {
    const long long cells = (long long)(r - 1) * (r - 1) * (r - 1);
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= cells) return;
    const int z = (int)(cell % (r - 1)), y = (int)((cell / (r - 1)) % (r - 1)), x = (int)(cell / ((long long)(r - 1) * (r - 1)));
    float v[8];
    const int c = mc_case(vol, r, x, y, z, level, v);
    const int8_t* row = table + c * width;
    int n = 0;
    while (n < width && row[n] >= 0) {
        const int e = row[n];
        const size_t g = ((size_t)(x + c_edge_origin[e][0]) * r + (y + c_edge_origin[e][1])) * r + (z + c_edge_origin[e][2]);
        edge_flag[3 * g + c_edge_axis[e]] = 1;
        ++n;
    }
    tri_count[cell] = n / 3;
}

// pass 2a: faces with the compacted vertex numbers of their edges
__global__ void mc_faces_kernel(const float* __restrict__ vol, int r, float level, const int8_t* __restrict__ table, int width,
                                const int* __restrict__ tri_off, const int* __restrict__ edge_off, int32_t* __restrict__ faces) {
    const long long cells = (long long)(r - 1) * (r - 1) * (r - 1);
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= cells) return;
    const int z = (int)(cell % (r - 1)), y = (int)((cell / (r - 1)) % (r - 1)), x = (int)(cell / ((long long)(r - 1) * (r - 1)));
    float v[8];
    const int c = mc_case(vol, r, x, y, z, level, v);
    const int8_t* row = table + c * width;
    int32_t* dst = faces + 3 * (size_t)tri_off[cell];
    for (int n = 0; n < width && row[n] >= 0; ++n) {
        const int e = row[n];
        const size_t g = ((size_t)(x + c_edge_origin[e][0]) * r + (y + c_edge_origin[e][1])) * r + (z + c_edge_origin[e][2]);
        dst[n] = edge_off[3 * g + c_edge_axis[e]];
    }
}

// pass 2b: one vertex per crossed grid edge, linear interpolation of the level in volume-index coordinates
__global__ void mc_verts_kernel(const float* __restrict__ vol, int r, float level, const int* __restrict__ edge_flag,
                                const int* __restrict__ edge_off, float* __restrict__ verts, int32_t* __restrict__ vert_edge) {
    const long long edges = 3ll * r * r * r;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= edges || !edge_flag[e]) return;
    const long long g = e / 3;
    const int axis = (int)(e % 3);
    const int z = (int)(g % r), y = (int)((g / r) % r), x = (int)(g / ((long long)r * r));
    const long long stride = axis == 0 ? (long long)r * r : (axis == 1 ? r : 1);
    const float va = vol[g], vb = vol[g + stride];
    float t = (level - va) / (vb - va);
    t = fminf(fmaxf(t, 0.f), 1.f);
    const int o = edge_off[e];
    verts[3 * (size_t)o] = x + (axis == 0 ? t : 0.f);
    verts[3 * (size_t)o + 1] = y + (axis == 1 ? t : 0.f);
    verts[3 * (size_t)o + 2] = z + (axis == 2 ? t : 0.f);
    vert_edge[o] = (int32_t)e;
}

// bisection state of the refinement (source/poco_utils.py:111-140): only vertices strictly inside their grid edge are refined
// (exactly one non-integer coordinate), and only when both end values exist
__global__ void refine_init_kernel(const float* __restrict__ vol, int r, const float* __restrict__ verts, const int32_t* __restrict__ vert_edge,
                                   long long nv, float step, float bmin_pad, float* __restrict__ va, float* __restrict__ vb,
                                   float* __restrict__ pa, float* __restrict__ pb, float* __restrict__ v, unsigned char* __restrict__ active) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const long long e = vert_edge[i];
    const long long g = e / 3;
    const int axis = (int)(e % 3);
    const int c[3] = {(int)(g / ((long long)r * r)), (int)((g / r) % r), (int)(g % r)};
    const long long stride = axis == 0 ? (long long)r * r : (axis == 1 ? r : 1);
    const float fa = vol[g], fb = vol[g + stride];
    const float p = verts[3 * i + axis];
    const bool inside_edge = p - floorf(p) > 0.f;
    active[i] = inside_edge && !isnan(fa) && !isnan(fb);
    pa[i] = fa;
    pb[i] = fb;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        // separately rounded multiply and add like numpy's `idx * step + bmin_pad` (no FMA contraction: bit-equal coordinates)
        va[3 * i + d] = __fadd_rn(__fmul_rn(float(c[d]), step), bmin_pad);
        vb[3 * i + d] = __fadd_rn(__fmul_rn(float(c[d] + (d == axis ? 1 : 0)), step), bmin_pad);
        v[3 * i + d] = __fadd_rn(__fmul_rn(verts[3 * i + d], step), bmin_pad);
    }
}

// one sweep (poco_utils.py:157-166): the end whose value has the sign of the prediction moves to the current vertex
__global__ void refine_update_kernel(const float* __restrict__ pred, long long n, float* __restrict__ va, float* __restrict__ vb,
                                     float* __restrict__ pa, float* __restrict__ pb, float* __restrict__ v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p = pred[i];
    const bool m1 = p * pa[i] > 0.f, m2 = p * pb[i] > 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float cur = v[3 * i + d];
        const float a = m1 ? cur : va[3 * i + d], b = m2 ? cur : vb[3 * i + d];
        va[3 * i + d] = a;
        vb[3 * i + d] = b;
        v[3 * i + d] = __fadd_rn(b, a) * 0.5f;
    }
    if (m1) pa[i] = p;
    if (m2) pb[i] = p;
}

static int mc_upload_tables(const int8_t* edge_corner, const int8_t* edge_axis, const int8_t* edge_origin) {
    PPS_CUDA(cudaMemcpyToSymbol(c_edge_corner, edge_corner, 24));
    PPS_CUDA(cudaMemcpyToSymbol(c_edge_axis, edge_axis, 12));
    PPS_CUDA(cudaMemcpyToSymbol(c_edge_origin, edge_origin, 36));
    return PPS_OK;
}

}  // namespace pps

using namespace pps;

extern "C" {

int pps_mc_set_edges(const int8_t* edge_corner_host, const int8_t* edge_axis_host, const int8_t* edge_origin_host) {
    PPS_CHECK_ARG(edge_corner_host && edge_axis_host && edge_origin_host, "pps_mc_set_edges: null pointer");
    return mc_upload_tables(edge_corner_host, edge_axis_host, edge_origin_host);
}

size_t pps_mc_workspace_bytes(int r) { return r >= 2 ? mc_layout(r).total : 0; }

int pps_mc_count(const float* volume, int r, float level, const int8_t* tri_table, int width, void* workspace, size_t workspace_bytes,
                 int64_t* counts_out, void* stream) {
    PPS_CHECK_ARG(volume && tri_table && workspace && counts_out && r >= 2 && r <= 1024 && width >= 3 && width % 3 == 0,
                  "pps_mc_count: bad arguments");
    const McLayout l = mc_layout(r);
    if (workspace_bytes < l.total) {
        set_error("pps_mc_count: workspace %zu < required %zu", workspace_bytes, l.total);
        return PPS_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(workspace);
    const long long cells = (long long)(r - 1) * (r - 1) * (r - 1), edges = 3ll * r * r * r;
    int* tri_count = reinterpret_cast<int*>(base + l.tri_count);
    int* tri_off = reinterpret_cast<int*>(base + l.tri_off);
    int* edge_flag = reinterpret_cast<int*>(base + l.edge_flag);
    int* edge_off = reinterpret_cast<int*>(base + l.edge_off);
    PPS_CUDA(cudaMemsetAsync(tri_count, 0, (cells + 1) * 4, st));
    PPS_CUDA(cudaMemsetAsync(edge_flag, 0, (edges + 1) * 4, st));
    mc_count_kernel<<<(unsigned)ceil_div(cells, 256), 256, 0, st>>>(volume, r, level, tri_table, width, tri_count, edge_flag);
    PPS_LAUNCH_CHECK();
    size_t tb = l.temp_bytes;
    PPS_CUDA(cub::DeviceScan::ExclusiveSum(base + l.temp, tb, tri_count, tri_off, (int)(cells + 1), st));
    tb = l.temp_bytes;
    PPS_CUDA(cub::DeviceScan::ExclusiveSum(base + l.temp, tb, edge_flag, edge_off, (int)(edges + 1), st));
    // the totals are the last entries of the exclusive scans (one trailing zero was appended to both inputs)
    PPS_CUDA(cudaMemsetAsync(counts_out, 0, 16, st));
    PPS_CUDA(cudaMemcpyAsync(counts_out, edge_off + edges, 4, cudaMemcpyDeviceToDevice, st));
    PPS_CUDA(cudaMemcpyAsync(counts_out + 1, tri_off + cells, 4, cudaMemcpyDeviceToDevice, st));
    return PPS_OK;
}

int pps_mc_emit(const float* volume, int r, float level, const int8_t* tri_table, int width, const void* workspace, float* verts_out,
                int32_t* vert_edge_out, int32_t* faces_out, void* stream) {
    PPS_CHECK_ARG(volume && tri_table && workspace && verts_out && vert_edge_out && faces_out && r >= 2, "pps_mc_emit: bad arguments");
    const McLayout l = mc_layout(r);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const char* base = static_cast<const char*>(workspace);
    const long long cells = (long long)(r - 1) * (r - 1) * (r - 1), edges = 3ll * r * r * r;
    const int* tri_off = reinterpret_cast<const int*>(base + l.tri_off);
    const int* edge_flag = reinterpret_cast<const int*>(base + l.edge_flag);
    const int* edge_off = reinterpret_cast<const int*>(base + l.edge_off);
    mc_faces_kernel<<<(unsigned)ceil_div(cells, 256), 256, 0, st>>>(volume, r, level, tri_table, width, tri_off, edge_off, faces_out);
    PPS_LAUNCH_CHECK();
    mc_verts_kernel<<<(unsigned)ceil_div(edges, 256), 256, 0, st>>>(volume, r, level, edge_flag, edge_off, verts_out, vert_edge_out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_refine_init(const float* volume, int r, const float* verts, const int32_t* vert_edge, int64_t nv, float step, float bmin_pad,
                    float* va, float* vb, float* pa, float* pb, float* v, unsigned char* active, void* stream) {
    PPS_CHECK_ARG(volume && verts && vert_edge && va && vb && pa && pb && v && active && nv >= 0, "pps_refine_init: bad arguments");
    if (nv == 0) return PPS_OK;
    refine_init_kernel<<<(unsigned)ceil_div(nv, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(volume, r, verts, vert_edge, nv, step,
                                                                                                 bmin_pad, va, vb, pa, pb, v, active);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_refine_update(const float* pred, int64_t n, float* va, float* vb, float* pa, float* pb, float* v, void* stream) {
    PPS_CHECK_ARG(pred && va && vb && pa && pb && v && n >= 0, "pps_refine_update: bad arguments");
    if (n == 0) return PPS_OK;
    refine_update_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, n, va, vb, pa, pb, v);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
}
