// Exact k-nearest-neighbour search on the GPU.
//
// Replaces the reference's CPU kd-tree (pykdtree, rebuilt on every call: source/poco_utils.py:257-273,
// source/base/proximity.py:40-89).  Index: the points sorted by a 3-D Morton code of a 2^L grid over their bounding
// cube, plus the start offset of every finest cell.  Because Morton order nests, the points of ANY octree node
// (level l, code c) are the contiguous range [start[c << 3(L-l)], start[(c+1) << 3(L-l)]), so one array is a whole
// implicit octree.  A query is one thread doing a depth-first, near-child-first traversal with box pruning against
// the current k-th distance; the k candidates live in a per-thread max-heap in shared memory ([slot][thread]
// layout, conflict free).  Results are exact under the total order (dist2, index) with float32 distances computed
// like pykdtree's float path: (dx*dx + dy*dy) + dz*dz, every operation rounded, no FMA.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pps {

constexpr int kMaxLevels = 7;  // 128^3 finest cells, 21-bit codes
constexpr int kLeafCount = 16;

struct KnnHeader {
    unsigned int min_enc[3];
    unsigned int max_enc[3];
    float origin[3];
    float cell;  // finest cell edge
    int levels;
    int n;
    int pad[4];
};
static_assert(sizeof(KnnHeader) == 64, "header is 64 bytes");

int knn_levels(int64_t n) {
    int l = 2;
    while (l < kMaxLevels && (int64_t(1) << (3 * l)) < n * 16) ++l;
    return l;
}

struct KnnLayout {
    size_t header, sorted, cell_start, keys0, keys1, vals0, vals1, temp, temp_bytes, total;
};

static KnnLayout knn_layout(int64_t n) {
    KnnLayout l;
    int lv = knn_levels(n);
    size_t cells = (size_t(1) << (3 * lv)) + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        off = align_up(off, 256);
        size_t r = off;
        off += bytes;
        return r;
    };
    l.header = take(sizeof(KnnHeader));
    l.sorted = take(size_t(n) * sizeof(float4));
    l.cell_start = take(cells * sizeof(int));
    l.keys0 = take(size_t(n) * 4);
    l.keys1 = take(size_t(n) * 4);
    l.vals0 = take(size_t(n) * 4);
    l.vals1 = take(size_t(n) * 4);
    l.temp_bytes = size_t(n) * 8 + (size_t(4) << 20);
    l.temp = take(l.temp_bytes);
    l.total = align_up(off, 256);
    return l;
}

__device__ __forceinline__ unsigned int enc_float(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_float(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ unsigned int spread3(unsigned int v) {
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ unsigned int morton3(unsigned int x, unsigned int y, unsigned int z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}

__global__ void knn_init_header(KnnHeader* h, int levels, int n) {
    for (int a = 0; a < 3; ++a) {
        h->min_enc[a] = 0xffffffffu;
        h->max_enc[a] = 0u;
    }
    h->levels = levels;
    h->n = n;
}

__global__ void knn_bbox(const float* __restrict__ pts, int n, KnnHeader* h) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = pts[3 * (size_t)i + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&h->min_enc[a], enc_float(mn[a]));
            atomicMax(&h->max_enc[a], enc_float(mx[a]));
        }
    }
}

__global__ void knn_codes(const float* __restrict__ pts, int n, KnnHeader* h, unsigned int* keys, unsigned int* vals) {
    int levels = h->levels;
    float ox = dec_float(h->min_enc[0]), oy = dec_float(h->min_enc[1]), oz = dec_float(h->min_enc[2]);
    float ext = fmaxf(fmaxf(dec_float(h->max_enc[0]) - ox, dec_float(h->max_enc[1]) - oy), dec_float(h->max_enc[2]) - oz);
    float cell = fmaxf(ext, 1e-20f) * 1.00001f / float(1 << levels);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        h->origin[0] = ox;
        h->origin[1] = oy;
        h->origin[2] = oz;
        h->cell = cell;
    }
    if (i >= n) return;
    int hi = (1 << levels) - 1;
    int cx = min(max(int(floorf((pts[3 * (size_t)i + 0] - ox) / cell)), 0), hi);
    int cy = min(max(int(floorf((pts[3 * (size_t)i + 1] - oy) / cell)), 0), hi);
    int cz = min(max(int(floorf((pts[3 * (size_t)i + 2] - oz) / cell)), 0), hi);
    keys[i] = morton3(cx, cy, cz);
    vals[i] = i;
}

// sorted[i] = (xyz of the i-th point in Morton order, original index); cell_start[c] = first i with key >= c
__global__ void knn_finalize(const float* __restrict__ pts, int n, const unsigned int* __restrict__ keys,
                             const unsigned int* __restrict__ vals, float4* sorted, int* cell_start, int num_cells) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned int src = vals[i];
    sorted[i] = make_float4(pts[3 * (size_t)src], pts[3 * (size_t)src + 1], pts[3 * (size_t)src + 2], __int_as_float(int(src)));
    unsigned int key = keys[i];
    unsigned int prev_plus = (i == 0) ? 0u : keys[i - 1] + 1u;
    for (unsigned int c = prev_plus; c <= key; ++c) cell_start[c] = i;  // total work over all threads = #cells
    if (i == n - 1)
        for (unsigned int c = key + 1; c <= (unsigned int)num_cells; ++c) cell_start[c] = n;
}

// ---- query -------------------------------------------------------------------------------------------------------

struct HeapRef {
    float* d;
    int* id;
    int stride;  // BS + 1
    int k;
    __device__ __forceinline__ float& dist(int s) { return d[s * stride]; }
    __device__ __forceinline__ int& idx(int s) { return id[s * stride]; }
};

__device__ __forceinline__ bool heap_greater(float da, int ia, float db, int ib) {
    return da > db || (da == db && ia > ib);
}

// place (nd, ni) at the root of a max-heap of size n (root is being replaced) and restore the heap property
__device__ __forceinline__ void heap_sift_root(HeapRef& h, int n, float nd, int ni) {
    int pos = 0;
    while (true) {
        int l = 2 * pos + 1;
        if (l >= n) break;
        int r = l + 1;
        float dl = h.dist(l);
        int il = h.idx(l);
        int c = l;
        float dc = dl;
        int ic = il;
        if (r < n) {
            float dr = h.dist(r);
            int ir = h.idx(r);
            if (heap_greater(dr, ir, dl, il)) {
                c = r;
                dc = dr;
                ic = ir;
            }
        }
        if (!heap_greater(dc, ic, nd, ni)) break;
        h.dist(pos) = dc;
        h.idx(pos) = ic;
        pos = c;
    }
    h.dist(pos) = nd;
    h.idx(pos) = ni;
}

template <int BS>
__global__ void __launch_bounds__(BS) knn_query_kernel(const KnnHeader* __restrict__ hdr, const float4* __restrict__ sorted,
                                                       const int* __restrict__ cell_start, const float* __restrict__ queries,
                                                       long long nq, int k, int32_t* __restrict__ idx_out,
                                                       float* __restrict__ d2_out) {
    extern __shared__ float smem[];
    const int stride = BS + 1;
    float* hd = smem;
    int* hi = reinterpret_cast<int*>(smem + (size_t)k * stride);
    const int tid = threadIdx.x;
    const long long q0 = (long long)blockIdx.x * BS;
    const long long qi = q0 + tid;
    HeapRef heap{hd + tid, hi + tid, stride, k};
    for (int s = 0; s < k; ++s) {
        heap.dist(s) = INFINITY;
        heap.idx(s) = 0x7fffffff;
    }
    if (qi < nq) {
        const int L = hdr->levels;
        const float ox = hdr->origin[0], oy = hdr->origin[1], oz = hdr->origin[2];
        const float cell = hdr->cell;
        const float pad = cell * 1e-3f;
        const float qx = queries[3 * qi], qy = queries[3 * qi + 1], qz = queries[3 * qi + 2];
        float worst = INFINITY;
        int worst_i = 0x7fffffff;

        auto box_dist = [&](int level, int cx, int cy, int cz) -> float {
            float size = cell * float(1 << (L - level));
            float lx = ox + cx * size - pad, ly = oy + cy * size - pad, lz = oz + cz * size - pad;
            float hx = lx + size + 2 * pad, hy = ly + size + 2 * pad, hz = lz + size + 2 * pad;
            float dx = fmaxf(fmaxf(lx - qx, qx - hx), 0.f);
            float dy = fmaxf(fmaxf(ly - qy, qy - hy), 0.f);
            float dz = fmaxf(fmaxf(lz - qz, qz - hz), 0.f);
            return (dx * dx + dy * dy + dz * dz) * 0.99999f;  // conservative lower bound
        };

        unsigned int stack[8 * kMaxLevels + 8];
        int sp = 0;
        stack[sp++] = 0u;  // root: level 0, cell (0,0,0)
        const unsigned char order[8] = {0, 1, 2, 4, 3, 5, 6, 7};
        while (sp > 0) {
            unsigned int nd = stack[--sp];
            int level = nd >> 21, cx = nd & 127, cy = (nd >> 7) & 127, cz = (nd >> 14) & 127;
            if (box_dist(level, cx, cy, cz) > worst) continue;
            unsigned int code = morton3(cx, cy, cz);
            int sh = 3 * (L - level);
            int lo = cell_start[code << sh], hi_ = cell_start[(code + 1u) << sh];
            int cnt = hi_ - lo;
            if (cnt == 0) continue;
            if (cnt <= kLeafCount || level == L) {
                for (int i = lo; i < hi_; ++i) {
                    float4 p = sorted[i];
                    float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
                    float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    int pi = __float_as_int(p.w);
                    if (heap_greater(worst, worst_i, d2, pi)) {
                        heap_sift_root(heap, k, d2, pi);
                        worst = heap.dist(0);
                        worst_i = heap.idx(0);
                    }
                }
            } else {
                float size = cell * float(1 << (L - level));
                int o = (qx >= ox + (cx + 0.5f) * size ? 1 : 0) | (qy >= oy + (cy + 0.5f) * size ? 2 : 0) |
                        (qz >= oz + (cz + 0.5f) * size ? 4 : 0);
                int sh2 = sh - 3;
                unsigned int base = code * 8u;
                for (int t = 7; t >= 0; --t) {  // far children first so that the nearest is popped first
                    int ch = o ^ order[t];
                    int clo = cell_start[(base + ch) << sh2], chi = cell_start[(base + ch + 1u) << sh2];
                    if (chi == clo) continue;
                    int ccx = cx * 2 + (ch & 1), ccy = cy * 2 + ((ch >> 1) & 1), ccz = cz * 2 + (ch >> 2);
                    if (box_dist(level + 1, ccx, ccy, ccz) > worst) continue;
                    stack[sp++] = ((unsigned int)(level + 1) << 21) | ccx | (ccy << 7) | (ccz << 14);
                }
            }
        }
        // in-place heapsort -> ascending (dist2, index)
        for (int end = k - 1; end > 0; --end) {
            float td = heap.dist(end);
            int ti = heap.idx(end);
            heap.dist(end) = heap.dist(0);
            heap.idx(end) = heap.idx(0);
            heap_sift_root(heap, end, td, ti);
        }
    }
    __syncthreads();
    // transposed, coalesced write-out of the block's [BS,k] results
    long long rows = nq - q0 < BS ? nq - q0 : BS;
    long long total = rows * k;
    for (long long e = tid; e < total; e += BS) {
        int r = int(e / k), s = int(e % k);
        idx_out[q0 * k + e] = hi[s * stride + r];
        if (d2_out) d2_out[q0 * k + e] = hd[s * stride + r];
    }
}

template <int BS>
static int launch_query(const KnnHeader* hdr, const float4* sorted, const int* cell_start, const float* queries,
                        int64_t q, int k, int32_t* idx_out, float* d2_out, cudaStream_t st) {
    size_t smem = size_t(k) * (BS + 1) * 8;
    PPS_CUDA(cudaFuncSetAttribute(knn_query_kernel<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_query_kernel<BS><<<(unsigned)ceil_div(q, BS), BS, smem, st>>>(hdr, sorted, cell_start, queries, q, k, idx_out, d2_out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int knn_query_impl(const void* index, int64_t n, const float* queries, int64_t q, int k, int32_t* idx_out,
                   float* d2_out, cudaStream_t st) {
    if (q == 0) return PPS_OK;
    PPS_CHECK_ARG(index && queries && idx_out, "pps_knn_query: null pointer");
    PPS_CHECK_ARG(n > 0 && n < (int64_t(1) << 31), "pps_knn_query: n=%lld out of range", (long long)n);
    PPS_CHECK_ARG(k >= 1 && k <= n && k <= 512, "pps_knn_query: k=%d must be in [1, min(n,512)] (n=%lld)", k, (long long)n);
    KnnLayout l = knn_layout(n);
    const char* base = static_cast<const char*>(index);
    const KnnHeader* hdr = reinterpret_cast<const KnnHeader*>(base + l.header);
    const float4* sorted = reinterpret_cast<const float4*>(base + l.sorted);
    const int* cell_start = reinterpret_cast<const int*>(base + l.cell_start);
    if (k <= 64) return launch_query<128>(hdr, sorted, cell_start, queries, q, k, idx_out, d2_out, st);
    if (k <= 200) return launch_query<64>(hdr, sorted, cell_start, queries, q, k, idx_out, d2_out, st);
    return launch_query<32>(hdr, sorted, cell_start, queries, q, k, idx_out, d2_out, st);
}

int knn_build_impl(const float* pts, int64_t n, void* index, size_t index_bytes, cudaStream_t st) {
    PPS_CHECK_ARG(pts && index, "pps_knn_build: null pointer");
    PPS_CHECK_ARG(n > 0 && n < (int64_t(1) << 31), "pps_knn_build: n=%lld out of range", (long long)n);
    KnnLayout l = knn_layout(n);
    if (index_bytes < l.total) {
        set_error("pps_knn_build: index buffer %zu < required %zu", index_bytes, l.total);
        return PPS_ERR_WORKSPACE;
    }
    char* base = static_cast<char*>(index);
    KnnHeader* hdr = reinterpret_cast<KnnHeader*>(base + l.header);
    float4* sorted = reinterpret_cast<float4*>(base + l.sorted);
    int* cell_start = reinterpret_cast<int*>(base + l.cell_start);
    unsigned int* keys0 = reinterpret_cast<unsigned int*>(base + l.keys0);
    unsigned int* keys1 = reinterpret_cast<unsigned int*>(base + l.keys1);
    unsigned int* vals0 = reinterpret_cast<unsigned int*>(base + l.vals0);
    unsigned int* vals1 = reinterpret_cast<unsigned int*>(base + l.vals1);
    int levels = knn_levels(n);
    int num_cells = 1 << (3 * levels);
    int in = int(n);
    int blocks = (int)ceil_div(n, 256);
    knn_init_header<<<1, 1, 0, st>>>(hdr, levels, in);
    PPS_LAUNCH_CHECK();
    knn_bbox<<<min(blocks, 4 * kNumSMs), 256, 0, st>>>(pts, in, hdr);
    PPS_LAUNCH_CHECK();
    knn_codes<<<blocks, 256, 0, st>>>(pts, in, hdr, keys0, vals0);
    PPS_LAUNCH_CHECK();
    cub::DoubleBuffer<unsigned int> dk(keys0, keys1), dv(vals0, vals1);
    size_t need = 0;
    PPS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, in, 0, 3 * levels, st));
    if (need > l.temp_bytes) {
        set_error("pps_knn_build: radix sort needs %zu temp bytes, reserved %zu", need, l.temp_bytes);
        return PPS_ERR_WORKSPACE;
    }
    PPS_CUDA(cub::DeviceRadixSort::SortPairs(base + l.temp, need, dk, dv, in, 0, 3 * levels, st));
    knn_finalize<<<blocks, 256, 0, st>>>(pts, in, dk.Current(), dv.Current(), sorted, cell_start, num_cells);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

}  // namespace pps

extern "C" {
size_t pps_knn_index_bytes(int64_t n) { return n > 0 ? pps::knn_layout(n).total : 0; }
int pps_knn_build(const float* pts, int64_t n, void* index, size_t index_bytes, void* stream) {
    return pps::knn_build_impl(pts, n, index, index_bytes, static_cast<cudaStream_t>(stream));
}
int pps_knn_query(const void* index, int64_t n, const float* queries, int64_t q, int k, int32_t* idx_out,
                  float* dist2_out, void* stream) {
    return pps::knn_query_impl(index, n, queries, q, k, idx_out, dist2_out, static_cast<cudaStream_t>(stream));
}
}
