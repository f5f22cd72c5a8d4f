// Exact k-nearest-neighbour search on the GPU.
//
// Replaces the reference's CPU kd-tree (pykdtree, rebuilt on every call: source/poco_utils.py:257-273,
// source/base/proximity.py:40-89).  Index: the points sorted by a 3-D Morton code of a 2^L grid over their bounding
// cube, plus the start offset of every finest cell.  Because Morton order nests, the points of ANY octree node
// (level l, code c) are the contiguous range [start[c << 3(L-l)], start[(c+1) << 3(L-l)]), so one array is a whole
// implicit octree.  A query is one WARP doing a depth-first, near-child-first traversal with box pruning against the
// current k-th distance; the k candidates live in registers as a sorted list spread over the lanes.  Results are exact
// under the total order (dist2, index) with float32 distances computed like pykdtree's float path:
// (dx*dx + dy*dy) + dz*dz, every operation rounded, no FMA.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>

#include "common.cuh"

namespace pps {

constexpr int kMaxLevels = 7;  // 128^3 finest cells, 21-bit codes
constexpr int kDirectMaxShift = 2;  // seeded queries whose search ball spans at most 3 cells of 4 finest cells go straight to those cells
constexpr int kScanMax = 1024;  // an unprunable node with at most this many points is scanned as one contiguous range
constexpr int kRun = 8;        // consecutive queries handled by one warp (each seeds the next one's pruning bound)

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

// finest cells per point (pps_debug_knn_cells): the octree depth is the first with 8^l >= factor * n.  Surface clouds occupy a thin
// shell of the cells: at 16 cells per point (round 1) a leaf of the 100k-point bench cloud holds ~2 points, i.e. 2 of 32 lanes per
// scan step; 2 cells per point measured 5 % faster on the dense grid (tools/knn_run_probe.py), identical results
static int g_knn_cell_factor = 2;
int knn_levels(int64_t n) {
    int l = 2;
    while (l < kMaxLevels && (int64_t(1) << (3 * l)) < n * g_knn_cell_factor) ++l;
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
// One WARP per query.  The k best candidates live in registers as a sorted list distributed over the lanes (element
// g = slot*32 + lane), the traversal stack is per warp in shared memory, and control flow is warp-uniform:
//   * a leaf (<= 32 points, or the finest level) is scanned 32 points at a time with one coalesced float4 load per
//     lane; survivors of the (dist2, index) < worst test are inserted one by one (ballot + shuffle-up);
//   * at an inner node lanes 0..7 fetch the ranges and box distances of the 8 children and push the non-empty,
//     non-pruned ones far-to-near (rank by shuffles), so the nearest child is popped first;
//   * a warp handles kRun consecutive queries: by the triangle inequality sqrt(kth_prev) + |q - q_prev| bounds the k-th
//     distance of the next query, so its traversal prunes from the first node on (grid-ordered queries: tight bound).

__device__ __forceinline__ bool cand_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }

template <int SLOTS>
__global__ void __launch_bounds__(256) knn_warp_kernel(const KnnHeader* __restrict__ hdr, const float4* __restrict__ sorted,
                                                       const int* __restrict__ cell_start, const float* __restrict__ queries,
                                                       long long nq, int k, int32_t* idx_out,
                                                       float* __restrict__ d2_out, int run, int pass, int scan_cap, int scan_child) {
    // per warp: node code, and the node's range in the sorted array (loaded by the parent: no second round trip on the pop)
    __shared__ unsigned int stack_s[8][8 * kMaxLevels + 8];
    __shared__ int stack_lo_s[8][8 * kMaxLevels + 8], stack_hi_s[8][8 * kMaxLevels + 8];
    const unsigned int full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // pass 1: runs of `run` consecutive queries.  A run whose scans have exceeded `scan_cap` points hands its REMAINING queries to
    // pass 2 (marks their first output index with -1) instead of serialising them on this warp: queries deep inside the surface scan
    // most of the cloud (0.3 .. 1 ms each), and 16 of them in a row set the duration of the whole launch.  pass 2: one warp per
    // marked query (unseeded; nothing to amortise for a query that scans everything), all of them in parallel.
    if (pass == 2) run = 1;
    const long long q_first = ((long long)blockIdx.x * 8 + warp) * run;
    if (q_first >= nq) return;
    if (pass == 2 && idx_out[q_first * k] != -1) return;
    int scanned = 0;
    unsigned int* stack = stack_s[warp];
    int* stack_lo = stack_lo_s[warp];
    int* stack_hi = stack_hi_s[warp];
    const int L = hdr->levels, n_pts = hdr->n;
    const float ox = hdr->origin[0], oy = hdr->origin[1], oz = hdr->origin[2];
    const float cell = hdr->cell;
    const float pad = cell * 1e-3f;
    const int wslot = (k - 1) >> 5, wlane = (k - 1) & 31;  // where the k-th best lives
    float bound = INFINITY;  // upper bound of this query's k-th distance, seeded from the previous query of the run
    float pqx = 0.f, pqy = 0.f, pqz = 0.f;
    float ld[SLOTS];  // the k best so far, sorted ascending by (dist2, index): element g = slot*32 + lane
    int li[SLOTS];    // original point index
    int lp[SLOTS];    // position in the Morton-sorted array (to re-evaluate the list for the next query of the run)

    for (int rq = 0; rq < run; ++rq) {
        const long long qi = q_first + rq;
        if (qi >= nq) break;
        if (pass == 1 && rq > 0 && scanned > scan_cap) {
            for (long long t = qi + lane; t < min(nq, q_first + run); t += 32) idx_out[t * k] = -1;
            break;
        }
        const float qx = queries[3 * qi], qy = queries[3 * qi + 1], qz = queries[3 * qi + 2];
        float worst = INFINITY;
        int worst_i = 0x7fffffff;
        // 32 < k <= 256 (the decoder's grid-ordered queries; list sizes 64 / 128 / 256): the previous list seeds this one
        constexpr bool kReseed = SLOTS == 2 || SLOTS == 4 || SLOTS == 8;
        // A query without a predecessor list (the first of a run, every query of the unseeded list sizes) starts from the SLOTS*32
        // points around its own position in the Morton order instead of an empty list: real points near the query, so the k-th of
        // them bounds the search from the first node on, and one sort replaces the one-by-one insertion of everything the first
        // leaves contain (~350 insertions of ~60 instructions for k = 64)
        const bool window = !(kReseed && rq > 0) && n_pts >= SLOTS * 32;
        if (window) {
            const int top = (1 << L) - 1;
            const int cx = min(max(int(floorf((qx - ox) / cell)), 0), top), cy = min(max(int(floorf((qy - oy) / cell)), 0), top),
                      cz = min(max(int(floorf((qz - oz) / cell)), 0), top);
            const int at = cell_start[morton3(cx, cy, cz)];
            const int first = min(max(at - SLOTS * 16, 0), n_pts - SLOTS * 32);
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                lp[s] = first + s * 32 + lane;
                li[s] = __float_as_int(sorted[lp[s]].w);
            }
        }
        const bool seeded = window || (kReseed && rq > 0);
        if (seeded) {
            // Consecutive grid queries share most of their neighbours.  Re-evaluate the previous query's list for this query
            // (its entries are real points, so the k-th of them bounds the k-th distance from the first node on) and sort it
            // by (dist2, index) with a bitonic network over the SLOTS*32 entries.  The traversal then only has to find the few
            // NEW neighbours; points already in the list are recognised by their index when they are scanned again.  (One-by-one
            // insertion of ~350 candidates per query was 58 % of the kernel's instructions.)
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const float4 p = sorted[lp[s]];
                const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
                ld[s] = li[s] == 0x7fffffff ? INFINITY  // an empty entry (fewer points than list entries) stays empty
                                            : __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            }
#pragma unroll
            for (int kk = 2; kk <= SLOTS * 32; kk <<= 1) {
#pragma unroll
                for (int j = kk >> 1; j >= 1; j >>= 1) {
                    if (j >= 32) {  // partner = slot s ^ (j / 32) of the same lane: compare-exchange inside the thread
                        const int ms = j >> 5;
#pragma unroll
                        for (int s = 0; s < SLOTS; ++s) {
                            if ((s & ms) != 0) continue;
                            const int t = s | ms;
                            if (t >= SLOTS) continue;
                            const bool asc = ((s * 32) & kk) == 0;  // bit kk of g is the same for both partners (kk > j)
                            const bool swap = asc ? cand_less(ld[t], li[t], ld[s], li[s]) : cand_less(ld[s], li[s], ld[t], li[t]);
                            if (swap) {
                                const float td = ld[s];
                                const int ti = li[s], tp = lp[s];
                                ld[s] = ld[t];
                                li[s] = li[t];
                                lp[s] = lp[t];
                                ld[t] = td;
                                li[t] = ti;
                                lp[t] = tp;
                            }
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < SLOTS; ++s) {
                            const int g = s * 32 + lane;
                            const float od = __shfl_xor_sync(full, ld[s], j);
                            const int oi = __shfl_xor_sync(full, li[s], j);
                            const int op = __shfl_xor_sync(full, lp[s], j);
                            const bool take_min = ((g & j) == 0) == ((g & kk) == 0);
                            const bool other_less = cand_less(od, oi, ld[s], li[s]);
                            if (take_min == other_less) {
                                ld[s] = od;
                                li[s] = oi;
                                lp[s] = op;
                            }
                        }
                    }
                }
            }
            worst = __shfl_sync(full, ld[wslot < SLOTS ? wslot : SLOTS - 1], wlane);
            worst_i = __shfl_sync(full, li[wslot < SLOTS ? wslot : SLOTS - 1], wlane);
            if (kReseed) bound = INFINITY;  // (the unseeded list sizes keep the previous k-th distance for the triangle inequality below)
        } else {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                ld[s] = INFINITY;
                li[s] = 0x7fffffff;
                lp[s] = 0;
            }
        }
        if (!kReseed && rq > 0) {
            // triangle inequality: the k neighbours of the previous query lie within sqrt(kth_prev) + |q - q_prev| of q
            const float dx = qx - pqx, dy = qy - pqy, dz = qz - pqz;
            const float r = sqrtf(bound) + sqrtf(dx * dx + dy * dy + dz * dz);
            bound = r * r * 1.0001f + 1e-30f;
        }
        auto box_dist = [&](int level, int cx, int cy, int cz) -> float {
            float size = cell * float(1 << (L - level));
            float lx = ox + cx * size - pad, ly = oy + cy * size - pad, lz = oz + cz * size - pad;
            float hx = lx + size + 2 * pad, hy = ly + size + 2 * pad, hz = lz + size + 2 * pad;
            float dx = fmaxf(fmaxf(lx - qx, qx - hx), 0.f);
            float dy = fmaxf(fmaxf(ly - qy, qy - hy), 0.f);
            float dz = fmaxf(fmaxf(lz - qz, qz - hz), 0.f);
            return (dx * dx + dy * dy + dz * dz) * 0.99999f;  // conservative lower bound
        };

        // scan the contiguous range [lo, hi) of the Morton-sorted points, 32 at a time
        auto scan_range = [&](int lo, int hi) {
            scanned += hi - lo;
            float4 pn = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lo + lane < hi) pn = sorted[lo + lane];
            for (int base = lo; base < hi; base += 32) {
                const int i = base + lane;
                const float4 p = pn;
                if (i + 32 < hi) pn = sorted[i + 32];  // the next 32 points are in flight while these are inserted
                float d2 = INFINITY;
                int pi = 0x7fffffff;
                if (i < hi) {
                    const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
                    d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    pi = __float_as_int(p.w);
                }
                unsigned int mask = __ballot_sync(full, i < hi && d2 <= bound && cand_less(d2, pi, worst, worst_i));
                if (seeded && mask) {
                    // points of this step that are already in the list (carried over from the previous query: most of the candidates
                    // that beat the k-th distance): every lane marks the list entries it holds by their position in the sorted
                    // array, one warp-wide OR per step replaces a list lookup per candidate
                    unsigned int mine = 0;
#pragma unroll
                    for (int s = 0; s < SLOTS; ++s) {
                        const unsigned int off = (unsigned int)(lp[s] - base);
                        if (off < 32u && li[s] != 0x7fffffff) mine |= 1u << off;
                    }
                    mask &= ~__reduce_or_sync(full, mine);
                }
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float vd = __shfl_sync(full, d2, src);
                    const int vi = __shfl_sync(full, pi, src);
                    const int vp = base + src;
                    if (!cand_less(vd, vi, worst, worst_i)) continue;  // the bound may have tightened meanwhile
                    int pos = 0;  // number of list elements smaller than the candidate
#pragma unroll
                    for (int s = 0; s < SLOTS; ++s) pos += __popc(__ballot_sync(full, cand_less(ld[s], li[s], vd, vi)));
#pragma unroll
                    for (int s = SLOTS - 1; s >= 0; --s) {
                        float pd = __shfl_up_sync(full, ld[s], 1);
                        int pj = __shfl_up_sync(full, li[s], 1);
                        int pp = __shfl_up_sync(full, lp[s], 1);
                        if (s > 0) {
                            const float cd = __shfl_sync(full, ld[s - 1], 31);
                            const int cj = __shfl_sync(full, li[s - 1], 31);
                            const int cp = __shfl_sync(full, lp[s - 1], 31);
                            if (lane == 0) {
                                pd = cd;
                                pj = cj;
                                pp = cp;
                            }
                        }
                        const int g = s * 32 + lane;
                        if (g == pos) {
                            ld[s] = vd;
                            li[s] = vi;
                            lp[s] = vp;
                        } else if (g > pos) {
                            ld[s] = pd;
                            li[s] = pj;
                            lp[s] = pp;
                        }
                    }
                    worst = __shfl_sync(full, ld[wslot < SLOTS ? wslot : SLOTS - 1], wlane);
                    worst_i = __shfl_sync(full, li[wslot < SLOTS ? wslot : SLOTS - 1], wlane);
                }
            }
        };

        // ---- seeded query: no tree walk.  Every point within sqrt(worst) of q lies in the cells that the ball's bounding box
        // touches; at the finest level where that box spans at most 3 cells per axis those are <= 27 cells = 27 contiguous
        // ranges of the sorted array, tested by 27 lanes at once and scanned one after the other
        bool walked = false;
        if (seeded && worst < INFINITY) {
            const float r = sqrtf(worst) * 1.0001f + pad;
            const int top = (1 << L) - 1;
            int lo_c[3], hi_c[3];
            {
                const float qq[3] = {qx, qy, qz}, oo[3] = {ox, oy, oz};
#pragma unroll
                for (int a = 0; a < 3; ++a) {  // the same cell expression as knn_codes (monotone in the coordinate)
                    lo_c[a] = min(max(int(floorf((qq[a] - r - oo[a]) / cell)), 0), top);
                    hi_c[a] = min(max(int(floorf((qq[a] + r - oo[a]) / cell)), 0), top);
                }
            }
            int shc = 0;
            while (shc < L && (((hi_c[0] >> shc) - (lo_c[0] >> shc)) > 2 || ((hi_c[1] >> shc) - (lo_c[1] >> shc)) > 2 ||
                               ((hi_c[2] >> shc) - (lo_c[2] >> shc)) > 2))
                ++shc;
            const int level = L - shc;
            if (shc <= kDirectMaxShift) {  // larger balls: the tree prunes better than a 27-cell box

            int clo = 0, chi = 0;
            float d = INFINITY;
            if (lane < 27) {
                const int cx = (lo_c[0] >> shc) + lane % 3, cy = (lo_c[1] >> shc) + (lane / 3) % 3, cz = (lo_c[2] >> shc) + lane / 9;
                if (cx <= (hi_c[0] >> shc) && cy <= (hi_c[1] >> shc) && cz <= (hi_c[2] >> shc)) {
                    const unsigned int code = morton3(cx, cy, cz);
                    clo = cell_start[code << (3 * shc)];
                    chi = cell_start[(code + 1u) << (3 * shc)];
                    d = box_dist(level, cx, cy, cz);
                }
            }
            unsigned int m = __ballot_sync(full, chi > clo && d <= worst);
            while (m) {
                const int t = __ffs(m) - 1;
                m &= m - 1;
                const int lo = __shfl_sync(full, clo, t), hi = __shfl_sync(full, chi, t);
                if (__shfl_sync(full, d, t) > worst) continue;  // the k-th distance may have shrunk meanwhile
                scan_range(lo, hi);
            }
            walked = true;
            }
        }

        int sp = walked ? 0 : 1;
        if (lane == 0) {  // root: level 0, cell (0,0,0), all points
            stack[0] = 0u;
            stack_lo[0] = 0;
            stack_hi[0] = hdr->n;
        }
        __syncwarp();
        while (sp > 0) {
            const unsigned int nd = stack[--sp];
            const int level = nd >> 21, cx = nd & 127, cy = (nd >> 7) & 127, cz = (nd >> 14) & 127;
            if (box_dist(level, cx, cy, cz) > fminf(worst, bound)) continue;
            const unsigned int code = morton3(cx, cy, cz);
            const int sh = 3 * (L - level);
            const int lo = stack_lo[sp], hi = stack_hi[sp];
            const int cnt = hi - lo;
            if (cnt == 0) continue;
            bool scan = cnt <= 32 || level == L;
            if (!scan) {
                const int sh2 = sh - 3;
                const unsigned int cbase = code * 8u;
                const float prune = fminf(worst, bound);
                bool ok = false, nonempty = false;
                float d = INFINITY;
                unsigned int child = 0;
                int clo = 0, chi = 0;
                if (lane < 8) {
                    clo = cell_start[(cbase + lane) << sh2];
                    chi = cell_start[(cbase + lane + 1u) << sh2];
                    const int ccx = cx * 2 + (lane & 1), ccy = cy * 2 + ((lane >> 1) & 1), ccz = cz * 2 + (lane >> 2);
                    d = box_dist(level + 1, ccx, ccy, ccz);
                    nonempty = chi > clo;
                    ok = nonempty && d <= prune;
                    child = ((unsigned int)(level + 1) << 21) | ccx | (ccy << 7) | (ccz << 14);
                }
                const unsigned int m = __ballot_sync(full, ok);
                // no child can be pruned and the node is small: its points are one contiguous range, stream through it
                // instead of paying the traversal for every grandchild (queries far from the surface see all points at
                // nearly the same distance and cannot prune)
                // ... or most of them: a child costs ~80 instructions of traversal before its first point is tested (and the leaves of a
                // surface cloud fill 2 of the 32 lanes of a scan step), a 32-point step of a contiguous range ~20
                if (cnt <= kScanMax && (m == __ballot_sync(full, nonempty) || cnt <= scan_child * __popc(m))) {
                    scan = true;
                } else {
                    int rank = 0;  // far children first -> the nearest one ends on top of the stack
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const float dt = __shfl_sync(full, d, t);
                        if (((m >> t) & 1u) && (dt > d || (dt == d && t < lane))) ++rank;
                    }
                    if (ok) {
                        stack[sp + rank] = child;
                        stack_lo[sp + rank] = clo;
                        stack_hi[sp + rank] = chi;
                    }
                    sp += __popc(m);
                    __syncwarp();
                }
            }
            if (scan) scan_range(lo, hi);
        }
        // the list is sorted ascending by (dist2, index): element g = slot*32 + lane
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int g = s * 32 + lane;
            if (g < k) {
                idx_out[qi * k + g] = li[s];
                if (d2_out) d2_out[qi * k + g] = ld[s];
            }
        }
        bound = worst;  // k-th distance of this query (the list is full: k <= n)
        pqx = qx;
        pqy = qy;
        pqz = qz;
    }
}

// points a run may scan before it defers its remaining queries to the second pass: a surface query scans 300..1500 points, a run of 16
// stays below; a query inside the closed surface scans 10^4..10^5
static int g_knn_scan_cap = 65536;  // (16384 until the nodes were scanned whole: one rank's share at 8 / 4 GPUs 5.6 / 10.8 ms at 16384, 4.7 / 7.9 ms at 65536; tools/knn_shard_probe2.py)
static long long g_knn_defer_below = 600000;  // launches of fewer queries use the second pass (see launch_query)
static int g_knn_scan_child = 192;  // pps_debug_knn_scan_child (131^3 grid, k = 64: 42.3 ms at 0, 33.9 at 64, 31.2 at 192, 32.7 at 512; tools/knn_scan_probe.py): a node of at most this many points PER UNPRUNED CHILD is scanned whole
static int g_knn_run = 16;  // pps_debug_knn_run: 16 measured 5 % faster than 8 on the dense grid, 32 and 64 slower (tools/knn_run_probe.py)

template <int SLOTS>
static int launch_query(const KnnHeader* hdr, const float4* sorted, const int* cell_start, const float* queries,
                        int64_t q, int k, int32_t* idx_out, float* d2_out, cudaStream_t st) {
    // seeded lists (k in 33..256: the decoder's grid-ordered queries) profit from longer runs, the others have nothing to amortise
    int run = (SLOTS == 2 || SLOTS == 4 || SLOTS == 8) ? g_knn_run : kRun;
    // small launches (the late sweeps of the region growing, the encoder's 39..2500-point levels): shorten the run until every SM
    // is offered 32 warps; the result does not depend on the run length (unique total order), only the seeding benefit does
    const int64_t want_warps = (int64_t)kNumSMs * 32;
    if (q < want_warps * run) run = (int)std::max<int64_t>(1, q / want_warps);
    // Deferral (pass 2) removes the serial chains that set the duration of a launch whose expensive runs are FEW: one rank's dealt
    // share of the 131^3 grid at 8 / 4 / 2 GPUs takes 7.1 / 11.7 / 22.9 ms instead of 16.6 / 19.5 / 25.6 ms, the late sweeps of the
    // region growing 1-2 ms instead of 10-19 ms.  A deferred query runs unseeded (~1.5x its seeded cost), so a launch that is work-bound
    // anyway -- a contiguous 606 k-query super-chunk of the dense grid, whose middle is ALL expensive queries -- is better off without
    // (whole grid 40 ms without, 45 ms with): deferral is used below that size.
    // (k <= 32 serves the encoder's self-queries -- points of the cloud looking for their neighbours in it, never expensive)
    const bool defer = SLOTS > 1 && run > 1 && q < g_knn_defer_below;
    knn_warp_kernel<SLOTS><<<(unsigned)ceil_div(q, 8 * run), 256, 0, st>>>(hdr, sorted, cell_start, queries, q, k, idx_out, d2_out, run, 1,
                                                                          defer ? g_knn_scan_cap : 0x7fffffff, g_knn_scan_child);
    PPS_LAUNCH_CHECK();
    if (defer) {  // the deferred queries (none on surface-hugging query sets: the blocks find no mark and exit)
        knn_warp_kernel<SLOTS><<<(unsigned)ceil_div(q, 8), 256, 0, st>>>(hdr, sorted, cell_start, queries, q, k, idx_out, d2_out, 1, 2, 0, g_knn_scan_child);
    }
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
    if (k <= 32) return launch_query<1>(hdr, sorted, cell_start, queries, q, k, idx_out, d2_out, st);
    if (k <= 64) return launch_query<2>(hdr, sorted, cell_start, queries, q, k, idx_out, d2_out, st);
    if (k <= 128) return launch_query<4>(hdr, sorted, cell_start, queries, q, k, idx_out, d2_out, st);
    if (k <= 256) return launch_query<8>(hdr, sorted, cell_start, queries, q, k, idx_out, d2_out, st);
    return launch_query<16>(hdr, sorted, cell_start, queries, q, k, idx_out, d2_out, st);
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

int knn_set_cell_factor(int f) {
    const int old = g_knn_cell_factor;
    if (f >= 1 && f <= 64) g_knn_cell_factor = f;
    return old;
}

int knn_set_scan_child(int v) {
    const int old = g_knn_scan_child;
    if (v >= 0 && v <= kScanMax) g_knn_scan_child = v;
    return old;
}

int knn_set_scan_cap(int v) {
    const int old = g_knn_scan_cap;
    if (v >= 1) g_knn_scan_cap = v;
    return old;
}

int knn_set_run(int run) {
    const int old = g_knn_run;
    if (run >= 1 && run <= 256) g_knn_run = run;
    return old;
}

}  // namespace pps

extern "C" {
// queries per warp run of the seeded search (k in 33..256); returns the previous value, run < 1 only reads
int pps_debug_knn_run(int run) { return pps::knn_set_run(run); }
// finest octree cells per point (indices built afterwards use it; an index must be queried under the setting it was built with)
int pps_debug_knn_cells(int factor) { return pps::knn_set_cell_factor(factor); }
// points per unpruned child up to which an inner node is scanned as one range instead of being traversed; returns the previous value
int pps_debug_knn_scan_child(int points) { return pps::knn_set_scan_child(points); }
// points a run may scan before it hands its remaining queries to the second pass (launches below 600 k queries); returns the previous value
int pps_debug_knn_scan_cap(int points) { return pps::knn_set_scan_cap(points); }
size_t pps_knn_index_bytes(int64_t n) { return n > 0 ? pps::knn_layout(n).total : 0; }
int pps_knn_build(const float* pts, int64_t n, void* index, size_t index_bytes, void* stream) {
    return pps::knn_build_impl(pts, n, index, index_bytes, static_cast<cudaStream_t>(stream));
}
int pps_knn_query(const void* index, int64_t n, const float* queries, int64_t q, int k, int32_t* idx_out,
                  float* dist2_out, void* stream) {
    return pps::knn_query_impl(index, n, queries, q, k, idx_out, dist2_out, static_cast<cudaStream_t>(stream));
}
}
