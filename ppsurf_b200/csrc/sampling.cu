// Quantised support sampling and the encoder's index tensors on the device.
//
// Replaces sampling_quantized (source/poco_data_loader.py:59-134: random rotation -> voxel grid -> one representative per
// voxel -> halve the voxel until enough points -> random trim) and get_fkaconv_ids (poco_data_loader.py:137-209: four
// samplings at ratio 1/4 and 13 kNN index tensors).  The reference runs the voxel grid through torch_geometric with a
// host loop; here every round is a fixed sequence of kernels steered by device-side state (no host synchronisation):
//   bbox/min of the rotated alive points -> insert into an open-addressing hash table keyed by the voxel (the lowest
//   point index of a voxel wins: deterministic) -> mark representatives, give them a hashed random key -> radix sort ->
//   accept all of them (and halve the voxel) or, in the last round, the first n_support - picked of the shuffled list.
// Rounds after completion are no-ops.  The reference's sampling is random by construction; parity is distributional.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pps {

int knn_build_impl(const float* pts, int64_t n, void* index, size_t index_bytes, cudaStream_t st);
int knn_query_impl(const void* index, int64_t n, const float* queries, int64_t q, int k, int32_t* idx_out, float* d2_out,
                   cudaStream_t st);

constexpr int kSampleRounds = 6;
constexpr int kIdStreams = 8;  // clouds of a batch are independent chains of small kernels: run them on side streams
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

struct SampleState {
    float vox;
    int picked;
    int nreps;
    int done;
    unsigned int mn[3];  // ordered-int encoding of the minimum of the rotated alive points
    unsigned int bb_min[3], bb_max[3];
    int n_support;
};

__device__ __forceinline__ unsigned int s_enc(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float s_dec(unsigned int u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
__device__ __forceinline__ unsigned int hash32(unsigned int x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

__global__ void sample_init(SampleState* st, unsigned char* alive, int n, int n_support) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) alive[i] = 1;
    if (i == 0) {
        st->picked = 0;
        st->nreps = 0;
        st->done = 0;
        st->n_support = n_support;
        for (int a = 0; a < 3; ++a) {
            st->mn[a] = 0xffffffffu;
            st->bb_min[a] = 0xffffffffu;
            st->bb_max[a] = 0u;
        }
    }
}

__global__ void sample_bbox(const float* __restrict__ pts, int n, SampleState* st) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = pts[3 * (size_t)i + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&st->bb_min[a], s_enc(mn[a]));
            atomicMax(&st->bb_max[a], s_enc(mx[a]));
        }
}

// voxel edge = ||bbox||_2 / sqrt(n_support)  (poco_data_loader.py:86-88)
__global__ void sample_vox(SampleState* st) {
    float d2 = 0.f;
    for (int a = 0; a < 3; ++a) {
        float e = s_dec(st->bb_max[a]) - s_dec(st->bb_min[a]);
        d2 += e * e;
    }
    st->vox = fmaxf(sqrtf(d2), 1e-20f) / sqrtf(float(st->n_support));
}

__device__ __forceinline__ void rotate(const float* __restrict__ r, const float* __restrict__ p, float& x, float& y, float& z) {
    x = r[0] * p[0] + r[1] * p[1] + r[2] * p[2];
    y = r[3] * p[0] + r[4] * p[1] + r[5] * p[2];
    z = r[6] * p[0] + r[7] * p[1] + r[8] * p[2];
}

__global__ void sample_min(const float* __restrict__ pts, const unsigned char* __restrict__ alive, int n, const float* __restrict__ rot,
                           SampleState* st) {
    if (st->done) return;
    float mn[3] = {INFINITY, INFINITY, INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (!alive[i]) continue;
        float x, y, z;
        rotate(rot, pts + 3 * (size_t)i, x, y, z);
        mn[0] = fminf(mn[0], x);
        mn[1] = fminf(mn[1], y);
        mn[2] = fminf(mn[2], z);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int a = 0; a < 3; ++a) atomicMin(&st->mn[a], s_enc(mn[a]));
}

__device__ __forceinline__ unsigned long long voxel_key(const float* __restrict__ rot, const float* __restrict__ p, const SampleState* st) {
    float x, y, z;
    rotate(rot, p, x, y, z);
    const float inv = 1.f / st->vox;
    unsigned long long cx = (unsigned long long)fminf(fmaxf(floorf((x - s_dec(st->mn[0])) * inv), 0.f), 2097151.f);
    unsigned long long cy = (unsigned long long)fminf(fmaxf(floorf((y - s_dec(st->mn[1])) * inv), 0.f), 2097151.f);
    unsigned long long cz = (unsigned long long)fminf(fmaxf(floorf((z - s_dec(st->mn[2])) * inv), 0.f), 2097151.f);
    return cx | (cy << 21) | (cz << 42);
}
__device__ __forceinline__ unsigned int key_slot(unsigned long long key, unsigned int mask) {
    return hash32((unsigned int)key ^ hash32((unsigned int)(key >> 32))) & mask;
}

__global__ void sample_insert(const float* __restrict__ pts, const unsigned char* __restrict__ alive, int n, const float* __restrict__ rot,
                              const SampleState* st, unsigned long long* keys, int* vals, unsigned int mask) {
    if (st->done) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !alive[i]) return;
    const unsigned long long key = voxel_key(rot, pts + 3 * (size_t)i, st);
    unsigned int h = key_slot(key, mask);
    while (true) {
        unsigned long long prev = atomicCAS(&keys[h], kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) {
            atomicMin(&vals[h], i);  // the lowest index of a voxel represents it
            return;
        }
        h = (h + 1) & mask;
    }
}

// representatives get a hashed random sort key (31 bits), everything else sorts to the end
__global__ void sample_mark(const float* __restrict__ pts, const unsigned char* __restrict__ alive, int n, const float* __restrict__ rot,
                            SampleState* st, const unsigned long long* __restrict__ keys, const int* __restrict__ vals, unsigned int mask,
                            unsigned int seed, unsigned int* sort_key, int* sort_val) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool rep = false;
    // the random rotation of this round is mixed into the seed: a CUDA-graph replay (seed baked into the launch) with fresh
    // rotations still draws a fresh random order
    seed ^= hash32(__float_as_uint(rot[1]) ^ hash32(__float_as_uint(rot[5])));
    if (i < n && !st->done && alive[i]) {
        const unsigned long long key = voxel_key(rot, pts + 3 * (size_t)i, st);
        unsigned int h = key_slot(key, mask);
        while (keys[h] != key) h = (h + 1) & mask;
        rep = vals[h] == i;
    }
    if (i < n) {
        sort_val[i] = i;
        sort_key[i] = rep ? (hash32(seed ^ hash32((unsigned int)i)) >> 1) : 0xFFFFFFFFu;
    }
    const unsigned int m = __ballot_sync(0xffffffffu, rep);  // whole warp: no early exit above
    if (rep && (threadIdx.x & 31) == (__ffs(m) - 1)) atomicAdd(&st->nreps, __popc(m));
}

// after the sort: the first nreps entries are the representatives in random order
__global__ void sample_apply(int n, const SampleState* st, const int* __restrict__ sorted_val, unsigned char* alive, int32_t* sel_out) {
    if (st->done) return;
    const bool all = st->picked + st->nreps < st->n_support;
    const int take = all ? st->nreps : st->n_support - st->picked;
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= take) return;
    const int i = sorted_val[e];
    sel_out[st->picked + e] = i;
    alive[i] = 0;
}

__global__ void sample_update(SampleState* st) {
    if (st->done) return;
    const bool all = st->picked + st->nreps < st->n_support;
    if (all) {
        st->picked += st->nreps;
        st->vox *= 0.5f;  // poco_data_loader.py:118
    } else {
        st->picked = st->n_support;
        st->done = 1;
    }
    st->nreps = 0;
    for (int a = 0; a < 3; ++a) st->mn[a] = 0xffffffffu;
}

// safety net (duplicate-heavy clouds): after the last round fill up with the remaining points in index order
__global__ void sample_finish(int n, SampleState* st, unsigned char* alive, int32_t* sel_out) {
    if (st->done) return;
    int p = st->picked;
    for (int i = 0; i < n && p < st->n_support; ++i)
        if (alive[i]) {
            sel_out[p++] = i;
            alive[i] = 0;
        }
    st->picked = p;
    st->done = 1;
}

__global__ void gather_points(const float* __restrict__ pts, const int32_t* __restrict__ sel, int m, float* __restrict__ out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * m) return;
    out[e] = pts[3 * (size_t)sel[e / 3] + e % 3];
}

struct SampleLayout {
    size_t state, alive, keys, vals, sk0, sk1, sv0, sv1, temp, temp_bytes, total;
    unsigned int mask;
};
static SampleLayout sample_layout(int64_t n) {
    SampleLayout l;
    size_t h = 1024;
    while (h < size_t(2 * n)) h <<= 1;
    l.mask = (unsigned int)(h - 1);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        off = align_up(off, 256);
        size_t r = off;
        off += bytes;
        return r;
    };
    l.state = take(sizeof(SampleState));
    l.alive = take(size_t(n));
    l.keys = take(h * 8);
    l.vals = take(h * 4);
    l.sk0 = take(size_t(n) * 4);
    l.sk1 = take(size_t(n) * 4);
    l.sv0 = take(size_t(n) * 4);
    l.sv1 = take(size_t(n) * 4);
    l.temp_bytes = size_t(n) * 8 + (size_t(4) << 20);
    l.temp = take(l.temp_bytes);
    l.total = align_up(off, 256);
    return l;
}

int sample_quantized_impl(const float* pts, int64_t n, int64_t n_support, const float* rotations, int n_rot, uint32_t seed, void* ws,
                          size_t ws_bytes, int32_t* sel_out, cudaStream_t st) {
    PPS_CHECK_ARG(pts && rotations && ws && sel_out, "pps_sample_quantized: null pointer");
    PPS_CHECK_ARG(n > 0 && n < (int64_t(1) << 30) && n_support > 0 && n_support <= n && n_rot >= 1,
                  "pps_sample_quantized: n=%lld n_support=%lld n_rot=%d", (long long)n, (long long)n_support, n_rot);
    SampleLayout l = sample_layout(n);
    if (ws_bytes < l.total) {
        set_error("pps_sample_quantized: workspace %zu < required %zu", ws_bytes, l.total);
        return PPS_ERR_WORKSPACE;
    }
    char* base = static_cast<char*>(ws);
    SampleState* state = reinterpret_cast<SampleState*>(base + l.state);
    unsigned char* alive = reinterpret_cast<unsigned char*>(base + l.alive);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(base + l.keys);
    int* vals = reinterpret_cast<int*>(base + l.vals);
    unsigned int* sk0 = reinterpret_cast<unsigned int*>(base + l.sk0);
    unsigned int* sk1 = reinterpret_cast<unsigned int*>(base + l.sk1);
    int* sv0 = reinterpret_cast<int*>(base + l.sv0);
    int* sv1 = reinterpret_cast<int*>(base + l.sv1);
    const int in = int(n);
    const int blocks = (int)ceil_div(n, 256);
    const int rblocks = blocks < 2 * kNumSMs ? blocks : 2 * kNumSMs;
    sample_init<<<blocks, 256, 0, st>>>(state, alive, in, (int)n_support);
    PPS_LAUNCH_CHECK();
    if (n_support == n) {  // the reference returns all points unchanged (poco_data_loader.py:79-82)
        PPS_CUDA(cudaMemsetAsync(alive, 1, size_t(n), st));
        sample_finish<<<1, 1, 0, st>>>(in, state, alive, sel_out);
        PPS_LAUNCH_CHECK();
        return PPS_OK;
    }
    sample_bbox<<<rblocks, 256, 0, st>>>(pts, in, state);
    PPS_LAUNCH_CHECK();
    sample_vox<<<1, 1, 0, st>>>(state);
    PPS_LAUNCH_CHECK();
    const int rounds = n_rot < kSampleRounds ? n_rot : kSampleRounds;
    for (int t = 0; t < rounds; ++t) {
        const float* rot = rotations + 9 * t;
        sample_min<<<rblocks, 256, 0, st>>>(pts, alive, in, rot, state);
        PPS_LAUNCH_CHECK();
        PPS_CUDA(cudaMemsetAsync(keys, 0xFF, (size_t(l.mask) + 1) * 8, st));
        PPS_CUDA(cudaMemsetAsync(vals, 0x7F, (size_t(l.mask) + 1) * 4, st));
        sample_insert<<<blocks, 256, 0, st>>>(pts, alive, in, rot, state, keys, vals, l.mask);
        PPS_LAUNCH_CHECK();
        sample_mark<<<blocks, 256, 0, st>>>(pts, alive, in, rot, state, keys, vals, l.mask, seed * 0x9E3779B9u + (unsigned)t, sk0, sv0);
        PPS_LAUNCH_CHECK();
        cub::DoubleBuffer<unsigned int> dk(sk0, sk1);
        cub::DoubleBuffer<int> dv(sv0, sv1);
        size_t need = 0;
        PPS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, in, 0, 32, st));
        if (need > l.temp_bytes) {
            set_error("pps_sample_quantized: radix sort needs %zu temp bytes, reserved %zu", need, l.temp_bytes);
            return PPS_ERR_WORKSPACE;
        }
        PPS_CUDA(cub::DeviceRadixSort::SortPairs(base + l.temp, need, dk, dv, in, 0, 32, st));
        sample_apply<<<blocks, 256, 0, st>>>(in, state, dv.Current(), alive, sel_out);
        PPS_LAUNCH_CHECK();
        sample_update<<<1, 1, 0, st>>>(state);
        PPS_LAUNCH_CHECK();
    }
    sample_finish<<<1, 1, 0, st>>>(in, state, alive, sel_out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

}  // namespace pps

using namespace pps;

extern "C" {

size_t pps_sample_workspace_bytes(int64_t n) { return n > 0 ? sample_layout(n).total : 0; }

int pps_sample_quantized(const float* pts, int64_t n, int64_t n_support, const float* rotations, int n_rot, uint32_t seed,
                         void* workspace, size_t workspace_bytes, int32_t* sel_out, void* stream) {
    return sample_quantized_impl(pts, n, n_support, rotations, n_rot, seed, workspace, workspace_bytes, sel_out,
                                 static_cast<cudaStream_t>(stream));
}

// level sizes of the encoder: n_{l+1} = max(1, int(n_l * 0.25))  (poco_data_loader.py:75,148-151)
static void level_sizes(int64_t n0, int64_t (&n)[5]) {
    n[0] = n0;
    for (int l = 1; l < 5; ++l) n[l] = n[l - 1] / 4 > 1 ? n[l - 1] / 4 : 1;
}

static size_t encoder_ids_slice_bytes(int64_t n0) {
    int64_t n[5];
    level_sizes(n0, n);
    size_t bytes = align_up(pps_sample_workspace_bytes(n0), 256);
    for (int l = 0; l < 5; ++l) bytes += align_up(pps_knn_index_bytes(n[l]), 256);
    bytes += align_up(size_t(n[1]) * 4, 256);  // selection scratch
    return bytes;
}

size_t pps_encoder_ids_workspace_bytes(int64_t n0) { return n0 > 0 ? kIdStreams * encoder_ids_slice_bytes(n0) : 0; }

int pps_encoder_ids(const float* pts, int64_t b, int64_t n0, const float* rotations, int n_rot, uint32_t seed, void* workspace,
                    size_t workspace_bytes, const pps_encoder_ids_out* out, void* stream) {
    PPS_CHECK_ARG(pts && rotations && workspace && out, "pps_encoder_ids: null pointer");
    PPS_CHECK_ARG(b >= 1 && n0 >= 1, "pps_encoder_ids: bad sizes");
    if (workspace_bytes < pps_encoder_ids_workspace_bytes(n0)) {
        set_error("pps_encoder_ids: workspace %zu < required %zu", workspace_bytes, pps_encoder_ids_workspace_bytes(n0));
        return PPS_ERR_WORKSPACE;
    }
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    // fork: side streams (created once per DEVICE) wait for the caller's stream, each owns one workspace slice
    struct DeviceStreams {
        cudaStream_t side[kIdStreams];
        cudaEvent_t fork, join[kIdStreams];
        bool ready;
    };
    static DeviceStreams per_device[kMaxDevices] = {};
    static unsigned char created[kMaxDevices] = {};
    int device = 0;
    PPS_CUDA(cudaGetDevice(&device));
    PPS_CHECK_ARG(device >= 0 && device < kMaxDevices, "pps_encoder_ids: device %d out of range", device);
    DeviceStreams& ds = per_device[device];
    if (first_use_on_device(created)) {
        PPS_CUDA(cudaEventCreateWithFlags(&ds.fork, cudaEventDisableTiming));
        for (int i = 0; i < kIdStreams; ++i) {
            PPS_CUDA(cudaStreamCreateWithFlags(&ds.side[i], cudaStreamNonBlocking));
            PPS_CUDA(cudaEventCreateWithFlags(&ds.join[i], cudaEventDisableTiming));
        }
        ds.ready = true;
    }
    PPS_CHECK_ARG(ds.ready, "pps_encoder_ids: the side streams of device %d could not be created", device);
    cudaStream_t* side = ds.side;
    cudaEvent_t ev_fork = ds.fork;
    cudaEvent_t* ev_join = ds.join;
    const int nstreams = b < kIdStreams ? (int)b : kIdStreams;
    PPS_CUDA(cudaEventRecord(ev_fork, user));
    for (int i = 0; i < nstreams; ++i) PPS_CUDA(cudaStreamWaitEvent(side[i], ev_fork, 0));
    int64_t n[5];
    level_sizes(n0, n);
    const size_t slice = encoder_ids_slice_bytes(n0);
    static const int pair16[9][2] = {{0, 0}, {0, 1}, {1, 1}, {1, 2}, {2, 2}, {2, 3}, {3, 3}, {3, 4}, {4, 4}};
    static const int pair1[4][2] = {{4, 3}, {3, 2}, {2, 1}, {1, 0}};
    // the loop body returns through `rc`: the join below must run even after a failure (the side streams still write the
    // caller's workspace and outputs)
    auto one_sample = [&](int64_t s) -> int {
        cudaStream_t st = side[s % nstreams];
        char* base = static_cast<char*>(workspace) + (s % nstreams) * slice;
        size_t off = 0;
        void* sample_ws = base;
        const size_t sample_bytes = align_up(pps_sample_workspace_bytes(n0), 256);
        off += sample_bytes;
        void* index[5];
        size_t index_bytes[5];
        for (int l = 0; l < 5; ++l) {
            index[l] = base + off;
            index_bytes[l] = pps_knn_index_bytes(n[l]);
            off += align_up(index_bytes[l], 256);
        }
        int32_t* sel = reinterpret_cast<int32_t*>(base + off);
        const float* lv[5];
        lv[0] = pts + s * n0 * 3;
        for (int l = 1; l < 5; ++l) {
            float* sup = out->support[l - 1] + s * n[l] * 3;
            PPS_TRY(sample_quantized_impl(lv[l - 1], n[l - 1], n[l], rotations + ((s * 4 + (l - 1)) * n_rot) * 9, n_rot,
                                          seed + (uint32_t)l, sample_ws, sample_bytes, sel, st));
            gather_points<<<(unsigned)ceil_div(3 * n[l], 256), 256, 0, st>>>(lv[l - 1], sel, (int)n[l], sup);
            PPS_LAUNCH_CHECK();
            lv[l] = sup;
        }
        for (int l = 0; l < 5; ++l) PPS_TRY(knn_build_impl(lv[l], n[l], index[l], index_bytes[l], st));
        for (int p = 0; p < 9; ++p) {
            const int a = pair16[p][0], c = pair16[p][1];
            const int k = (int)(n[a] < 16 ? n[a] : 16);
            PPS_TRY(knn_query_impl(index[a], n[a], lv[c], n[c], k, out->ids16[p] + s * n[c] * k, nullptr, st));
        }
        for (int p = 0; p < 4; ++p) {
            const int a = pair1[p][0], c = pair1[p][1];
            PPS_TRY(knn_query_impl(index[a], n[a], lv[c], n[c], 1, out->ids1[p] + s * n[c], nullptr, st));
        }
        return PPS_OK;
    };
    int rc = PPS_OK;
    for (int64_t s = 0; s < b && rc == PPS_OK; ++s) rc = one_sample(s);
    // join: the caller's stream continues when every side stream is done
    for (int i = 0; i < nstreams; ++i) {
        cudaError_t e = cudaEventRecord(ev_join[i], side[i]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(user, ev_join[i], 0);
        if (e != cudaSuccess && rc == PPS_OK) {
            set_error("pps_encoder_ids: join failed: %s", cudaGetErrorString(e));
            rc = PPS_ERR_CUDA;
        }
    }
    return rc;
}
}
