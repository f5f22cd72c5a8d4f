// Occupancy decoder: per-query kNN -> global attention-interpolation branch + local
// PointNet branch -> MLP -> softmax difference.
//
// Replaces  PPSurfNetwork.from_latent (source/ppsurf_model.py:82-117), InterpAttentionKHeadsNet.forward
// (source/poco_model.py:381-419), PointNetfeat.forward/STN/AttentionPoco (source/base/nn.py:305-373,162-190,84-96),
// MLP.forward (nn.py:415-417), _get_pts_local_ps / _predict_from_latent (source/poco_utils.py:67-82).
//
// Algebraic restructuring (exact in real arithmetic, fp32 rounding-level differences only):
//   * fc1([latent_j, q - p_j]) = U_j + W1_xyz.q  with the per-point table U_j = W1_lat.latent_j - W1_xyz.p_j + b1
//   * the attention weights sum to 1, so  fc8(sum_j a_j fc_value(h_j)) = (W8 Wv) (sum_j a_j h_j) + (W8 bv + b8)
//   * same for the PointNet attention pooling, which additionally commutes with the affine bn3(conv3(.))
// This file holds the fp32 SIMT path (path 0); the tensor-core path (path 1) lives in decode_tc.cu.
#include <algorithm>

#include "common.cuh"

namespace pps {

int linear_impl(const float* x, const float* w, const float* bias, const float* residual, const int32_t* gather,
                float* y, int64_t m, int n, int k, int ldx, int ldy, int act, cudaStream_t st);
int knn_query_impl(const void* index, int64_t n, const float* queries, int64_t q, int k, int32_t* idx_out,
                   float* d2_out, cudaStream_t st);
int projection_tc_impl(const pps_decoder_weights* w, const float* table, const float* queries, const int32_t* idx,
                       int k_stride, int64_t q, void* ws, size_t ws_bytes, float* pooled, cudaStream_t st);
size_t projection_tc_workspace(const pps_decoder_weights* w, int64_t chunk);
bool pointnet_tc_supported(const pps_decoder_weights* w);
bool chain_tc_supported(const pps_decoder_weights* w);
int mlp_tc_impl(const pps_decoder_weights* w, const float* pooled_proj, const float* pooled_pn, int64_t q, float* logits_out,
                float* occ_out, cudaStream_t st);
int pointnet_tc_impl(const pps_decoder_weights* w, const float* patches, int64_t q, float* a1, float* g, float* f1, float* f2,
                     float* tmat, float* pooled128, float* partial, cudaStream_t st);
size_t pointnet_tc_partial_floats(const pps_decoder_weights* w, int64_t q);

// ---------------------------------------------------------------------------------------------------------------
// patches (a7)
// ---------------------------------------------------------------------------------------------------------------
__global__ void patch_normalize_kernel(const float* __restrict__ pts, const float* __restrict__ queries,
                                       const int32_t* __restrict__ idx, const float* __restrict__ d2, long long q, int p,
                                       int ks, float* __restrict__ out) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= q * p) return;
    long long qi = e / p;
    int j = int(e % p);
    float r = __fsqrt_rn(d2[qi * ks + (p - 1)]);  // max_j ||p_j - q|| = distance of the farthest patch point
    int src = idx[qi * ks + j];
    float x = __fdiv_rn(__fsub_rn(pts[3 * (size_t)src + 0], queries[3 * qi + 0]), r);
    float y = __fdiv_rn(__fsub_rn(pts[3 * (size_t)src + 1], queries[3 * qi + 1]), r);
    float z = __fdiv_rn(__fsub_rn(pts[3 * (size_t)src + 2], queries[3 * qi + 2]), r);
    out[3 * e + 0] = x;
    out[3 * e + 1] = y;
    out[3 * e + 2] = z;
}

// ---------------------------------------------------------------------------------------------------------------
// global branch (a8)
// ---------------------------------------------------------------------------------------------------------------
// table[n,:] -= W1_xyz . p_n   (the W1_lat.latent + b1 part comes from linear_impl)
__global__ void table_xyz_kernel(const float* __restrict__ pts, const float* __restrict__ w1_xyz, long long n, int c,
                                 float* table) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * c) return;
    long long i = e / c;
    int ch = int(e % c);
    float v = w1_xyz[3 * ch] * pts[3 * i] + w1_xyz[3 * ch + 1] * pts[3 * i + 1] + w1_xyz[3 * ch + 2] * pts[3 * i + 2];
    table[e] -= v;
}

// h1[(q,j),:] = relu(U[idx[q,j],:] + W1_xyz . q);  one warp per row, C = 256 -> 2 float4 per lane
__global__ void proj_gather_kernel(const float* __restrict__ table, const float* __restrict__ queries,
                                   const int32_t* __restrict__ idx, const float* __restrict__ w1_xyz, long long q, int k,
                                   int ks, int c, float* __restrict__ h1) {
    long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= q * k) return;
    long long qi = row / k;
    int j = int(row % k);
    int src = idx[qi * ks + j];
    float qx = queries[3 * qi], qy = queries[3 * qi + 1], qz = queries[3 * qi + 2];
    const float4* urow = reinterpret_cast<const float4*>(table + (size_t)src * c);
    float4* orow = reinterpret_cast<float4*>(h1 + row * c);
    for (int v = lane; v < c / 4; v += 32) {
        float4 u = urow[v];
        float o[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int ch = 4 * v + t;
            float add = w1_xyz[3 * ch] * qx + w1_xyz[3 * ch + 1] * qy + w1_xyz[3 * ch + 2] * qz;
            o[t] = fmaxf(o[t] + add, 0.f);
        }
        orow[v] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// attention = mean_heads softmax_j(score[(q,j),h]);  pooled[q,:] = sum_j attention_j h3[(q,j),:]
// one block of 64 threads per query; k <= 64, heads == 64
__global__ void __launch_bounds__(64) attn_pool_kernel(const float* __restrict__ score, const float* __restrict__ h3,
                                                       int k, int heads, int c, float* __restrict__ pooled) {
    __shared__ float e[64][65];
    __shared__ float att[64];
    long long qi = blockIdx.x;
    int t = threadIdx.x;
    const float* s = score + qi * k * heads;
    if (t < heads) {
        float m = -INFINITY;
        for (int j = 0; j < k; ++j) m = fmaxf(m, s[j * heads + t]);
        float sum = 0.f;
        for (int j = 0; j < k; ++j) {
            float v = expf(s[j * heads + t] - m);
            e[j][t] = v;
            sum += v;
        }
        float inv = 1.f / sum;
        for (int j = 0; j < k; ++j) e[j][t] *= inv;
    }
    __syncthreads();
    if (t < k) {
        float a = 0.f;
        for (int h = 0; h < heads; ++h) a += e[t][h];
        att[t] = a / float(heads);
    }
    __syncthreads();
    const float* hrow = h3 + qi * k * c;
    for (int ch = t; ch < c; ch += 64) {
        float acc = 0.f;
        for (int j = 0; j < k; ++j) acc = fmaf(att[j], hrow[(size_t)j * c + ch], acc);
        pooled[qi * c + ch] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// local branch (a9)
// ---------------------------------------------------------------------------------------------------------------
// a0[m, 0..63] = relu(W0a . x_m + b0a);  16 threads per point, 4 channels each
__global__ void pn_conv0a_kernel(const float* __restrict__ patches, const float* __restrict__ w, const float* __restrict__ b,
                                 long long m, float* __restrict__ a0) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long pt = e >> 4;
    int cq = int(e & 15);
    if (pt >= m) return;
    float x = patches[3 * pt], y = patches[3 * pt + 1], z = patches[3 * pt + 2];
    float o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int ch = 4 * cq + t;
        o[t] = fmaxf(w[3 * ch] * x + w[3 * ch + 1] * y + w[3 * ch + 2] * z + b[ch], 0.f);
    }
    reinterpret_cast<float4*>(a0)[e] = make_float4(o[0], o[1], o[2], o[3]);
}

// g[q,c] = max_p t[(q,p),c]
__global__ void segment_max_kernel(const float* __restrict__ t, long long q, int p, int c, float* __restrict__ g) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= q * c) return;
    long long qi = e / c;
    int ch = int(e % c);
    float m = -INFINITY;
    for (int j = 0; j < p; ++j) m = fmaxf(m, t[(qi * p + j) * c + ch]);
    g[e] = m;
}

// x'[(q,p),i] = sum_j T[q,i,j] a1[(q,p),j]   (feature transform, nn.py:329); block = 256 threads per query
__global__ void __launch_bounds__(256) stn_apply_kernel(const float* __restrict__ tmat, const float* __restrict__ a1, int p,
                                                        float* __restrict__ out) {
    __shared__ float T[64][65];
    __shared__ float X[4][64];
    long long qi = blockIdx.x;
    int tid = threadIdx.x;
    const float* tq = tmat + qi * 4096;
    for (int e = tid; e < 4096; e += 256) T[e >> 6][e & 63] = tq[e];
    int i = tid & 63, sub = tid >> 6;
    for (int p0 = 0; p0 < p; p0 += 4) {
        __syncthreads();
        int pp = p0 + sub;
        if (pp < p) X[sub][i] = a1[(qi * p + pp) * 64 + i];
        __syncthreads();
        if (pp < p) {
            float acc = 0.f;
#pragma unroll 16
            for (int j = 0; j < 64; ++j) acc = fmaf(T[i][j], X[sub][j], acc);
            out[(qi * p + pp) * 64 + i] = acc;
        }
    }
}

// attention pooling over the patch: w = softmax_p(wq . c2_p + bq); pooled[q,:] = sum_p w_p c2[(q,p),:]  (128 channels)
__global__ void __launch_bounds__(128) pn_attpool_kernel(const float* __restrict__ c2, const float* __restrict__ wq, float bq,
                                                         int p, float* __restrict__ pooled) {
    extern __shared__ float sm[];  // logits[p]
    __shared__ float red[4];
    long long qi = blockIdx.x;
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* base = c2 + qi * p * 128;
    for (int j = warp; j < p; j += 4) {
        float4 v = reinterpret_cast<const float4*>(base + (size_t)j * 128)[lane];
        float4 ww = reinterpret_cast<const float4*>(wq)[lane];
        float d = v.x * ww.x + v.y * ww.y + v.z * ww.z + v.w * ww.w;
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0) sm[j] = d + bq;
    }
    __syncthreads();
    float m = -INFINITY;
    for (int j = tid; j < p; j += 128) m = fmaxf(m, sm[j]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float s = 0.f;
    for (int j = tid; j < p; j += 128) {
        float v = expf(sm[j] - m);
        sm[j] = v;
        s += v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float inv = 1.f / (red[0] + red[1] + red[2] + red[3]);
    float acc = 0.f;
    for (int j = 0; j < p; ++j) acc = fmaf(sm[j] * inv, base[(size_t)j * 128 + tid], acc);
    pooled[qi * 128 + tid] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// MLP head (a10) + occupancy (a11)
// ---------------------------------------------------------------------------------------------------------------
// logits = W2 . x + b2 (2 outputs), occ = softmax(l)[0] - softmax(l)[1];  one warp per query, C = 256
__global__ void mlp_head_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                long long q, int c, float* logits, float* occ) {
    long long qi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (qi >= q) return;
    float a0 = 0.f, a1 = 0.f;
    for (int ch = lane; ch < c; ch += 32) {
        float v = x[qi * c + ch];
        a0 = fmaf(v, w[ch], a0);
        a1 = fmaf(v, w[c + ch], a1);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if (lane == 0) {
        float l0 = a0 + b[0], l1 = a1 + b[1];
        if (logits) {
            logits[2 * qi] = l0;
            logits[2 * qi + 1] = l1;
        }
        if (occ) {
            float m = fmaxf(l0, l1);
            float e0 = expf(l0 - m), e1 = expf(l1 - m);
            float s = e0 + e1;
            occ[qi] = e0 / s - e1 / s;
        }
    }
}

__global__ void grid_queries_kernel(int r, float step, float bmin_pad, long long first, long long count, float* out) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    long long v = first + e;
    int iz = int(v % r), iy = int((v / r) % r), ix = int(v / ((long long)r * r));
    out[3 * e + 0] = __fadd_rn(__fmul_rn(float(ix), step), bmin_pad);
    out[3 * e + 1] = __fadd_rn(__fmul_rn(float(iy), step), bmin_pad);
    out[3 * e + 2] = __fadd_rn(__fmul_rn(float(iz), step), bmin_pad);
}

// ---------------------------------------------------------------------------------------------------------------
// host-side pipeline
// ---------------------------------------------------------------------------------------------------------------
static inline int kmax_of(const pps_decoder_weights* w) { return w->k > w->num_pts_local ? w->k : w->num_pts_local; }
// The neighbour search runs over a SUPER-CHUNK of kKnnSuper decode chunks at once: the larger the launch, the better the search's tail
// is hidden (the whole 131^3 grid in one launch: 40 ms; in four: 46).  The global branch runs over the super-chunk in slices of at
// most kProjQueries queries: one launch over 2.25 M queries loses the fc1 table from the L2 (110.8 ms instead of 106).
constexpr int kKnnSuper = 16;
constexpr int64_t kProjQueries = 606208;
static int64_t super_chunks(int64_t) { return kKnnSuper; }

struct DecodeBuffers {
    int32_t* idx;
    float* d2;
    float* bufA;
    float* bufB;
    float* score;
    float* a1;
    float* patches;
    float* pooled;
    float* pooled_super;  // global-branch output of a whole super-chunk (tensor-core path)
    float* feat_proj;
    float* g;
    float* f1;
    float* f2;
    float* tmat;
    float* pooled128;
    float* pn_partial;  // attention-pooling partials of patches that span several half-tiles (tensor-core path, P > 64)
    float* feat;
    float* m0;
    float* m1;
    void* tc_ws;
    size_t tc_ws_bytes;
};

static bool carve(const pps_decoder_weights* w, int64_t chunk, void* ws, size_t ws_bytes, DecodeBuffers& b, size_t* need) {
    Arena a(ws, ws_bytes);
    const int C = w->latent, P = w->num_pts_local, K = w->k, S = w->stn_size;
    const int kmax = kmax_of(w);
    size_t rows = (size_t)chunk * (size_t)(K > P ? K : P);
    size_t wide = C > S ? C : S;
    const size_t nsuper = (size_t)chunk * super_chunks(chunk);
    b.idx = a.take<int32_t>(nsuper * kmax);
    b.d2 = a.take<float>(nsuper * kmax);
    b.bufA = a.take<float>(rows * wide);
    b.bufB = a.take<float>(rows * wide);
    b.score = a.take<float>((size_t)chunk * K * w->heads);
    b.a1 = a.take<float>((size_t)chunk * ((P + 63) / 64) * 64 * 64 + 8192);  // tensor-core path: tile-major, 64 point slots per half-tile
    b.patches = a.take<float>((size_t)chunk * P * 3);
    b.pooled = a.take<float>((size_t)chunk * C);
    b.pooled_super = a.take<float>(nsuper * C);
    b.feat_proj = a.take<float>((size_t)chunk * C);
    b.g = a.take<float>((size_t)chunk * S);
    b.f1 = a.take<float>((size_t)chunk * (S / 2));
    b.f2 = a.take<float>((size_t)chunk * (S / 4));
    b.tmat = a.take<float>((size_t)chunk * 4096);
    b.pooled128 = a.take<float>((size_t)chunk * 128);
    b.pn_partial = a.take<float>(pointnet_tc_partial_floats(w, chunk) + 4);
    b.feat = a.take<float>((size_t)chunk * C);
    b.m0 = a.take<float>((size_t)chunk * C);
    b.m1 = a.take<float>((size_t)chunk * C);
    b.tc_ws_bytes = projection_tc_workspace(w, chunk);
    b.tc_ws = a.take<char>(b.tc_ws_bytes);
    if (need) *need = align_up(a.off, 256);
    return a.ok();
}

static int check_weights(const pps_decoder_weights* w) {
    PPS_CHECK_ARG(w, "decoder weights are null");
    PPS_CHECK_ARG(w->latent == 256 && w->heads == 64, "decoder kernels are built for latent=256, heads=64 (got %d, %d)",
                  w->latent, w->heads);
    PPS_CHECK_ARG(w->k >= 1 && w->k <= 64, "decoder k=%d must be in [1,64]", w->k);
    PPS_CHECK_ARG(w->num_pts_local >= 1 && w->num_pts_local <= 512, "num_pts_local=%d out of range", w->num_pts_local);
    PPS_CHECK_ARG(w->stn_size % 4 == 0 && w->stn_size >= 64, "stn_size=%d unsupported", w->stn_size);
    return PPS_OK;
}

// global branch for q queries with given neighbour ids -> feat_out [q,C]  (fc8 applied)
static int projection_run(const pps_decoder_weights* w, const float* table, const float* queries, const int32_t* idx,
                          int ks, int64_t q, DecodeBuffers& b, float* feat_out, int path, cudaStream_t st) {
    const int C = w->latent, K = w->k;
    if (path == 1) {
        PPS_TRY(projection_tc_impl(w, table, queries, idx, ks, q, b.tc_ws, b.tc_ws_bytes, b.pooled, st));
    } else {
        int64_t rows = q * K;
        proj_gather_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, st>>>(table, queries, idx, w->w1_xyz, q, K, ks, C, b.bufA);
        PPS_LAUNCH_CHECK();
        profile_begin(st);
        PPS_TRY(linear_impl(b.bufA, w->w2, w->b2, nullptr, nullptr, b.bufB, rows, C, C, C, C, 1, st));
        PPS_TRY(linear_impl(b.bufB, w->w3, w->b3, nullptr, nullptr, b.bufA, rows, C, C, C, C, 1, st));
        PPS_TRY(linear_impl(b.bufA, w->wq, w->bq, nullptr, nullptr, b.score, rows, w->heads, C, C, w->heads, 0, st));
        profile_end(st);
        attn_pool_kernel<<<(unsigned)q, 64, 0, st>>>(b.score, b.bufA, K, w->heads, C, b.pooled);
        PPS_LAUNCH_CHECK();
    }
    PPS_TRY(linear_impl(b.pooled, w->wv8, w->bv8, nullptr, nullptr, feat_out, q, C, C, C, C, 0, st));
    return PPS_OK;
}

// local branch for q patches [q,P,3] -> feat_out [q,C]; `residual` (nullable) is added (sum of the branches)
static int pointnet_run(const pps_decoder_weights* w, const float* patches, int64_t q, DecodeBuffers& b,
                        const float* residual, float* feat_out, int path, cudaStream_t st) {
    const int C = w->latent, P = w->num_pts_local, S = w->stn_size;
    if (path == 1 && pointnet_tc_supported(w)) {
        PPS_TRY(pointnet_tc_impl(w, patches, q, b.a1, b.g, b.f1, b.f2, b.tmat, b.pooled128, b.pn_partial, st));
        return linear_impl(b.pooled128, w->pnv_w, w->pnv_b, residual, nullptr, feat_out, q, C, 128, 128, C, 0, st);
    }
    int64_t m = q * P;
    pn_conv0a_kernel<<<(unsigned)ceil_div(m * 16, 256), 256, 0, st>>>(patches, w->pn0a_w, w->pn0a_b, m, b.bufA);
    PPS_LAUNCH_CHECK();
    PPS_TRY(linear_impl(b.bufA, w->pn0b_w, w->pn0b_b, nullptr, nullptr, b.a1, m, 64, 64, 64, 64, 1, st));
    PPS_TRY(linear_impl(b.a1, w->stn1_w, w->stn1_b, nullptr, nullptr, b.bufA, m, 64, 64, 64, 64, 1, st));
    PPS_TRY(linear_impl(b.bufA, w->stn2_w, w->stn2_b, nullptr, nullptr, b.bufB, m, 128, 64, 64, 128, 1, st));
    PPS_TRY(linear_impl(b.bufB, w->stn3_w, w->stn3_b, nullptr, nullptr, b.bufA, m, S, 128, 128, S, 1, st));
    segment_max_kernel<<<(unsigned)ceil_div(q * S, 256), 256, 0, st>>>(b.bufA, q, P, S, b.g);
    PPS_LAUNCH_CHECK();
    PPS_TRY(linear_impl(b.g, w->stnf1_w, w->stnf1_b, nullptr, nullptr, b.f1, q, S / 2, S, S, S / 2, 1, st));
    PPS_TRY(linear_impl(b.f1, w->stnf2_w, w->stnf2_b, nullptr, nullptr, b.f2, q, S / 4, S / 2, S / 2, S / 4, 1, st));
    PPS_TRY(linear_impl(b.f2, w->stnf3_w, w->stnf3_b, nullptr, nullptr, b.tmat, q, 4096, S / 4, S / 4, 4096, 0, st));
    stn_apply_kernel<<<(unsigned)q, 256, 0, st>>>(b.tmat, b.a1, P, b.bufA);
    PPS_LAUNCH_CHECK();
    PPS_TRY(linear_impl(b.bufA, w->pn1_w, w->pn1_b, nullptr, nullptr, b.bufB, m, 64, 64, 64, 64, 1, st));
    PPS_TRY(linear_impl(b.bufB, w->pn2_w, w->pn2_b, nullptr, nullptr, b.bufA, m, 128, 64, 64, 128, 1, st));
    pn_attpool_kernel<<<(unsigned)q, 128, P * sizeof(float), st>>>(b.bufA, w->pnq_w, w->pnq_b, P, b.pooled128);
    PPS_LAUNCH_CHECK();
    PPS_TRY(linear_impl(b.pooled128, w->pnv_w, w->pnv_b, residual, nullptr, feat_out, q, C, 128, 128, C, 0, st));
    return PPS_OK;
}

// one decode chunk; `idx`/`d2` [q,kmax] are this chunk's rows of the neighbour search
static int decode_chunk(const pps_decoder_weights* w, const float* pts, const float* table, const float* queries, int64_t q,
                        const int32_t* idx, const float* d2, DecodeBuffers& b, const float* pooled_proj, float* logits_out,
                        float* occ_out, int path, cudaStream_t st) {
    const int C = w->latent, P = w->num_pts_local;
    const int kmax = kmax_of(w);
    patch_normalize_kernel<<<(unsigned)ceil_div(q * P, 256), 256, 0, st>>>(pts, queries, idx, d2, q, P, kmax, b.patches);
    PPS_LAUNCH_CHECK();
    if (path == 1 && pointnet_tc_supported(w) && chain_tc_supported(w)) {
        // all-tensor-core tail: both pooled vectors go straight into the chain kernel (merged value matrices + MLP + head);
        // the global branch of the whole super-chunk has already run (decode_super), `pooled_proj` are this chunk's rows
        PPS_TRY(pointnet_tc_impl(w, b.patches, q, b.a1, b.g, b.f1, b.f2, b.tmat, b.pooled128, b.pn_partial, st));
        return mlp_tc_impl(w, pooled_proj, b.pooled128, q, logits_out, occ_out, st);
    }
    PPS_TRY(projection_run(w, table, queries, idx, kmax, q, b, b.feat_proj, path, st));
    PPS_TRY(pointnet_run(w, b.patches, q, b, b.feat_proj, b.feat, path, st));
    PPS_TRY(linear_impl(b.feat, w->m0_w, w->m0_b, nullptr, nullptr, b.m0, q, C, C, C, C, 1, st));
    PPS_TRY(linear_impl(b.m0, w->m1_w, w->m1_b, nullptr, nullptr, b.m1, q, C, C, C, C, 1, st));
    mlp_head_kernel<<<(unsigned)ceil_div(q * 32, 256), 256, 0, st>>>(b.m1, w->m2_w, w->m2_b, q, C, logits_out, occ_out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

// neighbour search for up to kKnnSuper chunks, then the chunks one after the other
static int decode_super(const pps_decoder_weights* w, const void* knn_index, const float* pts, const float* table, int64_t n,
                        const float* queries, int64_t q, int64_t chunk, DecodeBuffers& b, float* logits_out, float* occ_out,
                        int32_t* idx_out, int path, cudaStream_t st) {
    const int kmax = kmax_of(w);
    PPS_TRY(knn_query_impl(knn_index, n, queries, q, kmax, b.idx, b.d2, st));
    if (idx_out) PPS_CUDA(cudaMemcpyAsync(idx_out, b.idx, (size_t)q * kmax * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    // tensor-core path: the global branch runs over the whole super-chunk in ONE launch.  Its working set (the 97 MB fc1 table
    // + the weight pack) then stays in the L2 instead of being evicted between chunks by the local branch, which streams
    // ~0.6 GB of a1 / T per chunk through the cache
    const bool all_tc = path == 1 && pointnet_tc_supported(w) && chain_tc_supported(w);
    if (all_tc) {
        for (int64_t s = 0; s < q; s += kProjQueries) {
            const int64_t c = std::min<int64_t>(kProjQueries, q - s);
            PPS_TRY(projection_tc_impl(w, table, queries + 3 * s, b.idx + s * kmax, kmax, c, b.tc_ws, b.tc_ws_bytes,
                                       b.pooled_super + (size_t)s * w->latent, st));
        }
    }
    // equal chunks: a super-chunk of 7.4 nominal chunks (one rank's share of the 131^3 grid at 8 GPUs) runs as 8 launches of 93 %
    // instead of 7 full ones and a 42 % tail -- the persistent kernels of a chunk cost nearly the same whatever their fill
    const int64_t pieces = ceil_div(q, chunk);
    const int64_t even = pieces > 0 ? std::min<int64_t>(chunk, (int64_t)align_up((size_t)ceil_div(q, pieces), 256)) : chunk;
    for (int64_t s = 0; s < q; s += even) {
        int64_t c = q - s < even ? q - s : even;
        PPS_TRY(decode_chunk(w, pts, table, queries + 3 * s, c, b.idx + s * kmax, b.d2 + s * kmax, b,
                             all_tc ? b.pooled_super + (size_t)s * w->latent : nullptr, logits_out ? logits_out + 2 * s : nullptr,
                             occ_out ? occ_out + s : nullptr, path, st));
    }
    return PPS_OK;
}

}  // namespace pps

using namespace pps;

extern "C" {

int pps_patch_normalize(const float* pts, const float* queries, const int32_t* idx, const float* dist2, int64_t q,
                        int p, int k_stride, float* out, void* stream) {
    PPS_CHECK_ARG(pts && queries && idx && dist2 && out, "pps_patch_normalize: null pointer");
    PPS_CHECK_ARG(p >= 1 && p <= k_stride, "pps_patch_normalize: p=%d k_stride=%d", p, k_stride);
    if (q == 0) return PPS_OK;
    patch_normalize_kernel<<<(unsigned)ceil_div(q * p, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pts, queries, idx, dist2, q, p, k_stride, out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_grid_queries(int r, float step, float bmin_pad, int64_t first, int64_t count, float* out, void* stream) {
    PPS_CHECK_ARG(out && r > 0 && first >= 0 && first + count <= (int64_t)r * r * r, "pps_grid_queries: bad range");
    if (count == 0) return PPS_OK;
    grid_queries_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(r, step, bmin_pad, first,
                                                                                                    count, out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_decoder_point_table(const pps_decoder_weights* w, const float* pts, const float* latents, int64_t n,
                            float* table, void* stream) {
    PPS_TRY(check_weights(w));
    PPS_CHECK_ARG(pts && latents && table && n > 0, "pps_decoder_point_table: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int C = w->latent;
    PPS_TRY(linear_impl(latents, w->w1_lat, w->b1, nullptr, nullptr, table, n, C, C, C, C, 0, st));
    table_xyz_kernel<<<(unsigned)ceil_div(n * C, 256), 256, 0, st>>>(pts, w->w1_xyz, n, C, table);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

size_t pps_decoder_workspace_bytes(const pps_decoder_weights* w, int64_t chunk) {
    if (!w || chunk <= 0) return 0;
    DecodeBuffers b;
    size_t need = 0;
    carve(w, chunk, nullptr, 0, b, &need);
    return need;
}

int pps_decoder_decode(const pps_decoder_weights* w, const void* knn_index, const float* pts, const float* table,
                       int64_t n, const float* queries, int64_t q, int64_t chunk, void* workspace,
                       size_t workspace_bytes, float* logits_out, float* occ_out, int32_t* idx_out, int path,
                       void* stream) {
    PPS_TRY(check_weights(w));
    PPS_CHECK_ARG(knn_index && pts && table && queries && workspace, "pps_decoder_decode: null pointer");
    PPS_CHECK_ARG(chunk > 0 && q >= 0 && n >= kmax_of(w), "pps_decoder_decode: chunk=%lld q=%lld n=%lld (need n >= %d)",
                  (long long)chunk, (long long)q, (long long)n, kmax_of(w));
    PPS_CHECK_ARG(path == 0 || path == 1, "pps_decoder_decode: unknown path %d", path);
    DecodeBuffers b;
    size_t need = 0;
    if (!carve(w, chunk, workspace, workspace_bytes, b, &need)) {
        set_error("pps_decoder_decode: workspace %zu < required %zu", workspace_bytes, need);
        return PPS_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int kmax = kmax_of(w);
    const int64_t super = chunk * super_chunks(chunk);
    for (int64_t s = 0; s < q; s += super) {
        int64_t c = q - s < super ? q - s : super;
        PPS_TRY(decode_super(w, knn_index, pts, table, n, queries + 3 * s, c, chunk, b, logits_out ? logits_out + 2 * s : nullptr,
                             occ_out ? occ_out + s : nullptr, idx_out ? idx_out + s * kmax : nullptr, path, st));
    }
    return PPS_OK;
}

int pps_decoder_decode_host(const pps_decoder_weights* w, const void* knn_index, const float* pts, const float* table,
                            int64_t n, const float* queries_host, int64_t q, int64_t chunk, void* workspace,
                            size_t workspace_bytes, void* staging, size_t staging_bytes, float* occ_host, int path,
                            void* stream, void* copy_stream) {
    PPS_TRY(check_weights(w));
    PPS_CHECK_ARG(knn_index && pts && table && queries_host && workspace && staging && occ_host,
                  "pps_decoder_decode_host: null pointer");
    PPS_CHECK_ARG(chunk > 0 && q >= 0 && n >= kmax_of(w), "pps_decoder_decode_host: bad sizes");
    if (staging_bytes < (size_t)q * 16) {
        set_error("pps_decoder_decode_host: staging %zu < %zu", staging_bytes, (size_t)q * 16);
        return PPS_ERR_WORKSPACE;
    }
    DecodeBuffers b;
    size_t need = 0;
    if (!carve(w, chunk, workspace, workspace_bytes, b, &need)) {
        set_error("pps_decoder_decode_host: workspace %zu < required %zu", workspace_bytes, need);
        return PPS_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream), cs = static_cast<cudaStream_t>(copy_stream);
    float* dq = static_cast<float*>(staging);
    float* docc = dq + 3 * q;
    const int64_t super = chunk * super_chunks(chunk);
    int64_t nchunks = ceil_div(q, super);
    // upload on the copy stream, compute on `stream`, download on the copy stream; the upload of super-chunk i+1 is queued
    // before the download of super-chunk i so that it overlaps the decode.  Four events per device, created once
    // (pps::device_events) and used alternately: a cudaStreamWaitEvent captures the record that precedes it on the host.
    cudaEvent_t* ev = device_events(4);
    if (ev == nullptr) return PPS_ERR_CUDA;
    cudaEvent_t* up = ev;
    cudaEvent_t* done = ev + 2;
    int rc = PPS_OK;
    cudaError_t ce = cudaSuccess;
    auto upload = [&](int64_t i) {
        const int64_t s = i * super, c = q - s < super ? q - s : super;
        ce = cudaMemcpyAsync(dq + 3 * s, queries_host + 3 * s, (size_t)c * 12, cudaMemcpyHostToDevice, cs);
        if (ce == cudaSuccess) ce = cudaEventRecord(up[i & 1], cs);
    };
    if (nchunks > 0) upload(0);
    for (int64_t i = 0; i < nchunks && rc == PPS_OK && ce == cudaSuccess; ++i) {
        const int64_t s = i * super, c = q - s < super ? q - s : super;
        ce = cudaStreamWaitEvent(st, up[i & 1], 0);
        if (ce != cudaSuccess) break;
        if (i + 1 < nchunks) upload(i + 1);
        if (ce != cudaSuccess) break;
        rc = decode_super(w, knn_index, pts, table, n, dq + 3 * s, c, chunk, b, nullptr, docc + s, nullptr, path, st);
        if (rc != PPS_OK) break;
        ce = cudaEventRecord(done[i & 1], st);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(cs, done[i & 1], 0);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(occ_host + s, docc + s, (size_t)c * 4, cudaMemcpyDeviceToHost, cs);
    }
    // always drain both streams before returning: the caller owns the buffers again after this call
    cudaError_t e1 = cudaStreamSynchronize(st), e2 = cudaStreamSynchronize(cs);
    if (rc == PPS_OK && ce != cudaSuccess) {
        set_error("pps_decoder_decode_host: %s", cudaGetErrorString(ce));
        return PPS_ERR_CUDA;
    }
    if (rc != PPS_OK) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        set_error("pps_decoder_decode_host: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        return PPS_ERR_CUDA;
    }
    return PPS_OK;
}

int pps_decoder_projection(const pps_decoder_weights* w, const float* pts, const float* table, const float* queries,
                           const int32_t* idx, int k_stride, int64_t q, void* workspace, size_t workspace_bytes,
                           float* feat_out, int path, void* stream) {
    (void)pts;
    PPS_TRY(check_weights(w));
    PPS_CHECK_ARG(table && queries && idx && workspace && feat_out && k_stride >= w->k, "pps_decoder_projection: bad arguments");
    PPS_CHECK_ARG(path == 0 || path == 1, "pps_decoder_projection: unknown path %d", path);
    if (q == 0) return PPS_OK;
    DecodeBuffers b;
    size_t need = 0;
    if (!carve(w, q, workspace, workspace_bytes, b, &need)) {
        set_error("pps_decoder_projection: workspace %zu < required %zu", workspace_bytes, need);
        return PPS_ERR_WORKSPACE;
    }
    return projection_run(w, table, queries, idx, k_stride, q, b, feat_out, path, static_cast<cudaStream_t>(stream));
}

int pps_decoder_pointnet(const pps_decoder_weights* w, const float* patches, int64_t q, void* workspace,
                         size_t workspace_bytes, float* feat_out, int path, void* stream) {
    PPS_TRY(check_weights(w));
    PPS_CHECK_ARG(patches && workspace && feat_out, "pps_decoder_pointnet: null pointer");
    if (q == 0) return PPS_OK;
    DecodeBuffers b;
    size_t need = 0;
    if (!carve(w, q, workspace, workspace_bytes, b, &need)) {
        set_error("pps_decoder_pointnet: workspace %zu < required %zu", workspace_bytes, need);
        return PPS_ERR_WORKSPACE;
    }
    PPS_CHECK_ARG(path == 0 || path == 1, "pps_decoder_pointnet: unknown path %d", path);
    return pointnet_run(w, patches, q, b, nullptr, feat_out, path, static_cast<cudaStream_t>(stream));
}
}
