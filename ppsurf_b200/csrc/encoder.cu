// FKAConv encoder kernels, point-major activations.
//
// Replaces FKAConvLayer.forward (source/base/nn.py:592-652), max_pool (nn.py:677-680), the global max of
// FKAConvNetwork.forward (nn.py:531) and the latent scatter of the predict loop (source/poco_model.py:228-234).
//
// FKAConv = (1) a tiny "kernel-weight" MLP on the 16 centred neighbour offsets of every support point that yields a
// 16x16 matrix per point, with two InstanceNorm2d layers whose statistics run over ALL (point, neighbour) pairs of a
// sample -> three launches of fka_weight_kernel (stats 1, stats 2, weights); (2) feat[n,m,c] = sum_j mat[n,j,m] *
// x[ids[n,j],c]  (fused gather * weights, fka_feat_kernel); (3) a dense contraction with the [cout,16*cin] kernel
// (linear_impl, BatchNorm/ReLU folded into its epilogue).
#include "common.cuh"

namespace pps {

int linear_impl(const float* x, const float* w, const float* bias, const float* residual, const int32_t* gather,
                float* y, int64_t m, int n, int k, int ldx, int ldy, int act, cudaStream_t st);
bool fka_fused_supported(const pps_fkaconv_weights* w, int kn, int64_t n_s);
int fka_fused_impl(const pps_fkaconv_weights* w, const float* x, const float* pts, const float* support, const int32_t* ids,
                   int64_t b, int64_t n_in, int64_t n_s, double* stats, float* out, cudaStream_t st);

constexpr int kNbr = 16;   // max neighbours per support point (the reference always asks for 16, clamped to n_in)
constexpr float kInEps = 1e-5f;

struct FkaParams {
    float alpha, beta, inv_radius;
    int act;
};

__device__ __forceinline__ float fka_act(float v, int act) { return act == 1 ? v / (1.f + expf(-v)) : fmaxf(v, 0.f); }

// PHASE 1: sum / sumsq of fc1 outputs; PHASE 2: sum / sumsq of fc2 outputs; PHASE 3: write mat [b,ns,kn(j),16(m)]
// 16 lanes per support point (one per neighbour), 16 points per block of 256 threads: the max over the neighbourhood and
// the distance-weight normalisation are half-warp shuffles, the j-independent halves of fc2 / fc3 (the pooled maxima)
// are computed once per point (lane o owns output o) and shared by shuffles.
constexpr int kFkaPts = 16;
template <int PHASE>
__global__ void __launch_bounds__(256) fka_weight_kernel(const float* __restrict__ pts, const float* __restrict__ support,
                                                         const int32_t* __restrict__ ids, int n_in, int n_s, int kn, FkaParams prm,
                                                         const float* __restrict__ fc1, const float* __restrict__ fc2,
                                                         const float* __restrict__ fc3, const float* __restrict__ in1_w,
                                                         const float* __restrict__ in1_b, const float* __restrict__ in2_w,
                                                         const float* __restrict__ in2_b, double* stats, float* __restrict__ mat) {
    __shared__ float W1[16 * 3], W2[16 * 32], W3[16 * 32];
    __shared__ float A1[16], B1[16], A2[16], B2[16];
    __shared__ float red[8][32];
    const unsigned int full = 0xffffffffu;
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = tid & 15;  // neighbour slot of this lane
    double* st_b = stats + (size_t)b * 64;  // [phase(2)][sum16, sumsq16]
    for (int e = tid; e < 48; e += 256) W1[e] = fc1[e];
    for (int e = tid; e < 512; e += 256) {
        W2[e] = fc2[e];
        W3[e] = fc3[e];
    }
    const double cnt = double(n_s) * kn;
    // with a single neighbour the reference skips both InstanceNorms (nn.py:627-628,635-636)
    if (PHASE >= 2 && tid < 16) {
        double mean = st_b[tid] / cnt;
        double var = st_b[16 + tid] / cnt - mean * mean;
        float rstd = float(1.0 / sqrt(fmax(var, 0.0) + double(kInEps)));
        A1[tid] = kn == 1 ? 1.f : rstd * in1_w[tid];
        B1[tid] = kn == 1 ? 0.f : in1_b[tid] - float(mean) * rstd * in1_w[tid];
    }
    if (PHASE >= 3 && tid < 16) {
        double mean = st_b[32 + tid] / cnt;
        double var = st_b[48 + tid] / cnt - mean * mean;
        float rstd = float(1.0 / sqrt(fmax(var, 0.0) + double(kInEps)));
        A2[tid] = kn == 1 ? 1.f : rstd * in2_w[tid];
        B2[tid] = kn == 1 ? 0.f : in2_b[tid] - float(mean) * rstd * in2_w[tid];
    }
    __syncthreads();

    const int n = blockIdx.x * kFkaPts + (tid >> 4);
    const bool valid = n < n_s && j < kn;
    const size_t row = (size_t)b * n_s + (n < n_s ? n : 0);
    float rx = 0.f, ry = 0.f, rz = 0.f, wgt = 0.f;
    if (valid) {
        const size_t src = (size_t)b * n_in + ids[row * kn + j];
        rx = pts[3 * src] - support[3 * row];
        ry = pts[3 * src + 1] - support[3 * row + 1];
        rz = pts[3 * src + 2] - support[3 * row + 2];
        const float dist = sqrtf(rx * rx + ry * ry + rz * rz);
        rx *= prm.inv_radius;
        ry *= prm.inv_radius;
        rz *= prm.inv_radius;
        wgt = 1.f / (1.f + expf(-(-prm.alpha * dist + prm.beta)));
    }
    float dsum = wgt;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) dsum += __shfl_xor_sync(full, dsum, o, 16);
    dsum = dsum + (dsum == 0.f ? 1.f : 0.f) + 1e-6f;
    const float dw = wgt / dsum * float(kn);

    float y1[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) y1[c] = W1[3 * c] * rx + W1[3 * c + 1] * ry + W1[3 * c + 2] * rz;

    float s[16], ss[16];  // statistics of this lane (PHASE 1, 2)
    if (PHASE == 1) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            s[c] = valid ? y1[c] : 0.f;
            ss[c] = valid ? y1[c] * y1[c] : 0.f;
        }
    } else {
        // m1 = act(IN1(fc1 rel)); mp1 = max over the neighbourhood of m1 * dw
        float m1[16], mp[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            m1[c] = fka_act(y1[c] * A1[c] + B1[c], prm.act);
            float v = valid ? m1[c] * dw : -INFINITY;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(full, v, o, 16));
            mp[c] = v;
        }
        // fc2 on cat(m1, mp1): lane o owns the j-independent half of output o
        float cown = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) cown = fmaf(W2[j * 32 + 16 + c], mp[c], cown);
        float y2[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            float y = __shfl_sync(full, cown, o, 16);
#pragma unroll
            for (int c = 0; c < 16; ++c) y = fmaf(W2[o * 32 + c], m1[c], y);
            y2[o] = y;
        }
        if (PHASE == 2) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                s[c] = valid ? y2[c] : 0.f;
                ss[c] = valid ? y2[c] * y2[c] : 0.f;
            }
        } else {
            float m2[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                m2[c] = fka_act(y2[c] * A2[c] + B2[c], prm.act);
                float v = valid ? m2[c] * dw : -INFINITY;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(full, v, o, 16));
                mp[c] = v;
            }
            cown = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) cown = fmaf(W3[j * 32 + 16 + c], mp[c], cown);
            float outv[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                float y = __shfl_sync(full, cown, o, 16);
#pragma unroll
                for (int c = 0; c < 16; ++c) y = fmaf(W3[o * 32 + c], m2[c], y);
                outv[o] = fka_act(y, prm.act) * dw;
            }
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(mat + (row * kn + j) * 16);
                dst[0] = make_float4(outv[0], outv[1], outv[2], outv[3]);
                dst[1] = make_float4(outv[4], outv[5], outv[6], outv[7]);
                dst[2] = make_float4(outv[8], outv[9], outv[10], outv[11]);
                dst[3] = make_float4(outv[12], outv[13], outv[14], outv[15]);
            }
        }
    }
    if (PHASE <= 2) {
        // block reduction of the 32 statistics: lanes -> warp sums (recursive halving: lane l ends with statistic l),
        // warps -> shared memory, then one double atomic per statistic and block
        float v[32];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            v[c] = s[c];
            v[16 + c] = ss[c];
        }
#pragma unroll
        for (int off = 16, nn = 16; off >= 1; off >>= 1, nn >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < nn; ++i) {
                const float send = upper ? v[i] : v[i + nn];
                const float keep = upper ? v[i + nn] : v[i];
                v[i] = keep + __shfl_xor_sync(full, send, off);
            }
        }
        red[warp][lane] = v[0];
        __syncthreads();
        if (warp == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[w][lane];
            atomicAdd(st_b + (PHASE == 1 ? 0 : 32) + lane, double(t));
        }
    }
}

// feat[row, m*cin + c] = sum_j mat[row,j,m] * x[b*n_in + ids[row,j], c];  16 support points per block
constexpr int kFeatPts = 16;
template <bool VEC>
__global__ void __launch_bounds__(256) fka_feat_kernel(const float* __restrict__ x, const int32_t* __restrict__ ids,
                                                       const float* __restrict__ mat, int n_in, int n_s, int kn, int cin,
                                                       float* __restrict__ feat) {
    __shared__ float smat[kFeatPts][kNbr][16];
    __shared__ int sid[kFeatPts][kNbr];
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * kFeatPts;
    const int tid = threadIdx.x;
    const int npts = min(kFeatPts, n_s - n0);
    const size_t row0 = (size_t)b * n_s + n0;
    for (int e = tid; e < npts * kn * 16; e += 256) smat[e / (kn * 16)][(e / 16) % kn][e % 16] = mat[row0 * kn * 16 + e];
    for (int e = tid; e < npts * kn; e += 256) sid[e / kn][e % kn] = ids[row0 * kn + e];
    __syncthreads();
    const int cw = VEC ? cin / 4 : cin;  // work items per point
    for (int e = tid; e < npts * cw; e += 256) {
        int pl = e / cw, cq = e % cw;
        if (VEC) {
            float4 acc[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) acc[m] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int j = 0; j < kn; ++j) {
                float4 xv = reinterpret_cast<const float4*>(x + ((size_t)b * n_in + sid[pl][j]) * cin)[cq];
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    float wgt = smat[pl][j][m];
                    acc[m].x = fmaf(wgt, xv.x, acc[m].x);
                    acc[m].y = fmaf(wgt, xv.y, acc[m].y);
                    acc[m].z = fmaf(wgt, xv.z, acc[m].z);
                    acc[m].w = fmaf(wgt, xv.w, acc[m].w);
                }
            }
            float* dst = feat + (row0 + pl) * (size_t)(16 * cin);
#pragma unroll
            for (int m = 0; m < 16; ++m) reinterpret_cast<float4*>(dst + (size_t)m * cin)[cq] = acc[m];
        } else {
            float acc[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) acc[m] = 0.f;
            for (int j = 0; j < kn; ++j) {
                float xv = x[((size_t)b * n_in + sid[pl][j]) * cin + cq];
#pragma unroll
                for (int m = 0; m < 16; ++m) acc[m] = fmaf(smat[pl][j][m], xv, acc[m]);
            }
            float* dst = feat + (row0 + pl) * (size_t)(16 * cin);
#pragma unroll
            for (int m = 0; m < 16; ++m) dst[(size_t)m * cin + cq] = acc[m];
        }
    }
}

__global__ void gather_max_kernel(const float* __restrict__ x, const int32_t* __restrict__ ids, int n_in, int n_s, int c,
                                  int kn, float* __restrict__ out) {
    const int b = blockIdx.y;
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n_s * c) return;
    int n = int(e / c), ch = int(e % c);
    size_t row = (size_t)b * n_s + n;
    float m = -INFINITY;
    for (int j = 0; j < kn; ++j) m = fmaxf(m, x[((size_t)b * n_in + ids[row * kn + j]) * c + ch]);
    out[row * c + ch] = m;
}

__global__ void __launch_bounds__(256) global_max_kernel(const float* __restrict__ x, int n, int c, float* __restrict__ out) {
    __shared__ float red[8][33];
    const int b = blockIdx.y;
    const int ch = blockIdx.x * 32 + (threadIdx.x & 31);
    const int r = threadIdx.x >> 5;
    float m = -INFINITY;
    if (ch < c)
        for (int i = r; i < n; i += 8) m = fmaxf(m, x[((size_t)b * n + i) * c + ch]);
    red[r][threadIdx.x & 31] = m;
    __syncthreads();
    if (r == 0 && ch < c) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i][threadIdx.x & 31]);
        out[(size_t)b * c + ch] = m;
    }
}

__global__ void latent_accumulate_kernel(const float* __restrict__ partial, const int32_t* __restrict__ ids, long long n, int c,
                                         float* latent, float* counts) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * c) return;
    long long i = e / c;
    int ch = int(e % c);
    int dst = ids[i];
    latent[(size_t)dst * c + ch] += partial[e];
    if (ch == 0) counts[dst] += 1.f;
}

// same with a row selection: partial row rows[i] goes to point ids[i] (the caller removed duplicate ids of the pass)
__global__ void latent_accumulate_rows_kernel(const float* __restrict__ partial, const int32_t* __restrict__ rows,
                                              const int32_t* __restrict__ ids, long long n, int c, float* latent, float* counts) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * (c / 4)) return;
    const long long i = e / (c / 4);
    const int c4 = int(e % (c / 4));
    const int dst = ids[i];
    const float4 v = reinterpret_cast<const float4*>(partial + (size_t)rows[i] * c)[c4];
    float4* d = reinterpret_cast<float4*>(latent + (size_t)dst * c) + c4;
    float4 o = *d;
    o.x += v.x;
    o.y += v.y;
    o.z += v.z;
    o.w += v.w;
    *d = o;
    if (c4 == 0) counts[dst] += 1.f;
}

__global__ void latent_finalize_kernel(float* latent, const float* __restrict__ counts, long long n, int c) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * c) return;
    latent[e] = latent[e] / counts[e / c];
}

static bool g_fka_fused = true;  // pps_debug_fka_fused(0) forces the unfused fp32 kernels (parity tests compare the two)

struct FkaLayout {
    size_t stats, mat, feat, total;
};
static FkaLayout fka_layout(int64_t b, int64_t n_s, int cin) {  // sized for kn = 16
    FkaLayout l;
    size_t off = 0;
    l.stats = off;
    off = align_up(off + (size_t)b * 64 * sizeof(double), 256);
    l.mat = off;
    off = align_up(off + (size_t)b * n_s * kNbr * 16 * sizeof(float), 256);
    l.feat = off;
    off = align_up(off + (size_t)b * n_s * 16 * (size_t)cin * sizeof(float), 256);
    l.total = off;
    return l;
}

}  // namespace pps

using namespace pps;

extern "C" {

void pps_debug_fka_fused(int on) { g_fka_fused = on != 0; }

size_t pps_fkaconv_workspace_bytes(int64_t b, int64_t n_s, int cin) {
    if (b <= 0 || n_s <= 0 || cin <= 0) return 0;
    return fka_layout(b, n_s, cin).total;
}

size_t pps_fkaconv_workspace_bytes_for(const pps_fkaconv_weights* w, int kn, int64_t b, int64_t n_s) {
    if (!w || b <= 0 || n_s <= 0) return 0;
    if (g_fka_fused && fka_fused_supported(w, kn, n_s)) return align_up((size_t)b * 64 * sizeof(double), 256);  // statistics only
    return fka_layout(b, n_s, w->cin).total;
}

int pps_fkaconv_forward(const pps_fkaconv_weights* w, const float* x, const float* pts, const float* support,
                        const int32_t* ids, int kn, int64_t b, int64_t n_in, int64_t n_s, void* workspace,
                        size_t workspace_bytes, float* out, void* stream) {
    PPS_CHECK_ARG(w && x && pts && support && ids && workspace && out, "pps_fkaconv_forward: null pointer");
    PPS_CHECK_ARG(kn >= 1 && kn <= kNbr, "pps_fkaconv_forward: %d neighbours per support point, supported 1..16", kn);
    PPS_CHECK_ARG(b >= 1 && b <= 65535 && n_in >= 1 && n_s >= 1 && n_in < (1ll << 31) && n_s < (1ll << 31),
                  "pps_fkaconv_forward: bad sizes b=%lld n_in=%lld n_s=%lld", (long long)b, (long long)n_in, (long long)n_s);
    PPS_CHECK_ARG(w->cin >= 1 && w->cout >= 1 && (w->act == 0 || w->act == 1), "pps_fkaconv_forward: bad weights");
    FkaLayout l = fka_layout(b, n_s, w->cin);
    const bool fused = g_fka_fused && fka_fused_supported(w, kn, n_s);
    const size_t need = fused ? align_up((size_t)b * 64 * sizeof(double), 256) : l.total;
    if (workspace_bytes < need) {
        set_error("pps_fkaconv_forward: workspace %zu < required %zu", workspace_bytes, need);
        return PPS_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(workspace);
    double* stats = reinterpret_cast<double*>(base + l.stats);
    float* mat = reinterpret_cast<float*>(base + l.mat);
    float* feat = reinterpret_cast<float*>(base + l.feat);
    if (fused) return fka_fused_impl(w, x, pts, support, ids, b, n_in, n_s, stats, out, st);
    PPS_CUDA(cudaMemsetAsync(stats, 0, (size_t)b * 64 * sizeof(double), st));
    FkaParams prm{w->alpha, w->beta, 1.f / w->norm_radius, w->act};
    dim3 grid((unsigned)ceil_div(n_s, kFkaPts), (unsigned)b);
    fka_weight_kernel<1><<<grid, 256, 0, st>>>(pts, support, ids, (int)n_in, (int)n_s, kn, prm, w->fc1, w->fc2, w->fc3, w->in1_w,
                                               w->in1_b, w->in2_w, w->in2_b, stats, mat);
    PPS_LAUNCH_CHECK();
    fka_weight_kernel<2><<<grid, 256, 0, st>>>(pts, support, ids, (int)n_in, (int)n_s, kn, prm, w->fc1, w->fc2, w->fc3, w->in1_w,
                                               w->in1_b, w->in2_w, w->in2_b, stats, mat);
    PPS_LAUNCH_CHECK();
    fka_weight_kernel<3><<<grid, 256, 0, st>>>(pts, support, ids, (int)n_in, (int)n_s, kn, prm, w->fc1, w->fc2, w->fc3, w->in1_w,
                                               w->in1_b, w->in2_w, w->in2_b, stats, mat);
    PPS_LAUNCH_CHECK();
    dim3 fgrid((unsigned)ceil_div(n_s, kFeatPts), (unsigned)b);
    bool vec = (w->cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    if (vec)
        fka_feat_kernel<true><<<fgrid, 256, 0, st>>>(x, ids, mat, (int)n_in, (int)n_s, kn, w->cin, feat);
    else
        fka_feat_kernel<false><<<fgrid, 256, 0, st>>>(x, ids, mat, (int)n_in, (int)n_s, kn, w->cin, feat);
    PPS_LAUNCH_CHECK();
    int kdim = 16 * w->cin;
    return linear_impl(feat, w->cv_w, w->out_bias, nullptr, nullptr, out, b * n_s, w->cout, kdim, kdim, w->cout,
                       w->out_relu ? 1 : 0, st);
}

int pps_gather_max(const float* x, const int32_t* ids, int64_t b, int64_t n_in, int64_t n_s, int c, int kn, float* out,
                   void* stream) {
    PPS_CHECK_ARG(x && ids && out && b >= 1 && b <= 65535 && c >= 1 && kn >= 1, "pps_gather_max: bad arguments");
    dim3 grid((unsigned)ceil_div(n_s * c, 256), (unsigned)b);
    gather_max_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ids, (int)n_in, (int)n_s, c, kn, out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_global_max(const float* x, int64_t b, int64_t n, int c, float* out, void* stream) {
    PPS_CHECK_ARG(x && out && b >= 1 && b <= 65535 && n >= 1 && c >= 1, "pps_global_max: bad arguments");
    dim3 grid((unsigned)ceil_div(c, 32), (unsigned)b);
    global_max_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, (int)n, c, out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_latent_accumulate(const float* partial, const int32_t* ids, int64_t n, int c, float* latent, float* counts,
                          void* stream) {
    PPS_CHECK_ARG(partial && ids && latent && counts && n >= 0 && c >= 1, "pps_latent_accumulate: bad arguments");
    if (n == 0) return PPS_OK;
    latent_accumulate_kernel<<<(unsigned)ceil_div(n * c, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(partial, ids, n, c,
                                                                                                           latent, counts);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_latent_accumulate_rows(const float* partial, const int32_t* rows, const int32_t* ids, int64_t n, int c, float* latent,
                               float* counts, void* stream) {
    PPS_CHECK_ARG(partial && rows && ids && latent && counts && n >= 0 && c >= 4 && c % 4 == 0,
                  "pps_latent_accumulate_rows: bad arguments (c must be a multiple of 4)");
    if (n == 0) return PPS_OK;
    latent_accumulate_rows_kernel<<<(unsigned)ceil_div(n * (c / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        partial, rows, ids, n, c, latent, counts);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

int pps_latent_finalize(float* latent, const float* counts, int64_t n, int c, void* stream) {
    PPS_CHECK_ARG(latent && counts && n >= 0 && c >= 1, "pps_latent_finalize: bad arguments");
    if (n == 0) return PPS_OK;
    latent_finalize_kernel<<<(unsigned)ceil_div(n * c, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(latent, counts, n, c);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}
}
