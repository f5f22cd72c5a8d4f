// Generic fused pointwise layer in fp32 SIMT:  Y = act(X[gather] . W^T + bias + residual).
// Serves every 1x1 conv / linear (+ folded eval BatchNorm, ReLU, residual add, 1-NN up-sampling gather) of the
// encoder, the STN/PointNet and the MLP (source/base/nn.py:162-190,305-373,415-417,438-450,530-548), and is the
// fp32 reference path of the decoder GEMMs.  Classic 128xBN x16 shared-memory tiling, 8xTN register blocking.
#include <algorithm>

#include "common.cuh"

namespace pps {

constexpr int kBM = 128, kBK = 16, kThreads = 256;

template <int BN, bool VEC, bool SPLIT>
__global__ void __launch_bounds__(kThreads) linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* residual,
                                                          const int32_t* __restrict__ gather, float* y, long long m, int n,
                                                          int k, int ldx, int ldy, int act, int kslice) {
    constexpr int TN = BN / 16;     // columns per thread: 8 (BN=128) or 4 (BN=64)
    constexpr int NB = TN / 4;      // number of 4-wide column blocks per thread
    __shared__ float As[kBK][kBM + 4];
    __shared__ float Bs[kBK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * kBM;
    const int col0 = blockIdx.y * BN;

    // loader mapping: one float4 (4 consecutive k) of one row per thread and pass
    const int lrow = tid & 127, lkq = tid >> 7;  // A: 128 rows x 4 k-quads -> 2 passes
    long long arow[2];
    bool arow_ok[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        long long r = row0 + lrow;
        arow_ok[p] = r < m;
        long long src = arow_ok[p] ? (gather ? (long long)gather[r] : r) : 0;
        arow[p] = src * ldx;
    }

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // SPLIT: this block reduces only k in [blockIdx.z * kslice, ...) and adds its partial sums atomically into a zeroed y
    const int kbeg = SPLIT ? blockIdx.z * kslice : 0;
    const int kend = SPLIT ? min(k, kbeg + kslice) : k;
    for (int k0 = kbeg; k0 < kend; k0 += kBK) {
        // ---- load A tile (transposed into As[k][row])
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            int kq = lkq + 2 * p;
            int kk = k0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (arow_ok[p]) {
                if (VEC) {
                    if (kk < kend) v = *reinterpret_cast<const float4*>(x + arow[p] + kk);
                } else {
                    if (kk + 0 < kend) v.x = x[arow[p] + kk + 0];
                    if (kk + 1 < kend) v.y = x[arow[p] + kk + 1];
                    if (kk + 2 < kend) v.z = x[arow[p] + kk + 2];
                    if (kk + 3 < kend) v.w = x[arow[p] + kk + 3];
                }
            }
            As[kq * 4 + 0][lrow] = v.x;
            As[kq * 4 + 1][lrow] = v.y;
            As[kq * 4 + 2][lrow] = v.z;
            As[kq * 4 + 3][lrow] = v.w;
        }
        // ---- load W tile (Bs[k][col])
        constexpr int WPASS = BN * 4 / kThreads;  // BN rows x 4 k-quads
#pragma unroll
        for (int p = 0; p < WPASS; ++p) {
            int e = tid + p * kThreads;
            int wrow = e % BN, kq = e / BN;
            int kk = k0 + kq * 4;
            int c = col0 + wrow;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < n) {
                const float* wp = w + (size_t)c * k;
                if (VEC) {
                    if (kk < kend) v = *reinterpret_cast<const float4*>(wp + kk);
                } else {
                    if (kk + 0 < kend) v.x = wp[kk + 0];
                    if (kk + 1 < kend) v.y = wp[kk + 1];
                    if (kk + 2 < kend) v.z = wp[kk + 2];
                    if (kk + 3 < kend) v.w = wp[kk + 3];
                }
            }
            Bs[kq * 4 + 0][wrow] = v.x;
            Bs[kq * 4 + 1][wrow] = v.y;
            Bs[kq * 4 + 2][wrow] = v.z;
            Bs[kq * 4 + 3][wrow] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            float a[8], b[TN];
            float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][nb * 64 + tx * 4]);
                b[nb * 4 + 0] = bv.x; b[nb * 4 + 1] = bv.y; b[nb * 4 + 2] = bv.z; b[nb * 4 + 3] = bv.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long r = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= m) continue;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int c = col0 + nb * 64 + tx * 4 + j;
                if (c >= n) continue;
                float v = acc[i][nb * 4 + j];
                if (SPLIT) {
                    atomicAdd(y + r * ldy + c, v);
                    continue;
                }
                if (bias) v += bias[c];
                if (residual) v += residual[r * ldy + c];
                if (act == 1) v = fmaxf(v, 0.f);
                y[r * ldy + c] = v;
            }
        }
    }
}

// y = act(y + bias + residual) after a split-K accumulation
__global__ void linear_epilogue_kernel(float* y, const float* __restrict__ bias, const float* residual, long long m, int n, int ldy,
                                       int act) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m * n) return;
    long long r = e / n;
    int c = int(e % n);
    float v = y[r * ldy + c];
    if (bias) v += bias[c];
    if (residual) v += residual[r * ldy + c];
    if (act == 1) v = fmaxf(v, 0.f);
    y[r * ldy + c] = v;
}

template <int BN>
static void launch_linear(dim3 grid, bool vec, bool split, const float* x, const float* w, const float* bias, const float* residual,
                          const int32_t* gather, float* y, int64_t m, int n, int k, int ldx, int ldy, int act, int kslice,
                          cudaStream_t st) {
    dim3 block(kThreads);
    if (split) {
        if (vec)
            linear_kernel<BN, true, true><<<grid, block, 0, st>>>(x, w, bias, residual, gather, y, m, n, k, ldx, ldy, act, kslice);
        else
            linear_kernel<BN, false, true><<<grid, block, 0, st>>>(x, w, bias, residual, gather, y, m, n, k, ldx, ldy, act, kslice);
    } else {
        if (vec)
            linear_kernel<BN, true, false><<<grid, block, 0, st>>>(x, w, bias, residual, gather, y, m, n, k, ldx, ldy, act, kslice);
        else
            linear_kernel<BN, false, false><<<grid, block, 0, st>>>(x, w, bias, residual, gather, y, m, n, k, ldx, ldy, act, kslice);
    }
}

int linear_impl(const float* x, const float* w, const float* bias, const float* residual, const int32_t* gather,
                float* y, int64_t m, int n, int k, int ldx, int ldy, int act, cudaStream_t st) {
    PPS_CHECK_ARG(x && w && y, "pps_linear: null pointer");
    PPS_CHECK_ARG(m >= 0 && n > 0 && k > 0 && ldx >= k && ldy >= n, "pps_linear: bad shape m=%lld n=%d k=%d ldx=%d ldy=%d",
                  (long long)m, n, k, ldx, ldy);
    if (m == 0) return PPS_OK;
    bool vec = (k % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
               ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
    const int bn = n > 64 ? 128 : 64;
    dim3 grid((unsigned)ceil_div(m, kBM), (unsigned)ceil_div(n, bn));
    // skinny problems with a long reduction (the FKAConv contractions of the deep encoder levels: M = 39..625 rows,
    // K = 2048..8192) would occupy a handful of SMs: split K across blockIdx.z, accumulate with float atomics into a
    // zeroed y and apply bias / residual / activation in a second pass.  residual may alias y only without the split.
    const long long blocks = (long long)grid.x * grid.y;
    int slices = 1;
    if (blocks < 64 && k >= 1024 && residual != y) {
        slices = (int)std::min<long long>(ceil_div(2 * kNumSMs, blocks), k / 256);
    }
    if (slices > 1) {
        int kslice = (int)align_up(ceil_div(k, slices), kBK);
        grid.z = (unsigned)ceil_div(k, kslice);
        PPS_CUDA(cudaMemset2DAsync(y, (size_t)ldy * sizeof(float), 0, (size_t)n * sizeof(float), (size_t)m, st));
        if (bn == 128)
            launch_linear<128>(grid, vec, true, x, w, bias, residual, gather, y, m, n, k, ldx, ldy, act, kslice, st);
        else
            launch_linear<64>(grid, vec, true, x, w, bias, residual, gather, y, m, n, k, ldx, ldy, act, kslice, st);
        PPS_LAUNCH_CHECK();
        if (bias || residual || act)
            linear_epilogue_kernel<<<(unsigned)ceil_div(m * n, 256), 256, 0, st>>>(y, bias, residual, m, n, ldy, act);
        PPS_LAUNCH_CHECK();
        return PPS_OK;
    }
    if (bn == 128)
        launch_linear<128>(grid, vec, false, x, w, bias, residual, gather, y, m, n, k, ldx, ldy, act, k, st);
    else
        launch_linear<64>(grid, vec, false, x, w, bias, residual, gather, y, m, n, k, ldx, ldy, act, k, st);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

}  // namespace pps

extern "C" int pps_linear(const float* x, const float* w, const float* bias, const float* residual,
                          const int32_t* gather, float* y, int64_t m, int n, int k, int ldx, int ldy, int act,
                          void* stream) {
    return pps::linear_impl(x, w, bias, residual, gather, y, m, n, k, ldx, ldy, act, static_cast<cudaStream_t>(stream));
}
