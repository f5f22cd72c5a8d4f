// Training step (BASELINE config 5): ONE strided, batched GEMM entry point for every dense contraction of the forward and the
// backward pass  --  forward  Y = X.W^T,  data gradient  dX = dY.W,  weight gradient  dW = dY^T.X (split over the rows),  the
// per-query 64x64 feature transform of the PointNet (batched) and its two gradients.  The operands are addressed by element
// strides, so no transposed copy of an activation or a weight is ever made.
//     C_b[m,n] (+)= sum_k A_b[m,k] . B_b[k,n] (+ bias[n]),   A_b[m,k] = a[b*sa_b + m*sa_m + k*sa_k],  B_b[k,n] = b[b*sb_b + k*sb_k + n*sb_n]
// precision 0: fp32 SIMT (this file; the path of the gradient parity tests), precision 1: bf16 tcgen05 (train_gemm_tc.cu).
// Replaces torch.nn.functional.linear / conv1d / bmm and their autograd formulas in the reference's training step
// (source/poco_model.py:120-125 -> Lightning backward).
#include <algorithm>

#include "common.cuh"

namespace pps {

int gemm_tc_impl(const float* a, int64_t sa_b, int64_t sa_m, int64_t sa_k, const float* b, int64_t sb_b, int64_t sb_k, int64_t sb_n,
                 float* c, int64_t sc_b, int64_t ldc, int64_t batch, int64_t m, int n, int64_t k, const float* bias, int accumulate,
                 cudaStream_t st);
bool gemm_tc_supported(int64_t m, int n, int64_t k);

namespace train {

constexpr int kTM = 64, kTN = 64, kTK = 16, kThreads = 256;

// one 64x64 output tile per block (blockIdx.x = m tile, y = n tile, z = batch * splits + split); 4x4 outputs per thread.
// ATOMIC: split-K partial sums or accumulation into an existing C with red.global.add.f32
template <bool ATOMIC>
__global__ void __launch_bounds__(kThreads) gemm_strided_kernel(const float* __restrict__ a, long long sa_b, long long sa_m, long long sa_k,
                                                                const float* __restrict__ b, long long sb_b, long long sb_k, long long sb_n,
                                                                float* c, long long sc_b, long long ldc, long long m, int n, long long k,
                                                                const float* __restrict__ bias, int splits, long long kslice) {
    __shared__ float As[kTK][kTM + 4];
    __shared__ float Bs[kTK][kTN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long batch = blockIdx.z / splits;
    const int split = blockIdx.z % splits;
    const long long m0 = (long long)blockIdx.x * kTM;
    const int n0 = blockIdx.y * kTN;
    a += batch * sa_b;
    b += batch * sb_b;
    c += batch * sc_b;
    const long long kbeg = split * kslice;
    const long long kend = min(k, kbeg + kslice);
    // loader mapping: the unit-stride dimension of each operand runs across consecutive threads
    const bool a_kfast = sa_k == 1, b_kfast = sb_k == 1;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (long long k0 = kbeg; k0 < kend; k0 += kTK) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int e = tid + p * kThreads;  // 1024 elements per tile
            int kk, mm;
            if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
            const long long gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < m && gk < kend) ? a[gm * sa_m + gk * sa_k] : 0.f;
            int kb, nn;
            if (b_kfast) { kb = e & 15; nn = e >> 4; } else { nn = e & 63; kb = e >> 6; }
            const long long gkb = k0 + kb;
            const int gn = n0 + nn;
            Bs[kb][nn] = (gn < n && gkb < kend) ? b[gkb * sb_k + (long long)gn * sb_n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kTK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long gm = m0 + ty * 4 + i;
        if (gm >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= n) continue;
            float v = acc[i][j];
            if (bias && split == 0) v += bias[gn];
            if (ATOMIC)
                atomicAdd(c + gm * ldc + gn, v);
            else
                c[gm * ldc + gn] = v;
        }
    }
}

}  // namespace train

int gemm_impl(const float* a, int64_t sa_b, int64_t sa_m, int64_t sa_k, const float* b, int64_t sb_b, int64_t sb_k, int64_t sb_n, float* c,
              int64_t sc_b, int64_t ldc, int64_t batch, int64_t m, int n, int64_t k, const float* bias, int accumulate, int precision,
              cudaStream_t st) {
    using namespace train;
    PPS_CHECK_ARG(a && b && c, "pps_gemm: null pointer");
    PPS_CHECK_ARG(batch >= 0 && m >= 0 && n > 0 && k >= 0 && ldc >= n, "pps_gemm: bad shape batch=%lld m=%lld n=%d k=%lld ldc=%lld",
                  (long long)batch, (long long)m, n, (long long)k, (long long)ldc);
    PPS_CHECK_ARG(precision == 0 || precision == 1, "pps_gemm: precision must be 0 (fp32) or 1 (bf16 tensor cores)");
    if (batch == 0 || m == 0) return PPS_OK;
    if (precision == 1 && batch == 1 && gemm_tc_supported(m, n, k))
        return gemm_tc_impl(a, sa_b, sa_m, sa_k, b, sb_b, sb_k, sb_n, c, sc_b, ldc, batch, m, n, k, bias, accumulate, st);
    const long long tiles = ceil_div(m, kTM) * ceil_div(n, kTN) * batch;
    // few output tiles and a long reduction (weight gradients: k = number of rows): split k over blockIdx.z
    int splits = 1;
    if (tiles < 2 * kNumSMs && k >= 512) splits = (int)std::min<long long>(ceil_div(4 * kNumSMs, tiles), ceil_div(k, 128));
    const long long kslice = align_up(ceil_div(std::max<int64_t>(k, 1), splits), kTK);
    splits = (int)std::max<int64_t>(1, ceil_div(k, kslice));
    const bool atomic = splits > 1 || accumulate;
    const int64_t per_launch = std::max<int64_t>(1, 65535 / splits);  // grid.z limit: batches go in slices
    for (int64_t b0 = 0; b0 < batch; b0 += per_launch) {
        const int64_t nb = std::min<int64_t>(per_launch, batch - b0);
        const float* ab = a + b0 * sa_b;
        const float* bb = b + b0 * sb_b;
        float* cb = c + b0 * sc_b;
        if (atomic && !accumulate) {
            if (sc_b == m * ldc && ldc == n) {
                PPS_CUDA(cudaMemsetAsync(cb, 0, (size_t)nb * m * n * sizeof(float), st));
            } else {
                for (int64_t bi = 0; bi < nb; ++bi)
                    PPS_CUDA(cudaMemset2DAsync(cb + bi * sc_b, (size_t)ldc * sizeof(float), 0, (size_t)n * sizeof(float), (size_t)m, st));
            }
        }
        dim3 grid((unsigned)ceil_div(m, kTM), (unsigned)ceil_div(n, kTN), (unsigned)(nb * splits));
        if (atomic)
            gemm_strided_kernel<true><<<grid, kThreads, 0, st>>>(ab, sa_b, sa_m, sa_k, bb, sb_b, sb_k, sb_n, cb, sc_b, ldc, m, n, k, bias, splits, kslice);
        else
            gemm_strided_kernel<false><<<grid, kThreads, 0, st>>>(ab, sa_b, sa_m, sa_k, bb, sb_b, sb_k, sb_n, cb, sc_b, ldc, m, n, k, bias, splits, kslice);
        PPS_LAUNCH_CHECK();
    }
    return PPS_OK;
}

}  // namespace pps

extern "C" int pps_gemm(const float* a, int64_t sa_b, int64_t sa_m, int64_t sa_k, const float* b, int64_t sb_b, int64_t sb_k, int64_t sb_n,
                        float* c, int64_t sc_b, int64_t ldc, int64_t batch, int64_t m, int n, int64_t k, const float* bias, int accumulate,
                        int precision, void* stream) {
    return pps::gemm_impl(a, sa_b, sa_m, sa_k, b, sb_b, sb_k, sb_n, c, sc_b, ldc, batch, m, n, k, bias, accumulate, precision,
                          static_cast<cudaStream_t>(stream));
}
