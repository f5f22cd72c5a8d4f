// bf16 tensor-core path of pps_gemm (training step, BASELINE config 5: "fit, bf16"): the same strided contraction as
// train_gemm.cu on tcgen05.  fp32 master activations / weights / gradients stay in HBM; a CTA converts its 128 x 64 operand
// tiles to bf16 on the way into shared memory (the operands of one GEMM are read from HBM once, never re-materialised in bf16),
// the accumulator is fp32 in TMEM.  Because the tiles are written by threads, any element strides work: X.W^T, dY.W and
// dY^T.X (k = rows, split over the grid with red.global.add) all run through this kernel without a transposed copy.
//
// CTA = 256 threads, output tile 128 x BN (BN = 32 / 64 / 128 / 256, runtime: the whole n of the 256-wide layers sits in ONE tile, so
// the activation operand is read from HBM once), K step 64, two shared-memory stages so that two or three CTAs share an SM: the
// loads of stage t+1 are in flight while the tensor pipe works on stage t, and the other CTAs cover this one's barrier waits.
// Operand layout: UMMA canonical K-major without swizzle, [k8 block][row][8 x bf16], block pitch rows * 16 + 16 bytes
// (conflict-free for both loader mappings).  Epilogue: TMEM -> registers -> a padded per-warp shared-memory transpose -> 128-byte
// coalesced row segments (a thread-per-row store pattern costs 32 LSU wavefronts per instruction and ran the kernel at a fifth of
// the HBM rate).
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace pps {
namespace tc {
namespace gemm {

constexpr int kThreads = 256;
constexpr int kBM = 128, kBK = 64;
constexpr int kLbo = 128 * 16 + 16;        // 2064: pitch of a k8 block of the A tile
constexpr int kTile = 8 * kLbo;            // 16512 B: the A tile (128 rows x 64 k, bf16)
constexpr int kStages = 2;
constexpr int kOffB = kStages * kTile;
__host__ __device__ constexpr int lbo_b(int bn) { return bn * 16 + 16; }
__host__ __device__ constexpr int tile_b(int bn) { return 8 * lbo_b(bn); }
__host__ __device__ constexpr int off_bar(int bn) { return kOffB + kStages * tile_b(bn); }  // stage_free[2], accum_done, tmem slot
__host__ __device__ constexpr int smem_bytes(int bn) { return off_bar(bn) + 48; }
constexpr int kStagePitch = 36;            // floats per row of the epilogue's per-warp 32 x 32 transpose buffer
static_assert(8 * 32 * kStagePitch * 4 <= off_bar(32), "the epilogue staging reuses the operand stages (all MMAs are complete)");

// instruction descriptor kind::f16: D fp32, A / B bf16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

struct Operand {
    const float* p;
    long long s_row, s_k;  // element strides of the tile's row dimension (m or n) and of k
    long long rows;        // valid rows (m or n)
};

// 4 chunks (8 consecutive k of one row) per thread: global fp32 -> registers
template <bool KFAST>
__device__ __forceinline__ void load_chunks(const Operand& o, long long row0, long long k0, long long kend, int nrows, float (&v)[4][8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int chunk = threadIdx.x + i * kThreads;
        const int row = KFAST ? (chunk >> 3) : (chunk & 127);
        const int kb = KFAST ? (chunk & 7) : (chunk >> 7);
        const long long gr = row0 + row, gk = k0 + kb * 8;
        const bool row_ok = row < nrows && gr < o.rows;
        const float* src = o.p + gr * o.s_row + gk * o.s_k;
        if (KFAST && row_ok && gk + 8 <= kend && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const float4 u0 = reinterpret_cast<const float4*>(src)[0], u1 = reinterpret_cast<const float4*>(src)[1];
            v[i][0] = u0.x; v[i][1] = u0.y; v[i][2] = u0.z; v[i][3] = u0.w;
            v[i][4] = u1.x; v[i][5] = u1.y; v[i][6] = u1.z; v[i][7] = u1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] = (row_ok && gk + j < kend) ? src[j * o.s_k] : 0.f;
        }
    }
}

template <bool KFAST>
__device__ __forceinline__ void store_chunks(uint8_t* tile, int lbo, int nrows, const float (&v)[4][8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int chunk = threadIdx.x + i * kThreads;
        const int row = KFAST ? (chunk >> 3) : (chunk & 127);
        const int kb = KFAST ? (chunk & 7) : (chunk >> 7);
        if (row >= nrows) continue;  // the B tile holds bn rows only
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[i][2 * j], v[i][2 * j + 1]);
            w[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(tile + kb * lbo + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <bool AK, bool BK>
__global__ void __launch_bounds__(kThreads) gemm_bf16_kernel(Operand a, Operand b, float* __restrict__ c, long long ldc, int n, long long k,
                                                             int bn, const float* __restrict__ bias, int splits, long long kslice,
                                                             int atomic) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = (long long)blockIdx.x * kBM;
    const int n0 = blockIdx.y * bn;
    const int split = blockIdx.z;
    const long long kbeg = split * kslice, kend = min(k, kbeg + kslice);
    const int nk = (int)((kend - kbeg + kBK - 1) / kBK);
    const int lbo = lbo_b(bn), tileb = tile_b(bn), obar = off_bar(bn);
    const uint32_t tmem_cols = bn <= 32 ? 32u : (bn <= 64 ? 64u : (bn <= 128 ? 128u : 256u));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + obar + 32);
    if (tid == 0) {
        mbar_init(sbase + obar, 1);
        mbar_init(sbase + obar + 8, 1);
        mbar_init(sbase + obar + 16, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + obar + 32), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = idesc_bf16(bn);
    const int bhalves = (bn + 127) >> 7;  // the B tile is loaded 128 rows at a time (register budget)
    float va[4][8], vb[4][8];
    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt & 1;
        const long long k0 = kbeg + (long long)kt * kBK;
        uint8_t* bt = smem + kOffB + s * tileb;
        load_chunks<AK>(a, m0, k0, kend, kBM, va);
        load_chunks<BK>(b, n0, k0, kend, min(bn, 128), vb);
        if (kt >= kStages) mbar_wait(sbase + obar + 8 * s, ((kt >> 1) - 1) & 1);  // the MMAs that read this stage are done
        store_chunks<AK>(smem + s * kTile, kLbo, kBM, va);
        store_chunks<BK>(bt, lbo, min(bn, 128), vb);
        if (bhalves > 1) {
            load_chunks<BK>(b, n0 + 128, k0, kend, bn - 128, vb);
            store_chunks<BK>(bt + 128 * 16, lbo, bn - 128, vb);
        }
        fence_async_smem();
        __syncthreads();
        if (warp == 0) {  // the whole warp (warp-uniform control flow), one elected lane issues: see elect_one in tc_common.cuh
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint64_t ad = umma_desc(sbase + s * kTile + 2 * q * kLbo, kLbo, 128);
                    const uint64_t bd = umma_desc(sbase + kOffB + s * tileb + 2 * q * lbo, lbo, 128);
                    umma(tmem, ad, bd, idesc, (kt > 0 || q > 0) ? 1u : 0u);
                }
                tc_commit(sbase + obar + 8 * s);
                if (kt == nk - 1) tc_commit(sbase + obar + 16);
            }
            __syncwarp();
        }
    }
    if (nk > 0) {
        mbar_wait(sbase + obar + 16, 0);
        tc_fence_after();
    }
    // epilogue: warp w owns lanes 32 * (w % 4) .. +31 (output rows) and the 32-column blocks cb = w / 4, w / 4 + 2, ...; a block goes
    // TMEM -> registers (thread = row) -> this warp's padded transpose buffer -> global memory with lanes along the columns
    const int lane_grp = warp & 3;
    float* stage = reinterpret_cast<float*>(smem) + warp * 32 * kStagePitch;  // all MMAs are done: the operand stages are free
    const long long row0 = m0 + lane_grp * 32;
    const bool vec_ok = !atomic && (ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(c) & 15) == 0) && ((n0 & 3) == 0);
    for (int cb = warp >> 2; cb * 32 < bn; cb += 2) {
        const int col0 = cb * 32;
        if (n0 + col0 >= n) break;  // warp-uniform
        float v[32];
        if (nk > 0) {
            tmem_ld32(tmem + ((uint32_t)(lane_grp * 32) << 16) + col0, v);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stage + lane * kStagePitch + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        if (vec_ok && n0 + col0 + 32 <= n) {
            // 8 lanes x float4 cover the 32 columns of a row, 4 rows per instruction: four full 128-byte lines
            const int r4 = lane >> 3, c4 = (lane & 7) * 4;
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias && split == 0) bv = *reinterpret_cast<const float4*>(bias + n0 + col0 + c4);
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const int r = rr * 4 + r4;
                if (row0 + r < a.rows) {
                    float4 x = *reinterpret_cast<const float4*>(stage + r * kStagePitch + c4);
                    x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
                    *reinterpret_cast<float4*>(c + (row0 + r) * ldc + n0 + col0 + c4) = x;
                }
            }
        } else {
            const int gn = n0 + col0 + lane;
            const float bvs = (bias && split == 0 && gn < n) ? bias[gn] : 0.f;
            for (int r = 0; r < 32; ++r) {
                if (row0 + r >= a.rows || gn >= n) continue;
                const float x = stage[r * kStagePitch + lane] + bvs;
                if (atomic)
                    atomicAdd(c + (row0 + r) * ldc + gn, x);
                else
                    c[(row0 + r) * ldc + gn] = x;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    }
}

}  // namespace gemm
}  // namespace tc

// every shape runs (rows / columns / k beyond the problem are zero-filled by the loaders); tiny problems are not worth a 128-row tile
bool gemm_tc_supported(int64_t m, int n, int64_t k) { return m * (int64_t)n * k >= 4096; }

int gemm_tc_impl(const float* a, int64_t sa_b, int64_t sa_m, int64_t sa_k, const float* b, int64_t sb_b, int64_t sb_k, int64_t sb_n,
                 float* c, int64_t sc_b, int64_t ldc, int64_t batch, int64_t m, int n, int64_t k, const float* bias, int accumulate,
                 cudaStream_t st) {
    using namespace tc::gemm;
    (void)sa_b; (void)sb_b; (void)sc_b;
    PPS_CHECK_ARG(batch == 1, "pps_gemm: the tensor-core path takes one problem per call");
    static unsigned char configured[kMaxDevices] = {};
    if (first_use_on_device(configured)) {
        PPS_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(256)));
        PPS_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(256)));
        PPS_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(256)));
        PPS_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(256)));
    }
    // BN: the smallest of 32 / 64 / 128 / 256 that covers n in one tile, else 256
    const int bn = n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256));
    const int smem = smem_bytes(bn);
    const long long tiles = ceil_div(m, kBM) * ceil_div(n, bn);
    int splits = 1;
    if (tiles < 2 * kNumSMs && k >= 1024) splits = (int)std::min<long long>(ceil_div(3 * kNumSMs, tiles), k / 256);
    long long kslice = (long long)align_up((size_t)ceil_div(k, splits), kBK);
    splits = (int)ceil_div(k, kslice);
    const int atomic = (splits > 1 || accumulate) ? 1 : 0;
    if (atomic && !accumulate) {
        if (ldc == n) {
            PPS_CUDA(cudaMemsetAsync(c, 0, (size_t)m * n * sizeof(float), st));
        } else {
            PPS_CUDA(cudaMemset2DAsync(c, (size_t)ldc * sizeof(float), 0, (size_t)n * sizeof(float), (size_t)m, st));
        }
    }
    Operand oa{a, sa_m, sa_k, m}, ob{b, sb_n, sb_k, n};
    dim3 grid((unsigned)ceil_div(m, kBM), (unsigned)ceil_div(n, bn), (unsigned)splits);
    const bool ak = sa_k == 1, bk = sb_k == 1;
    if (ak && bk)
        gemm_bf16_kernel<true, true><<<grid, kThreads, smem, st>>>(oa, ob, c, ldc, n, k, bn, bias, splits, kslice, atomic);
    else if (ak)
        gemm_bf16_kernel<true, false><<<grid, kThreads, smem, st>>>(oa, ob, c, ldc, n, k, bn, bias, splits, kslice, atomic);
    else if (bk)
        gemm_bf16_kernel<false, true><<<grid, kThreads, smem, st>>>(oa, ob, c, ldc, n, k, bn, bias, splits, kslice, atomic);
    else
        gemm_bf16_kernel<false, false><<<grid, kThreads, smem, st>>>(oa, ob, c, ldc, n, k, bn, bias, splits, kslice, atomic);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

}  // namespace pps
