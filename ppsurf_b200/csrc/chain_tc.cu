// Tensor-core kernels for the per-QUERY layer chains of the decoder (path 1), M = queries of the chunk, tiles of 128 rows:
//
//   stn_fc_tc_kernel   g [q,256] -> stn.fc1 (+bn4, ReLU) -> stn.fc2 (+bn5, ReLU) -> stn.fc3 (+I) = T [q,4096]
//                      (source/base/nn.py:183-189); fc3 runs as 16 column blocks of 256 with two TMEM accumulators so the
//                      store of block i overlaps the MMAs of block i+1
//   mlp_tc_kernel      feat = (W8 Wv) pooled_proj + (Wv_att A3) pooled_pn + biases  (both branches accumulate into ONE
//                      accumulator: the sum of the branches is linear, source/ppsurf_model.py:100) -> mlp.layers.0 ->
//                      mlp.layers.1 -> mlp.layers.2 -> softmax difference (source/base/nn.py:415-417, poco_utils.py:79-80)
//
// Same split-fp16 tcgen05 scheme, operand layout and warp roles as decode_tc.cu (producer warp, MMA warp, 8 epilogue
// warps); these chains are small (< one wave of tiles per chunk), so phases are simply serialised.
#include "tc_common.cuh"

namespace pps {
namespace tc {
namespace chain {

constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kLbo = 128 * 16 + 16;        // 2064
constexpr int kABytes = 32 * kLbo;         // operand tile for K <= 256
constexpr int kSlot = 16384;
constexpr int kStages = 4;
constexpr int kOffAhi = 0;
constexpr int kOffAlo = kABytes;           // 66048
constexpr int kOffRing = 2 * kABytes;      // 132096
constexpr int kOffPar = kOffRing + kStages * kSlot;   // 197632: parameters (biases ...), up to 6144 floats
constexpr int kParFloats = 6144;
constexpr int kOffBar = kOffPar + kParFloats * 4;     // full[4] empty[4] accum ready afree tmem_empty[2]
constexpr int kOffTmem = kOffBar + 16 * 8;
constexpr int kSmemBytes = kOffTmem + 16 + 1024;      // ~224 KB

struct Ctx {
    uint8_t* smem;
    uint32_t sbase, tmem;
    uint32_t bar_full, bar_empty, bar_accum, bar_ready, bar_afree, bar_tfree, bar_accum2;
};

struct Ring {
    uint32_t slot = 0, phase = 0;
    __device__ __forceinline__ void advance() {
        if (++slot == kStages) {
            slot = 0;
            phase ^= 1;
        }
    }
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// producer: stream `nstages` stages of `bytes` each.  Like the MMA issuer it is run by ALL lanes of its warp (warp-uniform control flow)
// and issues through one elected lane: see elect_one in tc_common.cuh
__device__ __forceinline__ void produce(const Ctx& c, Ring& r, const uint8_t*& src, int nstages, uint32_t bytes) {
    for (int s = 0; s < nstages; ++s) {
        mbar_wait(c.bar_empty + 8 * r.slot, r.phase ^ 1);
        if (elect_one()) {
            mbar_expect_tx(c.bar_full + 8 * r.slot, bytes);
            bulk_copy(c.sbase + kOffRing + r.slot * kSlot, src, bytes, c.bar_full + 8 * r.slot);
        }
        __syncwarp();
        src += bytes;
        r.advance();
    }
}

// MMA issuer: D[128, n] (+)= X[128, 16*ksteps] . W[n, :]^T, W stages of 64*n bytes from the ring
__device__ __forceinline__ void mma_layer(const Ctx& c, Ring& r, uint32_t dcol, int n, int ksteps, bool accumulate_first) {
    const uint32_t idesc = umma_idesc(n);
    for (int s = 0; s < ksteps; ++s) {
        mbar_wait(c.bar_full + 8 * r.slot, r.phase);
        tc_fence_after();
        const uint64_t x_hi = umma_desc(c.sbase + kOffAhi + 2 * s * kLbo, kLbo, 128);
        const uint64_t x_lo = umma_desc(c.sbase + kOffAlo + 2 * s * kLbo, kLbo, 128);
        const uint32_t wst = c.sbase + kOffRing + r.slot * kSlot;
        const uint64_t w_hi = umma_desc(wst, n * 16, 128);
        const uint64_t w_lo = umma_desc(wst + n * 32, n * 16, 128);
        if (elect_one()) {
            umma(c.tmem + dcol, x_hi, w_hi, idesc, (s > 0 || accumulate_first) ? 1u : 0u);
            umma(c.tmem + dcol, x_lo, w_hi, idesc, 1u);
            umma(c.tmem + dcol, x_hi, w_lo, idesc, 1u);
            tc_commit(c.bar_empty + 8 * r.slot);
        }
        __syncwarp();
        r.advance();
    }
}

// transposed variant for a 256-row weight block: D^T[128 rows of W (half mh), 128 tile rows] at columns dcol + 128*mh, i.e. the
// weights are the M operand and the activation tile the N operand -- the accumulator then holds FEATURES on the TMEM lanes, and
// an epilogue thread's 32 registers are 32 consecutive tile rows of one feature
__device__ __forceinline__ void mma_layer_t(const Ctx& c, Ring& r, uint32_t dcol, int ksteps) {
    const uint32_t idesc = umma_idesc(128);
    for (int s = 0; s < ksteps; ++s) {
        mbar_wait(c.bar_full + 8 * r.slot, r.phase);
        tc_fence_after();
        const uint64_t x_hi = umma_desc(c.sbase + kOffAhi + 2 * s * kLbo, kLbo, 128);
        const uint64_t x_lo = umma_desc(c.sbase + kOffAlo + 2 * s * kLbo, kLbo, 128);
        const uint32_t wst = c.sbase + kOffRing + r.slot * kSlot;  // [hi: 2 k8 blocks x 256 rows x 16 B][lo: likewise]
        if (elect_one()) {
#pragma unroll
            for (int mh = 0; mh < 2; ++mh) {
                const uint64_t w_hi = umma_desc(wst + mh * 128 * 16, 256 * 16, 128);
                const uint64_t w_lo = umma_desc(wst + 256 * 32 + mh * 128 * 16, 256 * 16, 128);
                umma(c.tmem + dcol + 128 * mh, w_hi, x_hi, idesc, s > 0 ? 1u : 0u);
                umma(c.tmem + dcol + 128 * mh, w_hi, x_lo, idesc, 1u);
                umma(c.tmem + dcol + 128 * mh, w_lo, x_hi, idesc, 1u);
            }
            tc_commit(c.bar_empty + 8 * r.slot);
        }
        __syncwarp();
        r.advance();
    }
}

// epilogue warps: fp32 rows of `src` (row pitch ld, K columns, K in {64,128,256}) -> operand tile; rows >= nrows are zero
__device__ __forceinline__ void load_rows(const Ctx& c, const float* __restrict__ src, long long row0, long long nrows, int K, int ld,
                                          int ew, int lane) {
    const int kbs = K >> 3, rpi = 32 / kbs;  // k8 blocks per row, rows per warp instruction
    const int rs = lane / kbs, kb = lane % kbs;
    for (int i = 0; i < 16; i += rpi) {
        const int r = ew * 16 + i + rs;
        const long long gr = row0 + r;
        float v[8];
        if (gr < nrows) {
            const float4* p = reinterpret_cast<const float4*>(src + gr * ld) + 2 * kb;
            const float4 u0 = p[0], u1 = p[1];
            v[0] = u0.x; v[1] = u0.y; v[2] = u0.z; v[3] = u0.w;
            v[4] = u1.x; v[5] = u1.y; v[6] = u1.z; v[7] = u1.w;
        } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = 0.f;
        }
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4*>(c.smem + kOffAhi + kb * kLbo + r * 16) = hi;
        *reinterpret_cast<uint4*>(c.smem + kOffAlo + kb * kLbo + r * 16) = lo;
    }
}

__device__ __forceinline__ void tile_ready(const Ctx& c) { warp_arrive(c.bar_ready, threadIdx.x & 31); }

// epilogue warps: accumulator (128 x n at column dcol) -> v + bias (ReLU optional) -> operand tile
template <bool RELU>
__device__ __forceinline__ void epi_to_tile(const Ctx& c, uint32_t dcol, int n, const float* bias, int lane_grp, int half, int lane) {
    const int row = lane_grp * 32 + lane;
    const int per = n >> 1;  // columns per thread
    for (int cb = 0; cb < per; cb += 32) {
        const int col0 = half * per + cb;
        float v[32];
        tmem_ld32(c.tmem + ((uint32_t)(lane_grp * 32) << 16) + dcol + col0, v);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
            float x[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float y = v[kb * 8 + t] + bias[col0 + kb * 8 + t];
                x[t] = RELU ? fmaxf(y, 0.f) : y;
            }
            uint4 hi, lo;
            split8(x, hi, lo);
            const int kblk = (col0 >> 3) + kb;
            *reinterpret_cast<uint4*>(c.smem + kOffAhi + kblk * kLbo + row * 16) = hi;
            *reinterpret_cast<uint4*>(c.smem + kOffAlo + kblk * kLbo + row * 16) = lo;
        }
    }
}

__device__ __forceinline__ void setup(Ctx& c, uint8_t* smem_raw, int tid, int warp) {
    c.smem = smem_raw;  // used directly so that the compiler keeps the shared address space (LDS/STS)
    c.sbase = smem_u32(c.smem);
    c.bar_full = c.sbase + kOffBar;
    c.bar_empty = c.bar_full + 8 * kStages;
    c.bar_accum = c.bar_empty + 8 * kStages;
    c.bar_ready = c.bar_accum + 8;
    c.bar_afree = c.bar_ready + 8;
    c.bar_tfree = c.bar_afree + 8;  // two barriers
    c.bar_accum2 = c.bar_tfree + 16;  // accumulator-complete barrier of the second TMEM buffer
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(c.bar_full + 8 * i, 1);
            mbar_init(c.bar_empty + 8 * i, 1);
        }
        mbar_init(c.bar_accum, 1);
        mbar_init(c.bar_ready, kEpiThreads / 32);  // one elected arrival per warp
        mbar_init(c.bar_afree, 1);
        mbar_init(c.bar_tfree, kEpiThreads / 32);
        mbar_init(c.bar_tfree + 8, kEpiThreads / 32);
        mbar_init(c.bar_accum2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(c.sbase + kOffTmem), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
}

__device__ __forceinline__ void teardown(const Ctx& c, int warp) {
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"(512u) : "memory");
    }
}

// packed weight bytes: a [n,k] matrix takes 4*n*k bytes (fp16 hi + lo)
constexpr size_t kPackStnFc = size_t(4) * (128 * 256 + 64 * 128 + 4096 * 64);
// ... followed by the 4096 fp32 biases of the LAST FC.  That layer is packed MERGED with the local branch's conv1 (packing.py): its
// output is M_q = W1 (fc3(f2) + I) instead of the transform T_q = fc3(f2) + I, and pn_feat_kernel's first MMA is conv1 with M_q
constexpr size_t kPackStnFcTotal = kPackStnFc + size_t(4096) * 4;
constexpr size_t kPackMlp = size_t(4) * (256 * 256 + 256 * 128 + 256 * 256 + 256 * 256);

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
    stn_fc_tc_kernel(const float* __restrict__ g, long long nq, const uint8_t* __restrict__ wpack, const float* __restrict__ b1,
                     const float* __restrict__ b2, const float* __restrict__ b3, float* __restrict__ tmat) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Ctx c;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    setup(c, smem_raw, tid, warp);
    float* par = reinterpret_cast<float*>(c.smem + kOffPar);  // b1[128] b2[64] b3[4096]
    for (int e = tid; e < 4096; e += kThreads) {
        if (e < 128) par[e] = b1[e];
        if (e < 64) par[128 + e] = b2[e];
        par[192 + e] = b3[e];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem = *reinterpret_cast<volatile uint32_t*>(c.smem + kOffTmem);
    const long long ntiles = (nq + 127) / 128;

    if (warp == 0) {
        {
            Ring r;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const uint8_t* src = wpack;
                produce(c, r, src, 16, 8192);        // fc1: n=128, k=256
                produce(c, r, src, 8, 4096);         // fc2: n=64,  k=128
                produce(c, r, src, 16 * 4, 16384);   // fc3: 16 blocks of n=256, k=64
            }
        }
    } else if (warp == 1) {
        {
            Ring r;
            uint32_t ready_phase = 0, tfree_phase[2] = {0, 0};
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait(c.bar_ready, ready_phase);
                ready_phase ^= 1;
                tc_fence_after();
                mma_layer(c, r, 0, 128, 16, false);
                if (elect_one()) tc_commit(c.bar_accum);
                __syncwarp();
                mbar_wait(c.bar_ready, ready_phase);
                ready_phase ^= 1;
                tc_fence_after();
                mma_layer(c, r, 0, 64, 8, false);
                if (elect_one()) tc_commit(c.bar_accum);
                __syncwarp();
                mbar_wait(c.bar_ready, ready_phase);
                ready_phase ^= 1;
                tc_fence_after();
                for (int nb = 0; nb < 16; ++nb) {
                    const int buf = nb & 1;
                    // the epilogue must have drained this accumulator (two blocks ago; first use per tile: previous layers)
                    mbar_wait(c.bar_tfree + 8 * buf, tfree_phase[buf] ^ 1);
                    tfree_phase[buf] ^= 1;
                    tc_fence_after();
                    mma_layer_t(c, r, buf * 256, 4);  // fc3 block nb, transposed: coalesced T stores
                    // one completion barrier PER accumulator: the issuer can run a block ahead of the epilogue, and a single
                    // barrier advancing two phases would alias in the parity wait
                    if (elect_one()) tc_commit(buf ? c.bar_accum2 : c.bar_accum);
                    __syncwarp();
                }
            }
        }
    } else {
        const int ew = warp - 2, lane_grp = warp & 3, half = ew >> 2;
        const int row = lane_grp * 32 + lane;
        uint32_t accum_phase = 0, accum2_phase = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long row0 = tile * 128;
            load_rows(c, g, row0, nq, 256, 256, ew, lane);
            tile_ready(c);
            mbar_wait(c.bar_accum, accum_phase);
            accum_phase ^= 1;
            tc_fence_after();
            epi_to_tile<true>(c, 0, 128, par, lane_grp, half, lane);
            tile_ready(c);
            mbar_wait(c.bar_accum, accum_phase);
            accum_phase ^= 1;
            tc_fence_after();
            epi_to_tile<true>(c, 0, 64, par + 128, lane_grp, half, lane);
            tile_ready(c);
            for (int nb = 0; nb < 16; ++nb) {
                const int buf = nb & 1;
                if (buf) {
                    mbar_wait(c.bar_accum2, accum2_phase);
                    accum2_phase ^= 1;
                } else {
                    mbar_wait(c.bar_accum, accum_phase);
                    accum_phase ^= 1;
                }
                tc_fence_after();
                // accumulator columns [128*mh, +128) = the tile's queries for features nb*256 + mh*128 + (TMEM lane).  This warp:
                // lanes of its lane group, query chunks 2*half and 2*half+1.  For a fixed register the 32 lanes write 32
                // consecutive features of one query: one 128-byte line per store instruction (the row-per-thread layout wrote
                // 32 different lines per instruction and was bound by LSU wavefronts)
#pragma unroll 1
                for (int cb = 0; cb < 4; ++cb) {
                    const int mh = cb >> 1, qc = 2 * half + (cb & 1);
                    float v[32];
                    tmem_ld32(c.tmem + ((uint32_t)(lane_grp * 32) << 16) + buf * 256 + mh * 128 + qc * 32, v);
                    const int f = nb * 256 + mh * 128 + row;
                    const float bias = par[192 + f];
                    float* dst = tmat + (row0 + qc * 32) * 4096 + f;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (row0 + qc * 32 + j < nq) dst[(size_t)j * 4096] = v[j] + bias;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(c.bar_tfree + 8 * buf);
            }
        }
    }
    teardown(c, warp);
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
    mlp_tc_kernel(const float* __restrict__ pooled_proj, const float* __restrict__ pooled_pn, long long nq,
                  const uint8_t* __restrict__ wpack, const float* __restrict__ bias_feat, const float* __restrict__ b0,
                  const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                  float* __restrict__ logits_out, float* __restrict__ occ_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Ctx c;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    setup(c, smem_raw, tid, warp);
    float* par = reinterpret_cast<float*>(c.smem + kOffPar);  // bias_feat[256] b0[256] b1[256] w2[512] part[2][128][2]
    for (int e = tid; e < 512; e += kThreads) {
        if (e < 256) {
            par[e] = bias_feat[e];
            par[256 + e] = b0[e];
            par[512 + e] = b1[e];
        }
        par[768 + e] = w2[e];
    }
    float* s_part = par + 1280;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem = *reinterpret_cast<volatile uint32_t*>(c.smem + kOffTmem);
    const long long ntiles = (nq + 127) / 128;

    if (warp == 0) {
        {
            Ring r;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const uint8_t* src = wpack;
                produce(c, r, src, 16, 16384);  // (W8 Wv): n=256, k=256
                produce(c, r, src, 8, 16384);   // (Wv_att A3): n=256, k=128
                produce(c, r, src, 16, 16384);  // mlp.layers.0
                produce(c, r, src, 16, 16384);  // mlp.layers.1
            }
        }
    } else if (warp == 1) {
        {
            Ring r;
            uint32_t ready_phase = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait(c.bar_ready, ready_phase);
                ready_phase ^= 1;
                tc_fence_after();
                mma_layer(c, r, 0, 256, 16, false);
                if (elect_one()) tc_commit(c.bar_afree);
                __syncwarp();  // the operand tile may be reloaded with the second branch
                mbar_wait(c.bar_ready, ready_phase);
                ready_phase ^= 1;
                tc_fence_after();
                mma_layer(c, r, 0, 256, 8, true);  // accumulates onto the first branch
                if (elect_one()) tc_commit(c.bar_accum);
                __syncwarp();
                for (int layer = 0; layer < 2; ++layer) {
                    mbar_wait(c.bar_ready, ready_phase);
                    ready_phase ^= 1;
                    tc_fence_after();
                    mma_layer(c, r, layer == 0 ? 256 : 0, 256, 16, false);
                    if (elect_one()) tc_commit(c.bar_accum);
                    __syncwarp();
                }
            }
        }
    } else {
        const int ew = warp - 2, lane_grp = warp & 3, half = ew >> 2;
        const int row = lane_grp * 32 + lane;
        uint32_t accum_phase = 0, afree_phase = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long row0 = tile * 128;
            load_rows(c, pooled_proj, row0, nq, 256, 256, ew, lane);
            tile_ready(c);
            mbar_wait(c.bar_afree, afree_phase);
            afree_phase ^= 1;
            load_rows(c, pooled_pn, row0, nq, 128, 128, ew, lane);
            tile_ready(c);
            mbar_wait(c.bar_accum, accum_phase);
            accum_phase ^= 1;
            tc_fence_after();
            epi_to_tile<false>(c, 0, 256, par, lane_grp, half, lane);  // feat = sum of the branches (no activation)
            tile_ready(c);
            mbar_wait(c.bar_accum, accum_phase);
            accum_phase ^= 1;
            tc_fence_after();
            epi_to_tile<true>(c, 256, 256, par + 256, lane_grp, half, lane);
            tile_ready(c);
            mbar_wait(c.bar_accum, accum_phase);
            accum_phase ^= 1;
            tc_fence_after();
            // mlp.layers.1 epilogue fused with mlp.layers.2 (2 outputs): partial dots over this thread's 128 columns
            float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
            for (int cb = 0; cb < 4; ++cb) {
                const int col0 = half * 128 + cb * 32;
                float v[32];
                tmem_ld32(c.tmem + ((uint32_t)(lane_grp * 32) << 16) + col0, v);
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const float y = fmaxf(v[t] + par[512 + col0 + t], 0.f);
                    l0 = fmaf(y, par[768 + col0 + t], l0);
                    l1 = fmaf(y, par[768 + 256 + col0 + t], l1);
                }
            }
            s_part[(half * 128 + row) * 2] = l0;
            s_part[(half * 128 + row) * 2 + 1] = l1;
            tc_fence_before();
            epi_bar();
            if (half == 0) {
                const long long q = row0 + row;
                if (q < nq) {
                    const float a0 = s_part[row * 2] + s_part[(128 + row) * 2] + b2[0];
                    const float a1 = s_part[row * 2 + 1] + s_part[(128 + row) * 2 + 1] + b2[1];
                    if (logits_out) {
                        logits_out[2 * q] = a0;
                        logits_out[2 * q + 1] = a1;
                    }
                    if (occ_out) {
                        const float m = fmaxf(a0, a1);
                        const float e0 = expf(a0 - m), e1 = expf(a1 - m);
                        const float s = e0 + e1;
                        occ_out[q] = e0 / s - e1 / s;
                    }
                }
            }
            epi_bar();  // s_part and the operand tile are reused by the next tile
        }
    }
    teardown(c, warp);
}

}  // namespace chain
}  // namespace tc

bool chain_tc_supported(const pps_decoder_weights* w) {
    return w->tc_stn_fc != nullptr && w->tc_mlp != nullptr && w->latent == 256 && w->stn_size == 256;
}

static int chain_configure() {
    static unsigned char configured[kMaxDevices] = {};
    if (first_use_on_device(configured)) {
        PPS_CUDA(cudaFuncSetAttribute(tc::chain::stn_fc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::chain::kSmemBytes));
        PPS_CUDA(cudaFuncSetAttribute(tc::chain::mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::chain::kSmemBytes));
    }
    return PPS_OK;
}

int stn_fc_tc_impl(const pps_decoder_weights* w, const float* g, int64_t q, float* tmat, cudaStream_t st) {
    if (q == 0) return PPS_OK;
    PPS_TRY(chain_configure());
    const long long ntiles = (q + 127) / 128;
    const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
    tc::chain::stn_fc_tc_kernel<<<grid, tc::chain::kThreads, tc::chain::kSmemBytes, st>>>(
        g, q, static_cast<const uint8_t*>(w->tc_stn_fc), w->stnf1_b, w->stnf2_b,
        reinterpret_cast<const float*>(static_cast<const uint8_t*>(w->tc_stn_fc) + tc::chain::kPackStnFc), tmat);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

// pooled_proj [q,256] (global branch before fc_value/fc8), pooled_pn [q,128] (local branch before the value matrix)
int mlp_tc_impl(const pps_decoder_weights* w, const float* pooled_proj, const float* pooled_pn, int64_t q, float* logits_out,
                float* occ_out, cudaStream_t st) {
    if (q == 0) return PPS_OK;
    PPS_TRY(chain_configure());
    const long long ntiles = (q + 127) / 128;
    const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
    tc::chain::mlp_tc_kernel<<<grid, tc::chain::kThreads, tc::chain::kSmemBytes, st>>>(
        pooled_proj, pooled_pn, q, static_cast<const uint8_t*>(w->tc_mlp), w->tc_bias_feat, w->m0_b, w->m1_b, w->m2_w, w->m2_b,
        logits_out, occ_out);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

}  // namespace pps

extern "C" size_t pps_decoder_tc_stn_fc_bytes(void) { return pps::tc::chain::kPackStnFcTotal; }
extern "C" size_t pps_decoder_tc_mlp_bytes(void) { return pps::tc::chain::kPackMlp; }
