// Fused FKAConv layer for sm_100a (source/base/nn.py:592-652): after the two InstanceNorm statistics
// passes (fka_stats_kernel below: moments of the neighbour offsets, then fc2's sums) ONE kernel does
//   neighbour index gather -> kernel-weight MLP (fc1/IN1/act/max/fc2/IN2/act/max/fc3/act * distance weight)
//   -> neighbourhood weighted sum  feat[p,m,c] = sum_j mat[p,j,m] x[ids[p,j],c]
//   -> contraction with the [cout, 16 cin] kernel on the tensor cores (tcgen05, split-fp16, fp32 accumulate in TMEM)
//   -> folded BatchNorm bias + ReLU -> out
// Neither `mat` (1 KB per point) nor `feat` (64 cin B per point) ever reaches HBM: per tile of 128 support points the 16x16
// kernel-weight matrices are parked in TMEM (256 columns next to the accumulator  --  explicit spill space: every thread reads
// back exactly the 64 values it wrote), `feat` is produced 4 input channels (64 K columns) at a time straight into the
// shared-memory A operand (UMMA canonical K-major layout, fp16 hi + lo), and the neighbour feature rows arrive through a
// double-buffered cp.async gather.
//
// Warp roles (576 threads, one CTA per SM, persistent over tiles): warp 0 = bulk-copy producer of the weight stages, warp 1 = MMA
// issuer, warps 2..17 = compute group.  Kernel-weight MLP: 16 lanes per support point (lane = neighbour), half-warp shuffles for
// the max over the neighbourhood.  Weighted sum: thread (p, s) owns kernel elements m = 4s..4s+3 of point p (64 weights in
// registers) and produces, per chunk, the two 16-byte k8 blocks {4 m} x {2 channels} of its row  --  one conflict-free STS.128 each.
// K order: k = ((c / 2) * 4 + s) * 8 + (m % 4) * 2 + (c % 2)  (ppsurf_b200/packing.py packs cv_w accordingly).
#include "tc_common.cuh"

namespace pps {
namespace tc {
namespace fka {

constexpr int kComputeWarps = 16;
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kThreads = 64 + kComputeThreads;
constexpr int kTile = 128;                      // support points per tile = MMA M
constexpr int kLbo = kTile * 16 + 16;           // k8-block pitch of the A operand (16 B pad: conflict-free k-wise)
constexpr int kAHalf = 8 * kLbo;                // one chunk = 64 K columns = 8 k8 blocks
constexpr int kABuf = 2 * kAHalf;               // fp16 hi | fp16 lo
constexpr int kXBuf = 16 * kTile * 16;          // gathered neighbour rows of one chunk: [j][p][4 channels]
constexpr int kStages = 3;
constexpr int kSlot = 16384;                    // weight stage of one k16 step: 64 * N bytes, N <= 256
constexpr int kMaxSamples = 16;                 // samples a tile of 128 flattened rows may span
constexpr int kMatPitch = 32 * 16 + 16;         // staging of one MLP round: [s][j] rows of 32 points x 16 B (+16 B pad)
constexpr int kMatBytes = 64 * kMatPitch;       // 33792

constexpr int kOffA = 0;
constexpr int kOffX = kOffA + 2 * kABuf;        // 66048
constexpr int kOffRing = kOffX + 2 * kXBuf;     // 131584
constexpr int kOffIds = kOffRing + kStages * kSlot;  // 180736
constexpr int kOffPar = kOffIds + 16 * kTile * 4;    // 188928
constexpr int kParWb2 = 0, kParWb3 = 256, kParBias = 512, kParCoef = 768;  // wb = pooled halves of fc2 / fc3, [o][c]
constexpr int kParFloats = kParCoef + kMaxSamples * 64;  // 2352
constexpr int kOffBar = kOffPar + kParFloats * 4;        // 198336
constexpr int kOffTmem = kOffBar + 16 * 8;
constexpr int kSmemBytes = kOffTmem + 16;
static_assert(kMatBytes <= 2 * kXBuf && kMatBytes <= 2 * kABuf, "MLP staging aliases the gather / operand buffers");
static_assert(kOffX % 16 == 0 && kOffRing % 128 == 0 && kOffBar % 8 == 0, "alignment");

constexpr float kInEps = 1e-5f;

// Kernel-weight MLP parameters, passed BY VALUE inside the kernel parameter block: every access below has a compile-time index,
// so the weights are constant-bank operands of the FFMAs (no shared-memory loads, no load latency in the dependent chains).
// Only the halves of fc2 / fc3 that multiply the pooled maxima are indexed by the lane and live in shared memory.
struct MlpWeights {
    float w1[48];    // fc1 [16][3]
    float w2a[256];  // fc2[:, 0:16]  [o][c]
    float w3a[256];  // fc3[:, 0:16]
};

template <int ACT>
__device__ __forceinline__ float act_t(float v) {
    // SiLU: ex2.approx (2 ulp) + rcp.approx (1 ulp), two MUFU operations per value
    return ACT == 1 ? __fdividef(v, 1.f + __expf(-v)) : fmaxf(v, 0.f);
}
__device__ __forceinline__ float hmax16(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, 16));
    return v;
}
// one (point, neighbour) lane of the kernel-weight MLP (source/base/nn.py:626-643).  cf = this sample's folded InstanceNorm
// coefficients [A1 16 | B1 16 | A2 16 | B2 16]; wb2 / wb3 = row `j` of the pooled halves of fc2 / fc3 in shared memory.
// STAGE 1 returns fc1's outputs, STAGE 2 fc2's (both before the InstanceNorm: the statistics passes), STAGE 3 the weights.
template <int ACT, int STAGE>
__device__ __forceinline__ void mlp_lane(const MlpWeights& W, float rx, float ry, float rz, float dw, const float* cf,
                                         const float* wb2, const float* wb3, float (&out)[16]) {
    float m1[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float y1 = fmaf(W.w1[3 * c + 2], rz, fmaf(W.w1[3 * c + 1], ry, W.w1[3 * c] * rx));
        if (STAGE == 1)
            out[c] = y1;
        else
            m1[c] = act_t<ACT>(fmaf(y1, cf[c], cf[16 + c]));
    }
    if (STAGE == 1) return;
    float cown = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) cown = fmaf(wb2[c], hmax16(m1[c] * dw), cown);
    float m2[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        float y = __shfl_sync(0xffffffffu, cown, o, 16);
#pragma unroll
        for (int c = 0; c < 16; ++c) y = fmaf(W.w2a[o * 16 + c], m1[c], y);
        if (STAGE == 2)
            out[o] = y;
        else
            m2[o] = act_t<ACT>(fmaf(y, cf[32 + o], cf[48 + o]));
    }
    if (STAGE == 2) return;
    cown = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) cown = fmaf(wb3[c], hmax16(m2[c] * dw), cown);
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        float y = __shfl_sync(0xffffffffu, cown, o, 16);
#pragma unroll
        for (int c = 0; c < 16; ++c) y = fmaf(W.w3a[o * 16 + c], m2[c], y);
        out[o] = act_t<ACT>(y) * dw;
    }
}

// neighbour offset (normalised) and distance weight of lane j of a support point (nn.py:597-624)
struct LaneGeom {
    float rx, ry, rz, wgt;
    int src;
};
__device__ __forceinline__ LaneGeom lane_geom(const float* __restrict__ pts, const float* __restrict__ support,
                                              const int32_t* __restrict__ ids, long long row, int j, int smp, int n_in, float alpha,
                                              float beta, float inv_radius) {
    LaneGeom g;
    g.src = smp * n_in + ids[row * 16 + j];
    g.rx = pts[3 * (size_t)g.src] - support[3 * row];
    g.ry = pts[3 * (size_t)g.src + 1] - support[3 * row + 1];
    g.rz = pts[3 * (size_t)g.src + 2] - support[3 * row + 2];
    const float dist = sqrtf(g.rx * g.rx + g.ry * g.ry + g.rz * g.rz);
    g.rx *= inv_radius;
    g.ry *= inv_radius;
    g.rz *= inv_radius;
    g.wgt = 1.f / (1.f + expf(-(-alpha * dist + beta)));
    return g;
}
__device__ __forceinline__ float lane_dw(float wgt) {
    float dsum = wgt;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o, 16);
    dsum = dsum + (dsum == 0.f ? 1.f : 0.f) + 1e-6f;
    return wgt / dsum * 16.f;
}
constexpr uint32_t kMatCol = 256;  // TMEM columns [256, 512): parked kernel-weight matrices; [0, 256): accumulator

struct Params {
    const float* x;
    const float* pts;
    const float* support;
    const int32_t* ids;
    const float *fc1, *fc2, *fc3, *in1_w, *in1_b, *in2_w, *in2_b;
    double* stats;           // [b][64]: [0,9) moments of the neighbour offsets (InstanceNorm 1), [32,64) sum / sumsq of fc2's outputs
    MlpWeights mlp;
    const uint8_t* wpack;    // per N slice: K/16 stages of [hi kb0 | hi kb1 | lo kb0 | lo kb1], block = [nsl rows][8 fp16]
    const float* bias;       // [cout] or null
    float* out;              // [rows, cout]
    long long rows;          // b * n_s
    int n_in, n_s, cin, cout, nsl, nslices, act, relu;
    float alpha, beta, inv_radius, out_scale;
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

__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }


// folded InstanceNorm coefficients of channel c of one sample from its statistics.  InstanceNorm 1 normalises fc1's outputs
// y = w . r, a LINEAR map of the neighbour offset r, so its mean and variance follow from the first and second moments of r
// over the sample: mean = w . E[r], E[y^2] = w^T E[r r^T] w  (9 numbers instead of 32 sums; the moments pass runs no MLP).
__device__ __forceinline__ void in_coef(const double* st, double cnt, int c, const float* fc1, const float* in1_w, const float* in1_b,
                                        const float* in2_w, const float* in2_b, float* cf) {
    const double w0 = fc1[3 * c], w1 = fc1[3 * c + 1], w2 = fc1[3 * c + 2];
    double mean = (w0 * st[0] + w1 * st[1] + w2 * st[2]) / cnt;
    double ey2 = (w0 * w0 * st[3] + w1 * w1 * st[6] + w2 * w2 * st[8] + 2.0 * (w0 * w1 * st[4] + w0 * w2 * st[5] + w1 * w2 * st[7])) / cnt;
    float rstd = float(1.0 / sqrt(fmax(ey2 - mean * mean, 0.0) + double(kInEps)));
    cf[c] = rstd * in1_w[c];
    cf[16 + c] = in1_b[c] - float(mean) * rstd * in1_w[c];
    mean = st[32 + c] / cnt;
    const double var = st[48 + c] / cnt - mean * mean;
    rstd = float(1.0 / sqrt(fmax(var, 0.0) + double(kInEps)));
    cf[32 + c] = rstd * in2_w[c];
    cf[48 + c] = in2_b[c] - float(mean) * rstd * in2_w[c];
}

// Statistics passes over the flattened rows: 16 support points per block of 256 threads (lane = neighbour).  A block touches at
// most two samples (n_s >= 16): its sums are reduced once per sample present.  STAGE 1: moments of r; STAGE 2: sum / sumsq of
// fc2's outputs (runs fc1 -> IN1 -> act -> max -> fc2 with the coefficients from the moments).
template <int ACT, int STAGE>
__global__ void __launch_bounds__(256) fka_stats_kernel(const __grid_constant__ Params prm) {
    constexpr int NV = STAGE == 1 ? 9 : 32;
    __shared__ float red[8][32];
    __shared__ float cf_s[2][64];
    __shared__ float wb2[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, j = tid & 15;
    const long long row0 = (long long)blockIdx.x * 16;
    const long long row = row0 + (tid >> 4);
    const bool valid = row < prm.rows;
    const long long rowc = valid ? row : prm.rows - 1;
    const int smp = (int)(rowc / prm.n_s);
    const int smp_a = (int)(row0 / prm.n_s);
    const long long last = row0 + 15 < prm.rows ? row0 + 15 : prm.rows - 1;
    const int smp_b = (int)(last / prm.n_s);
    if (STAGE == 2) {
        wb2[tid] = prm.fc2[(tid >> 4) * 32 + 16 + (tid & 15)];
        if (tid < 32) {
            const int which = tid >> 4;
            in_coef(prm.stats + (size_t)(which ? smp_b : smp_a) * 64, double(prm.n_s) * 16.0, tid & 15, prm.fc1, prm.in1_w, prm.in1_b,
                    prm.in2_w, prm.in2_b, cf_s[which]);  // the IN2 half is not valid yet and not used by this stage
        }
        __syncthreads();
    }
    const LaneGeom lg = lane_geom(prm.pts, prm.support, prm.ids, rowc, j, smp, prm.n_in, prm.alpha, prm.beta, prm.inv_radius);
    float v[32];
    if (STAGE == 1) {
        v[0] = lg.rx; v[1] = lg.ry; v[2] = lg.rz;
        v[3] = lg.rx * lg.rx; v[4] = lg.rx * lg.ry; v[5] = lg.rx * lg.rz;
        v[6] = lg.ry * lg.ry; v[7] = lg.ry * lg.rz; v[8] = lg.rz * lg.rz;
    } else {
        const float dw = lane_dw(lg.wgt);
        float y2[16];
        mlp_lane<ACT, 2>(prm.mlp, lg.rx, lg.ry, lg.rz, dw, cf_s[smp == smp_a ? 0 : 1], wb2 + j * 16, nullptr, y2);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            v[c] = y2[c];
            v[16 + c] = y2[c] * y2[c];
        }
    }
    for (int pass = 0; pass < (smp_b != smp_a ? 2 : 1); ++pass) {
        const int target = pass ? smp_b : smp_a;
        const bool mine = valid && smp == target;
        float t[32];
#pragma unroll
        for (int i = 0; i < NV; ++i) t[i] = mine ? v[i] : 0.f;
        if (STAGE == 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t[i] += __shfl_xor_sync(0xffffffffu, t[i], o);
            if (lane == 0)
#pragma unroll
                for (int i = 0; i < 9; ++i) red[warp][i] = t[i];
        } else {
            // recursive halving: lane l ends with statistic l of the warp
#pragma unroll
            for (int off = 16, nn = 16; off >= 1; off >>= 1, nn >>= 1) {
                const bool upper = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < nn; ++i) {
                    const float send = upper ? t[i] : t[i + nn];
                    const float keep = upper ? t[i + nn] : t[i];
                    t[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            red[warp][lane] = t[0];
        }
        __syncthreads();
        if (tid < NV) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) sum += red[w][tid];
            atomicAdd(prm.stats + (size_t)target * 64 + (STAGE == 1 ? 0 : 32) + tid, double(sum));
        }
        __syncthreads();
    }
}

template <int ACT>
__global__ void __launch_bounds__(kThreads, 1) fka_fused_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_wfull = sbase + kOffBar, bar_wempty = bar_wfull + 8 * kStages, bar_aready = bar_wempty + 8 * kStages,
                   bar_afree = bar_aready + 16, bar_accum = bar_afree + 16;
    float* par = reinterpret_cast<float*>(smem + kOffPar);
    int* ids_s = reinterpret_cast<int*>(smem + kOffIds);

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar_wfull + 8 * i, 1);
            mbar_init(bar_wempty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_aready + 8 * i, kComputeWarps);
            mbar_init(bar_afree + 8 * i, 1);
        }
        mbar_init(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kOffTmem), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = tid; e < 256; e += kThreads) {
        par[kParWb2 + e] = prm.fc2[(e >> 4) * 32 + 16 + (e & 15)];
        par[kParWb3 + e] = prm.fc3[(e >> 4) * 32 + 16 + (e & 15)];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + kOffTmem);

    const long long ntiles_m = (prm.rows + kTile - 1) / kTile;
    const long long ntiles = ntiles_m * prm.nslices;
    const int nchunks = prm.cin >> 2;          // 4 input channels = 64 K columns per chunk
    const int nsl = prm.nsl;                   // output channels of this CTA's slice (MMA N)
    const uint32_t stage_bytes = 64u * nsl;
    const size_t slice_bytes = (size_t)stage_bytes * 4 * nchunks;

    // warps 0 and 1 run their loops with all lanes (warp-uniform control flow) and issue through one elected lane: see elect_one
    if (warp == 0) {
        {
            Ring r;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const uint8_t* src = prm.wpack + (size_t)(t % prm.nslices) * slice_bytes;
                for (int s = 0; s < 4 * nchunks; ++s) {
                    mbar_wait(bar_wempty + 8 * r.slot, r.phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(bar_wfull + 8 * r.slot, stage_bytes);
                        bulk_copy(sbase + kOffRing + r.slot * kSlot, src, stage_bytes, bar_wfull + 8 * r.slot);
                    }
                    __syncwarp();
                    src += stage_bytes;
                    r.advance();
                }
            }
        }
    } else if (warp == 1) {
        {
            Ring r;
            const uint32_t idesc = umma_idesc(nsl);
            uint32_t g = 0;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                for (int kc = 0; kc < nchunks; ++kc, ++g) {
                    const uint32_t buf = g & 1u, use = g >> 1;
                    mbar_wait(bar_aready + 8 * buf, use & 1u);
                    tc_fence_after();
                    const uint32_t abase = sbase + kOffA + buf * kABuf;
#pragma unroll 1
                    for (int s = 0; s < 4; ++s) {
                        mbar_wait(bar_wfull + 8 * r.slot, r.phase);
                        tc_fence_after();
                        const uint64_t a_hi = umma_desc(abase + 2 * s * kLbo, kLbo, 128);
                        const uint64_t a_lo = umma_desc(abase + kAHalf + 2 * s * kLbo, kLbo, 128);
                        const uint32_t wst = sbase + kOffRing + r.slot * kSlot;
                        const uint64_t w_hi = umma_desc(wst, nsl * 16, 128);
                        const uint64_t w_lo = umma_desc(wst + nsl * 32, nsl * 16, 128);
                        if (elect_one()) {
                            umma(tmem, a_hi, w_hi, idesc, (kc > 0 || s > 0) ? 1u : 0u);
                            umma(tmem, a_lo, w_hi, idesc, 1u);
                            umma(tmem, a_hi, w_lo, idesc, 1u);
                            tc_commit(bar_wempty + 8 * r.slot);
                        }
                        __syncwarp();
                        r.advance();
                    }
                    if (elect_one()) tc_commit(bar_afree + 8 * buf);
                    __syncwarp();
                }
                if (elect_one()) tc_commit(bar_accum);
                __syncwarp();
            }
        }
    } else {
        const int cw = warp - 2;               // compute warp 0..15
        const int q = warp & 3;                // TMEM lane quarter this warp may access
        const int s = cw >> 2;                 // kernel-element slice (m = 4s..4s+3) in the weighted-sum phase
        const int ct = tid - 64;               // 0..511
        const unsigned full = 0xffffffffu;
        const float* coef = par + kParCoef;
        uint32_t g = 0, accum_phase = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const long long tm = t / prm.nslices;
            const int sl = (int)(t % prm.nslices);
            const long long row0 = tm * kTile;
            const long long last_row = (row0 + kTile - 1 < prm.rows ? row0 + kTile - 1 : prm.rows - 1);
            const int b_first = (int)(row0 / prm.n_s);
            const int nsamp = (int)(last_row / prm.n_s) - b_first + 1;
            // ---- InstanceNorm coefficients of the samples this tile touches, bias of this slice
            if (ct < nsamp * 16)
                in_coef(prm.stats + (size_t)(b_first + (ct >> 4)) * 64, double(prm.n_s) * 16.0, ct & 15, prm.fc1, prm.in1_w, prm.in1_b,
                        prm.in2_w, prm.in2_b, par + kParCoef + (ct >> 4) * 64);
            if (ct < nsl) par[kParBias + ct] = prm.bias ? prm.bias[sl * nsl + ct] : 0.f;
            compute_bar();

            // ---- kernel-weight MLP, 4 rounds of 32 points: lane = neighbour j of point pl
#pragma unroll 1
            for (int r = 0; r < 4; ++r) {
                const int pl = 32 * r + 2 * cw + (lane >> 4);
                const int j = lane & 15;
                const long long row = row0 + pl;
                const bool valid = row < prm.rows;
                const long long rowc = valid ? row : prm.rows - 1;  // rows past the end recompute the last row; their result is dropped
                const int smp = (int)(rowc / prm.n_s);
                const LaneGeom lg = lane_geom(prm.pts, prm.support, prm.ids, rowc, j, smp, prm.n_in, prm.alpha, prm.beta, prm.inv_radius);
                ids_s[j * kTile + pl] = lg.src;
                const float dw = lane_dw(lg.wgt);
                float outv[16];
                mlp_lane<ACT, 3>(prm.mlp, lg.rx, lg.ry, lg.rz, dw, coef + (smp - b_first) * 64, par + kParWb2 + j * 16,
                                 par + kParWb3 + j * 16, outv);
                uint8_t* stage = smem + ((r & 1) ? kOffA : kOffX);
#pragma unroll
                for (int sq = 0; sq < 4; ++sq)
                    *reinterpret_cast<float4*>(stage + (sq * 16 + j) * kMatPitch + (pl & 31) * 16) =
                        valid ? make_float4(outv[4 * sq], outv[4 * sq + 1], outv[4 * sq + 2], outv[4 * sq + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                compute_bar();
                if (q == r) {
                    // park this round's matrices in TMEM: thread = (point 32 r + lane, slice s), 64 values [j][m % 4]
                    const uint8_t* rd = stage + (s * 16) * kMatPitch + lane * 16;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float v[32];
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const float4 f = *reinterpret_cast<const float4*>(rd + (h * 8 + jj) * kMatPitch);
                            v[4 * jj] = f.x;
                            v[4 * jj + 1] = f.y;
                            v[4 * jj + 2] = f.z;
                            v[4 * jj + 3] = f.w;
                        }
                        tmem_st32(tmem + ((uint32_t)(32 * q) << 16) + kMatCol + s * 64 + h * 32, v);
                    }
                    tmem_st_wait();
                }
            }
            tc_fence_before();
            compute_bar();  // every staging buffer has been drained: the gather / operand buffers are free again
            tc_fence_after();

            // ---- weighted sum + operand tile: thread = (point p, slice s)
            const int p = 32 * q + lane;
            float w[64];
            {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + kMatCol + s * 64, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) w[i] = v[i];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + kMatCol + s * 64 + 32, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) w[32 + i] = v[i];
            }
            // gather of chunk kc: 2048 16-byte pieces x[ids[p,j], 4 kc .. 4 kc + 3] -> xs[kc & 1][j][p]
            auto gather = [&](int kc) {
                const uint32_t dst = sbase + kOffX + (kc & 1) * kXBuf;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = ct + i * kComputeThreads;  // = j * 128 + p'
                    cp_async16(dst + e * 16, prm.x + (size_t)ids_s[e] * prm.cin + 4 * kc);
                }
                cp_async_commit();
            };
            gather(0);
#pragma unroll 1
            for (int kc = 0; kc < nchunks; ++kc, ++g) {
                cp_async_wait_all();
                compute_bar();  // chunk kc has landed for everybody; everybody is done reading the other buffer
                if (kc + 1 < nchunks) gather(kc + 1);
                const uint32_t buf = g & 1u, use = g >> 1;
                const uint8_t* xs = smem + kOffX + (kc & 1) * kXBuf + p * 16;
                float acc[4][4];
#pragma unroll
                for (int mm = 0; mm < 4; ++mm)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[mm][c] = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 xv = *reinterpret_cast<const float4*>(xs + j * (kTile * 16));
#pragma unroll
                    for (int mm = 0; mm < 4; ++mm) {
                        const float wv = w[4 * j + mm];
                        acc[mm][0] = fmaf(wv, xv.x, acc[mm][0]);
                        acc[mm][1] = fmaf(wv, xv.y, acc[mm][1]);
                        acc[mm][2] = fmaf(wv, xv.z, acc[mm][2]);
                        acc[mm][3] = fmaf(wv, xv.w, acc[mm][3]);
                    }
                }
                mbar_wait(bar_afree + 8 * buf, (use & 1u) ^ 1u);  // the MMAs that read this buffer two chunks ago are done
                uint8_t* abuf = smem + kOffA + buf * kABuf + p * 16;
#pragma unroll
                for (int cp = 0; cp < 2; ++cp) {
                    float v8[8];
#pragma unroll
                    for (int mm = 0; mm < 4; ++mm) {
                        v8[2 * mm] = acc[mm][2 * cp];
                        v8[2 * mm + 1] = acc[mm][2 * cp + 1];
                    }
                    uint4 hi, lo;
                    split8(v8, hi, lo);
                    *reinterpret_cast<uint4*>(abuf + (cp * 4 + s) * kLbo) = hi;
                    *reinterpret_cast<uint4*>(abuf + kAHalf + (cp * 4 + s) * kLbo) = lo;
                }
                warp_arrive(bar_aready + 8 * buf, lane);
            }

            // ---- epilogue: accumulator (128 x nsl) + bias, ReLU -> out; slice s stores columns [32 s, 32 s + 32) (nsl <= 128)
            // or [64 s, 64 s + 64) (nsl = 256)
            mbar_wait(bar_accum, accum_phase);
            accum_phase ^= 1;
            tc_fence_after();
            const int per = nsl > 128 ? 64 : 32;
            const long long orow = row0 + p;
            for (int cb = 0; cb < per; cb += 32) {
                const int col0 = s * per + cb;
                if (col0 < nsl) {  // warp-uniform
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + col0, v);
                    if (orow < prm.rows) {
                        float4* dst = reinterpret_cast<float4*>(prm.out + orow * prm.cout + sl * nsl + col0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 o;
                            o.x = fmaf(v[4 * i], prm.out_scale, par[kParBias + col0 + 4 * i]);
                            o.y = fmaf(v[4 * i + 1], prm.out_scale, par[kParBias + col0 + 4 * i + 1]);
                            o.z = fmaf(v[4 * i + 2], prm.out_scale, par[kParBias + col0 + 4 * i + 2]);
                            o.w = fmaf(v[4 * i + 3], prm.out_scale, par[kParBias + col0 + 4 * i + 3]);
                            if (prm.relu) {
                                o.x = fmaxf(o.x, 0.f);
                                o.y = fmaxf(o.y, 0.f);
                                o.z = fmaxf(o.z, 0.f);
                                o.w = fmaxf(o.w, 0.f);
                            }
                            dst[i] = o;
                        }
                    }
                }
            }
            tc_fence_before();
            compute_bar();  // the accumulator, the coefficient table and ids_s are reused by the next tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace fka
}  // namespace tc

// rows = b * n_s flattened; returns false when the shape is outside what the fused kernel was built for
bool fka_fused_supported(const pps_fkaconv_weights* w, int kn, int64_t n_s) {
    const int nsl = w->cout > 256 ? 256 : w->cout;
    return w->tc_pack != nullptr && w->mlp_host != nullptr && kn == 16 && n_s >= 16 && w->cin >= 4 && w->cin % 4 == 0 &&
           w->cout % nsl == 0 && (nsl == 32 || nsl == 64 || nsl == 128 || nsl == 256);
}

template <int ACT>
static int fka_fused_launch(const tc::fka::Params& p, cudaStream_t st) {
    static unsigned char configured[kMaxDevices] = {};
    if (first_use_on_device(configured)) {
        PPS_CUDA(cudaFuncSetAttribute(tc::fka::fka_fused_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::fka::kSmemBytes));
    }
    const unsigned sblocks = (unsigned)((p.rows + 15) / 16);
    tc::fka::fka_stats_kernel<ACT, 1><<<sblocks, 256, 0, st>>>(p);
    PPS_LAUNCH_CHECK();
    tc::fka::fka_stats_kernel<ACT, 2><<<sblocks, 256, 0, st>>>(p);
    PPS_LAUNCH_CHECK();
    const long long ntiles = ((p.rows + tc::fka::kTile - 1) / tc::fka::kTile) * p.nslices;
    const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
    tc::fka::fka_fused_kernel<ACT><<<grid, tc::fka::kThreads, tc::fka::kSmemBytes, st>>>(p);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

// the whole layer: statistics (zeroed here), moments pass, fc2 statistics pass, fused kernel
int fka_fused_impl(const pps_fkaconv_weights* w, const float* x, const float* pts, const float* support, const int32_t* ids,
                   int64_t b, int64_t n_in, int64_t n_s, double* stats, float* out, cudaStream_t st) {
    tc::fka::Params p;
    p.x = x;
    p.pts = pts;
    p.support = support;
    p.ids = ids;
    p.fc1 = w->fc1;
    p.fc2 = w->fc2;
    p.fc3 = w->fc3;
    p.in1_w = w->in1_w;
    p.in1_b = w->in1_b;
    p.in2_w = w->in2_w;
    p.in2_b = w->in2_b;
    p.stats = stats;
    const float* h = w->mlp_host;  // fc1 [16,3] | fc2 [16,32] | fc3 [16,32] on the HOST: they travel in the kernel parameter block
    for (int e = 0; e < 48; ++e) p.mlp.w1[e] = h[e];
    for (int o = 0; o < 16; ++o)
        for (int c = 0; c < 16; ++c) {
            p.mlp.w2a[o * 16 + c] = h[48 + o * 32 + c];
            p.mlp.w3a[o * 16 + c] = h[48 + 512 + o * 32 + c];
        }
    p.wpack = static_cast<const uint8_t*>(w->tc_pack);
    p.bias = w->out_bias;
    p.out = out;
    p.rows = b * n_s;
    p.n_in = (int)n_in;
    p.n_s = (int)n_s;
    p.cin = w->cin;
    p.cout = w->cout;
    p.nsl = w->cout > 256 ? 256 : w->cout;
    p.nslices = w->cout / p.nsl;
    p.act = w->act;
    p.relu = w->out_relu;
    p.alpha = w->alpha;
    p.beta = w->beta;
    p.inv_radius = 1.f / w->norm_radius;
    p.out_scale = w->tc_out_scale;
    PPS_CUDA(cudaMemsetAsync(stats, 0, (size_t)b * 64 * sizeof(double), st));
    return w->act == 1 ? fka_fused_launch<1>(p, st) : fka_fused_launch<0>(p, st);
}

}  // namespace pps
