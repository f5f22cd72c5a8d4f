// Fused FKAConv layer for sm_100a (source/base/nn.py:592-652): after the two InstanceNorm statistics
// passes (encoder.cu) ONE kernel does
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
constexpr int kParW1 = 0, kParW2 = 48, kParW3 = 560, kParBias = 1072, kParCoef = 1328;
constexpr int kParFloats = kParCoef + kMaxSamples * 64;  // 2352
constexpr int kOffBar = kOffPar + kParFloats * 4;        // 198336
constexpr int kOffTmem = kOffBar + 16 * 8;
constexpr int kSmemBytes = kOffTmem + 16;
static_assert(kMatBytes <= 2 * kXBuf && kMatBytes <= 2 * kABuf, "MLP staging aliases the gather / operand buffers");
static_assert(kOffX % 16 == 0 && kOffRing % 128 == 0 && kOffBar % 8 == 0, "alignment");

constexpr float kInEps = 1e-5f;
constexpr uint32_t kMatCol = 256;  // TMEM columns [256, 512): parked kernel-weight matrices; [0, 256): accumulator

struct Params {
    const float* x;
    const float* pts;
    const float* support;
    const int32_t* ids;
    const float *fc1, *fc2, *fc3, *in1_w, *in1_b, *in2_w, *in2_b;
    const double* stats;     // [b][64]: sum / sumsq of fc1 outputs, sum / sumsq of fc2 outputs
    const uint8_t* wpack;    // per N slice: K/16 stages of [hi kb0 | hi kb1 | lo kb0 | lo kb1], block = [nsl rows][8 fp16]
    const float* bias;       // [cout] or null
    float* out;              // [rows, cout]
    long long rows;          // b * n_s
    int n_in, n_s, cin, cout, nsl, nslices, act, relu;
    float alpha, beta, inv_radius;
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

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// SiLU / ReLU of the kernel-weight MLP.  ex2.approx (2 ulp) + rcp.approx (1 ulp): two MUFU operations per value; the accurate
// expf + IEEE division of the statistics kernels costs 4x the issue slots and this MLP is the SIMT hot spot of the layer.
__device__ __forceinline__ float act_fn(float v, int act) { return act == 1 ? __fdividef(v, 1.f + __expf(-v)) : fmaxf(v, 0.f); }

__global__ void __launch_bounds__(kThreads, 1) fka_fused_kernel(const Params prm) {
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
    for (int e = tid; e < 512; e += kThreads) {
        if (e < 48) par[kParW1 + e] = prm.fc1[e];
        par[kParW2 + e] = prm.fc2[e];
        par[kParW3 + e] = prm.fc3[e];
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

    if (warp == 0) {
        if (lane == 0) {
            Ring r;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const uint8_t* src = prm.wpack + (size_t)(t % prm.nslices) * slice_bytes;
                for (int s = 0; s < 4 * nchunks; ++s) {
                    mbar_wait(bar_wempty + 8 * r.slot, r.phase ^ 1);
                    mbar_expect_tx(bar_wfull + 8 * r.slot, stage_bytes);
                    bulk_copy(sbase + kOffRing + r.slot * kSlot, src, stage_bytes, bar_wfull + 8 * r.slot);
                    src += stage_bytes;
                    r.advance();
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
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
                        umma(tmem, a_hi, w_hi, idesc, (kc > 0 || s > 0) ? 1u : 0u);
                        umma(tmem, a_lo, w_hi, idesc, 1u);
                        umma(tmem, a_hi, w_lo, idesc, 1u);
                        tc_commit(bar_wempty + 8 * r.slot);
                        r.advance();
                    }
                    tc_commit(bar_afree + 8 * buf);
                }
                tc_commit(bar_accum);
            }
        }
    } else {
        const int cw = warp - 2;               // compute warp 0..15
        const int q = warp & 3;                // TMEM lane quarter this warp may access
        const int s = cw >> 2;                 // kernel-element slice (m = 4s..4s+3) in the weighted-sum phase
        const int ct = tid - 64;               // 0..511
        const unsigned full = 0xffffffffu;
        const float* W1 = par + kParW1;
        const float* W2 = par + kParW2;
        const float* W3 = par + kParW3;
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
            if (ct < nsamp * 16) {
                const int smp = ct >> 4, c = ct & 15;
                const double* st = prm.stats + (size_t)(b_first + smp) * 64;
                const double cnt = double(prm.n_s) * 16.0;
                double mean = st[c] / cnt, var = st[16 + c] / cnt - mean * mean;
                float rstd = float(1.0 / sqrt(fmax(var, 0.0) + double(kInEps)));
                float* cf = par + kParCoef + smp * 64;
                cf[c] = rstd * prm.in1_w[c];
                cf[16 + c] = prm.in1_b[c] - float(mean) * rstd * prm.in1_w[c];
                mean = st[32 + c] / cnt;
                var = st[48 + c] / cnt - mean * mean;
                rstd = float(1.0 / sqrt(fmax(var, 0.0) + double(kInEps)));
                cf[32 + c] = rstd * prm.in2_w[c];
                cf[48 + c] = prm.in2_b[c] - float(mean) * rstd * prm.in2_w[c];
            }
            if (ct < nsl) par[kParBias + ct] = prm.bias ? prm.bias[sl * nsl + ct] : 0.f;
            compute_bar();

            // ---- kernel-weight MLP, 4 rounds of 32 points: lane = neighbour j of point pl
#pragma unroll 1
            for (int r = 0; r < 4; ++r) {
                const int pl = 32 * r + 2 * cw + (lane >> 4);
                const int j = lane & 15;
                const long long row = row0 + pl;
                const bool valid = row < prm.rows;
                const long long rowc = valid ? row : 0;
                const int smp = (int)(rowc / prm.n_s);
                const float* cf = coef + (smp - b_first) * 64;
                float rx = 0.f, ry = 0.f, rz = 0.f, wgt = 0.f;
                int src = 0;
                if (valid) {
                    src = smp * prm.n_in + prm.ids[rowc * 16 + j];
                    rx = prm.pts[3 * (size_t)src] - prm.support[3 * rowc];
                    ry = prm.pts[3 * (size_t)src + 1] - prm.support[3 * rowc + 1];
                    rz = prm.pts[3 * (size_t)src + 2] - prm.support[3 * rowc + 2];
                    const float dist = sqrtf(rx * rx + ry * ry + rz * rz);
                    rx *= prm.inv_radius;
                    ry *= prm.inv_radius;
                    rz *= prm.inv_radius;
                    wgt = 1.f / (1.f + expf(-(-prm.alpha * dist + prm.beta)));
                }
                ids_s[j * kTile + pl] = src;
                float dsum = wgt;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) dsum += __shfl_xor_sync(full, dsum, o, 16);
                dsum = dsum + (dsum == 0.f ? 1.f : 0.f) + 1e-6f;
                const float dw = wgt / dsum * 16.f;

                float m1[16], mp[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float y1 = W1[3 * c] * rx + W1[3 * c + 1] * ry + W1[3 * c + 2] * rz;
                    m1[c] = act_fn(y1 * cf[c] + cf[16 + c], prm.act);
                    float v = valid ? m1[c] * dw : -INFINITY;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(full, v, o, 16));
                    mp[c] = v;
                }
                float cown = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) cown = fmaf(W2[j * 32 + 16 + c], mp[c], cown);
                float m2[16];
#pragma unroll
                for (int o = 0; o < 16; ++o) {
                    float y = __shfl_sync(full, cown, o, 16);
#pragma unroll
                    for (int c = 0; c < 16; ++c) y = fmaf(W2[o * 32 + c], m1[c], y);
                    m2[o] = act_fn(y * cf[32 + o] + cf[48 + o], prm.act);
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    float v = valid ? m2[c] * dw : -INFINITY;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(full, v, o, 16));
                    mp[c] = v;
                }
                cown = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) cown = fmaf(W3[j * 32 + 16 + c], mp[c], cown);
                uint8_t* stage = smem + ((r & 1) ? kOffA : kOffX);
#pragma unroll
                for (int sq = 0; sq < 4; ++sq) {
                    float o4[4];
#pragma unroll
                    for (int mm = 0; mm < 4; ++mm) {
                        const int o = 4 * sq + mm;
                        float y = __shfl_sync(full, cown, o, 16);
#pragma unroll
                        for (int c = 0; c < 16; ++c) y = fmaf(W3[o * 32 + c], m2[c], y);
                        o4[mm] = valid ? act_fn(y, prm.act) * dw : 0.f;
                    }
                    *reinterpret_cast<float4*>(stage + (sq * 16 + j) * kMatPitch + (pl & 31) * 16) = make_float4(o4[0], o4[1], o4[2], o4[3]);
                }
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
                            o.x = v[4 * i] + par[kParBias + col0 + 4 * i];
                            o.y = v[4 * i + 1] + par[kParBias + col0 + 4 * i + 1];
                            o.z = v[4 * i + 2] + par[kParBias + col0 + 4 * i + 2];
                            o.w = v[4 * i + 3] + par[kParBias + col0 + 4 * i + 3];
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
    return w->tc_pack != nullptr && kn == 16 && n_s >= 16 && w->cin >= 4 && w->cin % 4 == 0 && w->cout % nsl == 0 &&
           (nsl == 32 || nsl == 64 || nsl == 128 || nsl == 256);
}

int fka_fused_impl(const pps_fkaconv_weights* w, const float* x, const float* pts, const float* support, const int32_t* ids,
                   int64_t b, int64_t n_in, int64_t n_s, const double* stats, float* out, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        PPS_CUDA(cudaFuncSetAttribute(tc::fka::fka_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::fka::kSmemBytes));
        configured = true;
    }
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
    const long long ntiles = ((p.rows + tc::fka::kTile - 1) / tc::fka::kTile) * p.nslices;
    const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
    tc::fka::fka_fused_kernel<<<grid, tc::fka::kThreads, tc::fka::kSmemBytes, st>>>(p);
    PPS_LAUNCH_CHECK();
    return PPS_OK;
}

}  // namespace pps
