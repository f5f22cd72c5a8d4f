// Tensor-core path of the decoder's local branch (PointNetfeat with feature-STN and attention pooling,
// source/base/nn.py:305-373,162-190,84-96) for patches of P <= 256 points: two persistent warp-specialised tcgen05 kernels
// on tiles of 128 rows = two half-tiles of 64 point slots (rows >= P are padding; patches of up to 256 points span up to
// four half-tiles, merged by an atomic max / an online-softmax combine), same split-fp16 scheme and roles as decode_tc.cu.
//
//   pn_stn_kernel   patches -> conv0a (SIMT, K=3) -> conv0b -> [a1 to global] -> stn.conv1 -> stn.conv2 -> stn.conv3
//                   computed TRANSPOSED (weights as the M=128 operand, the activation tile as the N=128 operand) so that
//                   the max over a patch's points is a max over accumulator COLUMNS inside one thread -> g [q,S]
//   (SIMT linears)  g -> stn.fc1 -> stn.fc2 -> stn.fc3 (+I) = T [q,64,64]          (M = queries: batched over the chunk)
//   pn_feat_kernel  a1, T -> x' = T_q . a1 (per-query operand built in shared memory) -> conv1 -> conv2 -> attention
//                   logits, softmax over the patch, pooled [q,128]
// The pooled vector then goes through the merged (att.fc_value . bn3 . conv3) matrix in linear_impl.
#include "tc_common.cuh"

namespace pps {
namespace tc {

constexpr int kPnRows = 128;
constexpr int kPnLbo = kPnRows * 16 + 16;  // 2064: padded k8-block pitch of an activation tile
constexpr int kPnThreads = 320;
constexpr int kPnEpiThreads = 256;

__device__ __forceinline__ void pn_epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kPnEpiThreads) : "memory"); }

// bytes of the packed weights
// Patches of P <= 64 points fill one 64-row half-tile per query; larger patches (ppsurf_200nn: P = 200) span G = ceil(P/64)
// consecutive half-tiles.  Half-tile ht = 2*tile + (row >> 6) belongs to query ht / G and holds its points [64*(ht % G), +64).
struct PnRow {
    long long q;
    int p;
};
__device__ __forceinline__ PnRow pn_row(long long tile, int row, int G) {
    const long long ht = 2 * tile + (row >> 6);
    PnRow r;
    if (G == 1) {  // the common case (P <= 64) must not pay for 64-bit divisions: they cost 6 % of the whole decode
        r.q = ht;
        r.p = row & 63;
    } else {
        r.q = ht / G;
        r.p = int(ht - r.q * G) * 64 + (row & 63);
    }
    return r;
}

constexpr int kPartialStride = 132;  // floats per (query, half-tile) partial of the attention pooling: pooled[128], max, sum, pad

constexpr size_t kPackStnBytes = size_t(4) * 4096 * 2 + size_t(4) * 8192 + size_t(16) * 8192;  // conv0b, stn1, stn2, stn3
constexpr size_t kPackFeatBytes = size_t(4) * 4096 + size_t(4) * 8192;                         // conv1, conv2

// ---------------------------------------------------------------------------------------------------------------------
// kernel A: conv0a .. stn.conv3 + max
// ---------------------------------------------------------------------------------------------------------------------
namespace stn {
// The kernel runs as CTA PAIRS (cluster of 2 on one TPC, tcgen05 cta_group::2) like projection_tc_kernel: every MMA is M = 256.
//   conv0b / stn.conv1 / stn.conv2   D[256 rows, n] = X . W^T: A = the two CTAs' 128-row tiles, B = n/2 weight rows from each CTA
//   stn.conv3 (transposed)           D^T[256 features, 256 rows] = W . X^T: A = 128 of the 256 features from each CTA, B = the two tiles
// so a CTA streams HALF of every weight stage through its ring and reads half of the weight operand per MMA.  With one CTA per MMA the
// kernel was bound by shared-memory bandwidth: an M128 N128 K16 instruction reads 8 KB in its 64 tensor cycles (= the 128 B/clk of the
// SM), the N = 64 layers need 192 B/clk, and the epilogues' operand stores and the ring refill (192 KB per tile) come on top --
// ~1 MB of shared-memory traffic per tile for 4.6 k cycles of tensor work.  Pairs: 0.65 MB, ring 96 KB per CTA and tile.
// After stn.conv3 a CTA holds ITS 128 features of all four queries of the pair-tile: the max over a patch stays a max over
// accumulator columns inside one thread and nothing crosses the pair.
// Rank 0 issues the MMAs; the barriers that gate them live in rank 0 and collect arrivals from both CTAs; completions are multicast.
constexpr int kOffAhi = 0;
constexpr int kABytes = 16 * kPnLbo;                 // up to 128 columns
constexpr int kOffAlo = kOffAhi + kABytes;           // 33024
constexpr int kOffRing = kOffAlo + kABytes;          // 66048
// Weight ring of one CTA: 2 slots of 16 KB.  Per pair-tile a CTA receives 7 slot fills: conv0b (8 KB: 4 k16 steps x 32 weight rows),
// stn.conv1 (8 KB), stn.conv2 (16 KB: 4 steps x 64 rows), stn.conv3 (4 x 16 KB: 2 steps x 128 features each)
constexpr int kSlot = 16384;
constexpr int kStages = 2;
constexpr int kFills = 7;
constexpr int kOffPar = kOffRing + kStages * kSlot;  // 98816: w0a[192] b0a[64] b0b[64] bs1[64] bs2[128] bs3[256]
constexpr int kParFloats = 192 + 64 + 64 + 64 + 128 + 256;
constexpr int kOffBar = kOffPar + kParFloats * 4;    // full[kStages] empty[kStages] accum aready
constexpr int kSubBytes = ((kOffBar + (2 * kStages + 2) * 8 + 1023) / 1024) * 1024;  // one chain (sub-block): 100 KB
constexpr int kSmemBytes = 2 * kSubBytes + 16 + 1024;  // two chains per CTA + the TMEM address
static_assert(kSmemBytes <= 232448, "one CTA per SM");
constexpr int kTmemCols = 256;                        // per chain
// packed weights: per fill [CTA 0's slot | CTA 1's slot]
__device__ __forceinline__ uint32_t fill_bytes(int f) { return f < 2 ? 8192u : 16384u; }
__device__ __forceinline__ uint32_t fill_offset(int f, uint32_t rank) {  // byte offset of CTA `rank`'s part of fill f
    const uint32_t base = f < 2 ? 16384u * f : 32768u + 32768u * (f - 2);
    return base + rank * fill_bytes(f);
}
}  // namespace stn

// MULTI = false: every patch fits one half-tile (P <= 64, G == 1): the instantiation carries no divisions and no partials
template <bool MULTI>
__global__ void __launch_bounds__(2 * kPnThreads, 1)
    pn_stn_kernel(const float* __restrict__ patches, long long nq, int P, int G_arg, const uint8_t* __restrict__ wpack,
                  const float* __restrict__ w0a, const float* __restrict__ b0a, const float* __restrict__ b0b,
                  const float* __restrict__ bs1, const float* __restrict__ bs2, const float* __restrict__ bs3,
                  float* __restrict__ a1_out, float* __restrict__ g_out) {
    using namespace stn;
    const int G = MULTI ? G_arg : 1;
    extern __shared__ __align__(1024) uint8_t smem_cta[];  // used directly: the compiler keeps the shared address space (LDS/STS)
    // Two independent CHAINS per CTA (sub-blocks of 320 threads with their own operand tile, ring, barriers and 256 TMEM columns), one
    // CTA per SM: two CTAs per SM would each allocate their 256 columns on their own, and nothing makes the two CTAs of a pair receive
    // the SAME columns on their two SMs -- which cta_group::2 needs (one TMEM address per MMA for both CTAs)
    const int sub = threadIdx.x >= kPnThreads ? 1 : 0;
    uint8_t* smem = smem_cta + sub * kSubBytes;
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x - sub * kPnThreads, warp = tid >> 5, lane = tid & 31;  // within the sub-block
    float* s_par = reinterpret_cast<float*>(smem + kOffPar);
    float* s_w0a = s_par;
    float* s_b0a = s_par + 192;
    float* s_b0b = s_b0a + 64;
    float* s_bs1 = s_b0b + 64;
    float* s_bs2 = s_bs1 + 64;
    float* s_bs3 = s_bs2 + 128;
    volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem_cta + 2 * kSubBytes);
    const uint32_t bar_full = sbase + kOffBar, bar_empty = bar_full + 8 * kStages, bar_accum = bar_empty + 8 * kStages,
                   bar_aready = bar_accum + 8;
    const uint32_t crank = cluster_ctarank();  // 0 = leader of the pair
    const uint32_t lead_full = map_to_cta(bar_full, 0), lead_aready = map_to_cta(bar_aready, 0);

    for (int e = tid; e < 256; e += kPnThreads) {
        if (e < 192) s_w0a[e] = w0a[e];
        if (e < 64) {
            s_b0a[e] = b0a[e];
            s_b0b[e] = b0b[e];
            s_bs1[e] = bs1[e];
        }
        if (e < 128) s_bs2[e] = bs2[e];
        s_bs3[e] = bs3[e];
    }
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar_full + 8 * i, crank == 0 ? 2 : 1);  // leader: own producer + the peer's relay
            mbar_init(bar_empty + 8 * i, 1);
        }
        mbar_init(bar_accum, 1);
        mbar_init(bar_aready, 2 * kPnEpiThreads / 32);  // one elected arrival per epilogue warp of the pair (used in the leader)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x >> 5 == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_cta + 2 * kSubBytes)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the barriers of both CTAs are initialised before any remote arrival or multicast commit
    tc_fence_after();
    const uint32_t tmem = *s_tmem + (uint32_t)(sub * kTmemCols);
    // pair-tile = the two tiles of a pair of sub-blocks; both run the same number of iterations (a tile past the end has no valid row)
    const long long ntiles = (nq * G + 1) / 2, npt = (ntiles + 1) / 2;
    const long long pair = (blockIdx.x >> 1) * 2 + sub, npairs = (gridDim.x >> 1) * 2;
    const long long iters = npt > pair ? (npt - pair + npairs - 1) / npairs : 0;

    // warps 0 and 1 run their loops with all lanes (warp-uniform control flow) and issue through one elected lane: see elect_one
    if (warp == 0) {
        uint32_t slot = 0, phase = 0;
        for (long long it = 0; it < iters; ++it) {
            for (int f = 0; f < kFills; ++f) {
                mbar_wait(bar_empty + 8 * slot, phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(bar_full + 8 * slot, fill_bytes(f));
                    bulk_copy(sbase + kOffRing + slot * kSlot, wpack + fill_offset(f, crank), fill_bytes(f), bar_full + 8 * slot);
                }
                __syncwarp();
                if (++slot == kStages) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (crank == 0) {
            // ---- MMA issuer of the pair
            uint32_t slot = 0, phase = 0, ready_phase = 0;
            for (long long it = 0; it < iters; ++it) {
                for (int layer = 0; layer < 4; ++layer) {
                    // the layer's first weight slot BEFORE the operand tile: it has usually landed long ago, and a wait on a completed
                    // mbarrier still costs ~90 cycles that would otherwise sit between the tile's release and the first MMA
                    mbar_wait_cluster(bar_full + 8 * slot, phase);
                    mbar_wait_cluster(bar_aready, ready_phase);
                    ready_phase ^= 1;
                    tc_fence_after();
                    if (layer < 3) {
                        // D[256 rows, n] = X[rows, 64] . W[n, 64]^T: one ring slot, 4 k16 steps of n/2 weight rows per CTA
                        const int n = layer < 2 ? 64 : 128;
                        const uint32_t idesc = umma_idesc2(n);
                        const uint32_t step_bytes = 32u * n;
                        if (elect_one()) {
#pragma unroll
                            for (int s = 0; s < 4; ++s) {
                                const uint64_t x_hi = umma_desc(sbase + kOffAhi + 2 * s * kPnLbo, kPnLbo, 128);
                                const uint64_t x_lo = umma_desc(sbase + kOffAlo + 2 * s * kPnLbo, kPnLbo, 128);
                                const uint32_t bst = sbase + kOffRing + slot * kSlot + s * step_bytes;
                                const uint64_t w_hi = umma_desc(bst, n * 8, 128);
                                const uint64_t w_lo = umma_desc(bst + n * 16, n * 8, 128);
                                umma2(tmem, x_hi, w_hi, idesc, s > 0 ? 1u : 0u);
                                umma2(tmem, x_lo, w_hi, idesc, 1u);
                                umma2(tmem, x_hi, w_lo, idesc, 1u);
                            }
                            tc_commit2(bar_empty + 8 * slot);
                        }
                        __syncwarp();
                        if (++slot == kStages) {
                            slot = 0;
                            phase ^= 1;
                        }
                    } else {
                        // transposed: D^T[256 features, 256 rows] = W[features, 128] . X[rows, 128]^T, 2 k16 steps per ring slot
                        const uint32_t idesc = umma_idesc2(256);
                        for (int s0 = 0; s0 < 8; s0 += 2) {
                            if (s0 > 0) {
                                mbar_wait_cluster(bar_full + 8 * slot, phase);
                                tc_fence_after();
                            }
                            if (elect_one()) {
#pragma unroll
                                for (int sub = 0; sub < 2; ++sub) {
                                    const int s = s0 + sub;
                                    const uint64_t x_hi = umma_desc(sbase + kOffAhi + 2 * s * kPnLbo, kPnLbo, 128);
                                    const uint64_t x_lo = umma_desc(sbase + kOffAlo + 2 * s * kPnLbo, kPnLbo, 128);
                                    const uint32_t wst = sbase + kOffRing + slot * kSlot + sub * 8192;
                                    const uint64_t w_hi = umma_desc(wst, 128 * 16, 128);
                                    const uint64_t w_lo = umma_desc(wst + 4096, 128 * 16, 128);
                                    umma2(tmem, w_hi, x_hi, idesc, s > 0 ? 1u : 0u);
                                    umma2(tmem, w_hi, x_lo, idesc, 1u);
                                    umma2(tmem, w_lo, x_hi, idesc, 1u);
                                }
                                tc_commit2(bar_empty + 8 * slot);
                            }
                            __syncwarp();
                            if (++slot == kStages) {
                                slot = 0;
                                phase ^= 1;
                            }
                        }
                    }
                    if (elect_one()) tc_commit2(bar_accum);  // accumulator of this layer complete, in both CTAs
                    __syncwarp();
                }
            }
        } else if (lane == 0) {
            // ---- peer: tell the leader when my half of a slot is here
            uint32_t slot = 0, phase = 0;
            for (long long n = 0; n < iters * kFills; ++n) {
                mbar_wait(bar_full + 8 * slot, phase);
                mbar_arrive_cluster(lead_full + 8 * slot);
                if (++slot == kStages) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        const int ew = warp - 2, et = tid - 64;
        const int lane_grp = (threadIdx.x >> 5) & 3, half = ew >> 2;  // the TMEM lane group follows the warp's index in the CTA
        const int row = lane_grp * 32 + lane;  // TMEM lane of this thread
        uint32_t accum_phase = 0;
        for (long long it = 0; it < iters; ++it) {
            const long long tile = 2 * (pair + it * npairs) + crank;  // a tile >= ntiles has no valid row
            // ---- gather + conv0a (SIMT, K=3): thread = (row, half of the 64 channels).  (Fetching the next tile's point a whole tile
            // ahead into three registers was measured and is slower: 52.6 vs 49.8 ms per 131^3 grid.)
            {
                const int r = et & 127, hf = et >> 7;
                const PnRow pr = pn_row(tile, r, G);
                const long long q = pr.q;
                const int p = pr.p;
                const bool valid = q < nq && p < P;
                float x = 0.f, y = 0.f, z = 0.f;
                if (valid) {
                    const float* src = patches + (q * P + p) * 3;
                    x = src[0];
                    y = src[1];
                    z = src[2];
                }
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    float v[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int ch = hf * 32 + kb * 8 + c;
                        v[c] = valid ? fmaxf(s_w0a[3 * ch] * x + s_w0a[3 * ch + 1] * y + s_w0a[3 * ch + 2] * z + s_b0a[ch], 0.f) : 0.f;
                    }
                    uint4 hi, lo;
                    split8(v, hi, lo);
                    *reinterpret_cast<uint4*>(smem + kOffAhi + (hf * 4 + kb) * kPnLbo + r * 16) = hi;
                    *reinterpret_cast<uint4*>(smem + kOffAlo + (hf * 4 + kb) * kPnLbo + r * 16) = lo;
                }
            }
            warp_arrive_cluster(lead_aready, lane);
            if ((et & 7) == 0 && it + 1 < iters) {  // the next tile's patch points stream from HBM: one L2 prefetch per 8 rows (96 bytes)
                const PnRow pn = pn_row(tile + 2 * npairs, et & 127, G);
                if (pn.q < nq && pn.p < P) prefetch_l2(patches + (pn.q * P + pn.p) * 3);
            }

            const PnRow prow = pn_row(tile, row, G);
            const bool row_valid = prow.q < nq && prow.p < P;
            // ---- conv0b / stn.conv1 (64 wide) and stn.conv2 (128 wide): bias + ReLU -> operand tile (in place)
#pragma unroll
            for (int layer = 0; layer < 3; ++layer) {
                mbar_wait(bar_accum, accum_phase);
                accum_phase ^= 1;
                tc_fence_after();
                const float* bias = layer == 0 ? s_b0b : (layer == 1 ? s_bs1 : s_bs2);
                const int nload = layer < 2 ? 1 : 2;  // 32 or 64 columns per thread
                float v[32];
#pragma unroll
                for (int cb = 0; cb < nload; ++cb) {
                    const int col0 = (layer < 2 ? half * 32 : half * 64) + cb * 32;
                    tmem_ld32(tmem + ((uint32_t)(lane_grp * 32) << 16) + col0, v);
                    bias_relu32(v, bias + col0);  // padding rows carry finite values nobody reads
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        float x8[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) x8[c] = v[kb * 8 + c];
                        uint4 hi, lo;
                        split8(x8, hi, lo);
                        const int kblk = (col0 >> 3) + kb;
                        *reinterpret_cast<uint4*>(smem + kOffAhi + kblk * kPnLbo + row * 16) = hi;
                        *reinterpret_cast<uint4*>(smem + kOffAlo + kblk * kPnLbo + row * 16) = lo;
                    }
                }
                warp_arrive_cluster(lead_aready, lane);
                if (layer == 0 && row_valid) {
                    // a1 feeds the feature transform of pn_feat_kernel -- stored AFTER the hand-off, while stn.conv1's MMAs run.  Global
                    // layout = the operand tile's: [tile][k8 block][row][8 floats], so the 32 lanes (= 32 consecutive rows) of a store
                    // instruction cover 1 KB instead of 32 different 256-byte rows (the row-major layout made this epilogue
                    // LSU-wavefront bound)
                    const int col0 = half * 32;
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        float4* dst = reinterpret_cast<float4*>(a1_out + ((tile * 8 + (col0 >> 3) + kb) * 128 + row) * 8);
                        dst[0] = make_float4(v[8 * kb], v[8 * kb + 1], v[8 * kb + 2], v[8 * kb + 3]);
                        dst[1] = make_float4(v[8 * kb + 4], v[8 * kb + 5], v[8 * kb + 6], v[8 * kb + 7]);
                    }
                }
            }
            // ---- stn.conv3 transposed: TMEM lane = feature 128 * crank + row, columns = the 256 rows of the pair-tile (tile of rank 0,
            // then tile of rank 1); max over the points of each of the four half-tiles.  This warp takes the tile of rank `half`
            mbar_wait(bar_accum, accum_phase);
            accum_phase ^= 1;
            tc_fence_after();
            {
                const long long ctile = 2 * (pair + it * npairs) + half;
                const int f = 128 * (int)crank + row;
                const float bias3 = s_bs3[f];
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    const PnRow ph = pn_row(ctile, hh * 64, G);  // this column block = one half-tile of one query
                    float m = -INFINITY;
                    for (int cb = 0; cb < 2; ++cb) {
                        float v[32];
                        tmem_ld32(tmem + ((uint32_t)(lane_grp * 32) << 16) + half * 128 + hh * 64 + cb * 32, v);
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (ph.p + cb * 32 + c < P) m = fmaxf(m, v[c]);
                    }
                    if (ph.q < nq) {
                        const float val = fmaxf(m + bias3, 0.f);  // ReLU and max commute
                        if (G == 1)
                            g_out[ph.q * 256 + f] = val;
                        else  // the patch spans several half-tiles: max over them (val >= 0, g zeroed by the host: int order = float order)
                            atomicMax(reinterpret_cast<int*>(g_out + ph.q * 256 + f), __float_as_int(val));
                    }
                }
            }
            tc_fence_before();
            // the next tile's gather overwrites the operand tile: the MMAs that read it are complete (accum barrier)
        }
    }
    tc_fence_before();
    __syncthreads();
    __syncwarp();
    cluster_sync_all();  // neither CTA exits (or frees its TMEM) while the peer may still arrive on its barriers or read its tile
    if (threadIdx.x >> 5 == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel C: feature transform, conv1, conv2, attention pooling
// ---------------------------------------------------------------------------------------------------------------------
namespace feat {
// Three CTAs per SM (round 2: two; the kernel is a serial chain load -> conv1 -> conv2 -> pooling per tile and bound by its
// latencies, so what counts is the number of chains in flight and the number of hand-offs in a chain).
// The feature transform is not a layer of its own: conv1(T_q . a1) = (W1 T_q) . a1, and W1 T_q is LINEAR in the output of the STN's
// last FC, so stn_fc_tc_kernel's fc3 is packed with the merged weights (packing.py) and emits M_q = W1 (T_q) directly -- this kernel's
// first MMA is conv1 with a per-query weight matrix (one MMA layer, one epilogue and two hand-offs less per tile than transform + conv1).
// Shared memory per CTA: the 64-column operand tile (33 KB) and ONE more 33 KB region R that holds the per-query matrices until their
// MMAs are done and the conv2 weights (32 KB) afterwards -- no weight ring: the copy lands while the epilogue converts conv1's output.
// One control warp issues the MMAs and the copy (it waits for its own commits), eight epilogue warps: 288 threads, <= 72 registers.
constexpr int kThreads = 288;
constexpr int kEpiThreads = 256;
constexpr int kOffAhi = 0;
constexpr int kABytes = 8 * kPnLbo;                       // 64-column operand tile (a1, x', conv1 output)
constexpr int kOffAlo = kOffAhi + kABytes;                // 16512
constexpr int kTLbo = 64 * 16 + 16;                       // 1040: k8-block pitch of a per-query 64x64 transform
constexpr int kTBytes = 8 * kTLbo;                        // 8320 per (query, hi/lo)
constexpr int kOffT = kOffAlo + kABytes;                  // 33024: region R: [query][hi,lo] conv1 matrices M_q, then conv2's weights
constexpr int kW2Offset = 16384;                          // conv2 in the weight pack (behind the unmerged conv1, which this kernel no longer reads)
constexpr int kW2Bytes = 32768;                           // conv2: 4 k16 steps of 8 KB
static_assert(kW2Bytes <= 4 * kTBytes, "conv2 must fit into the region of the per-query matrices");
constexpr int kOffPar = kOffT + 4 * kTBytes;              // 66304: b1[64] b2[128] wq[128] part[2][128] pool[4][128]
constexpr int kParFloats = 64 + 128 + 128 + 256 + 512;
constexpr int kOffBar = kOffPar + kParFloats * 4;         // full, (unused), accum, aready
constexpr int kOffTmem = kOffBar + 4 * 8;
constexpr int kSmemBytes = kOffTmem + 16 + 1024;          // ~70 KB
constexpr int kTmemCols = 128;
constexpr int kCtasPerSm = 3;
static_assert(kCtasPerSm * (kSmemBytes + 1024) <= 233472, "three CTAs per SM");
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
}  // namespace feat

template <bool MULTI>
__global__ void __launch_bounds__(feat::kThreads, feat::kCtasPerSm)
    pn_feat_kernel(const float* __restrict__ a1, const float* __restrict__ tmat, long long nq, int P, int G_arg,
                   const uint8_t* __restrict__ wpack, const float* __restrict__ b1, const float* __restrict__ b2,
                   const float* __restrict__ wq, float* __restrict__ pooled, float* __restrict__ partial) {
    using namespace feat;
    const int G = MULTI ? G_arg : 1;
    extern __shared__ __align__(1024) uint8_t smem[];  // used directly: the compiler keeps the shared address space (LDS/STS)
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* s_b1 = reinterpret_cast<float*>(smem + kOffPar);
    float* s_b2 = s_b1 + 64;
    float* s_wq = s_b2 + 128;
    float* s_part = s_wq + 128;    // [2][128] partial attention logits of the two column halves
    float* s_pool = s_part + 256;  // [4 lane groups][128] partial pooled sums of the warps' 32 rows
    volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + kOffTmem);
    const uint32_t bar_full = sbase + kOffBar, bar_accum = bar_full + 16, bar_aready = bar_accum + 8;

    for (int e = tid; e < 128; e += kThreads) {
        if (e < 64) s_b1[e] = b1[e];
        s_b2[e] = b2[e];
        s_wq[e] = wq[e];
    }
    if (tid == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_full + 8, 1);
        mbar_init(bar_accum, 1);
        mbar_init(bar_aready, kEpiThreads / 32);  // one elected arrival per warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kOffTmem), "r"((uint32_t)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const long long ntiles = (nq * G + 1) / 2;

    if (warp == 0) {
        // ---- control warp: all lanes run the loop (warp-uniform control flow), one elected lane issues (see elect_one)
        uint32_t ready_phase = 0, accum_phase = 0, full_phase = 0;
        const uint32_t w_r = sbase + kOffT;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            // conv1 with the per-query matrices: D[rows, ql*64 + o] = a1[rows, :] . M_ql[o, :]   (operands built by the epilogue warps;
            // a row reads the block of ITS query afterwards)
            mbar_wait(bar_aready, ready_phase);
            ready_phase ^= 1;
            tc_fence_after();
            if (elect_one()) {
                const uint32_t idesc = umma_idesc(64);
#pragma unroll
                for (int ql = 0; ql < 2; ++ql) {
                    const uint32_t t_hi = sbase + kOffT + (2 * ql) * kTBytes, t_lo = t_hi + kTBytes;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const uint64_t x_hi = umma_desc(sbase + kOffAhi + 2 * s * kPnLbo, kPnLbo, 128);
                        const uint64_t x_lo = umma_desc(sbase + kOffAlo + 2 * s * kPnLbo, kPnLbo, 128);
                        const uint64_t w_hi = umma_desc(t_hi + 2 * s * kTLbo, kTLbo, 128);
                        const uint64_t w_lo = umma_desc(t_lo + 2 * s * kTLbo, kTLbo, 128);
                        umma(tmem + ql * 64, x_hi, w_hi, idesc, s > 0 ? 1u : 0u);
                        umma(tmem + ql * 64, x_lo, w_hi, idesc, 1u);
                        umma(tmem + ql * 64, x_hi, w_lo, idesc, 1u);
                    }
                }
                tc_commit(bar_accum);
            }
            __syncwarp();
            // the per-query matrices have been read: conv2's weights take their place (they land while the epilogue converts conv1)
            mbar_wait(bar_accum, accum_phase);
            accum_phase ^= 1;
            if (elect_one()) {
                mbar_expect_tx(bar_full, kW2Bytes);
                bulk_copy(w_r, wpack + kW2Offset, kW2Bytes, bar_full);
            }
            __syncwarp();
            // conv2 (128 wide): 4 k16 steps of 8 KB.  The weights before the operand tile (see pn_stn_kernel)
            mbar_wait(bar_full, full_phase);
            full_phase ^= 1;
            mbar_wait(bar_aready, ready_phase);
            ready_phase ^= 1;
            tc_fence_after();
            if (elect_one()) {
                const uint32_t idesc = umma_idesc(128);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const uint64_t x_hi = umma_desc(sbase + kOffAhi + 2 * s * kPnLbo, kPnLbo, 128);
                    const uint64_t x_lo = umma_desc(sbase + kOffAlo + 2 * s * kPnLbo, kPnLbo, 128);
                    const uint32_t bst = w_r + s * 8192;
                    const uint64_t w_hi = umma_desc(bst, 128 * 16, 128);
                    const uint64_t w_lo = umma_desc(bst + 128 * 32, 128 * 16, 128);
                    umma(tmem, x_hi, w_hi, idesc, s > 0 ? 1u : 0u);
                    umma(tmem, x_lo, w_hi, idesc, 1u);
                    umma(tmem, x_hi, w_lo, idesc, 1u);
                }
                tc_commit(bar_accum);
            }
            __syncwarp();
            // second completion of the tile: keeps this warp's phase in step (the epilogue warps hold R until they have seen it)
            mbar_wait(bar_accum, accum_phase);
            accum_phase ^= 1;
        }
    } else {
        const int ew = warp - 1, et = tid - 32;  // epilogue warps 1..8: TMEM lane groups 1,2,3,0,1,2,3,0
        const int lane_grp = warp & 3, half = ew >> 2;
        const int row = lane_grp * 32 + lane;
        uint32_t accum_phase = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            // ---- load a1 (tile-major [tile][k8 block][row][8 floats], written by pn_stn_kernel): warp ew takes k8 block ew,
            // lanes are rows -> 1 KB per load instruction, conflict-free 512-byte shared-memory stores
#pragma unroll 2
            for (int i = 0; i < 4; ++i) {
                const int r = i * 32 + lane, kb = ew;
                const PnRow pr = pn_row(tile, r, G);
                const long long q = pr.q;
                const int p = pr.p;
                float v[8];
                if (q < nq && p < P) {
                    const float4* src = reinterpret_cast<const float4*>(a1 + ((tile * 8 + kb) * 128 + r) * 8);
                    const float4 u0 = src[0], u1 = src[1];
                    v[0] = u0.x; v[1] = u0.y; v[2] = u0.z; v[3] = u0.w;
                    v[4] = u1.x; v[5] = u1.y; v[6] = u1.z; v[7] = u1.w;
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[c] = 0.f;
                }
                uint4 hi, lo;
                split8(v, hi, lo);
                *reinterpret_cast<uint4*>(smem + kOffAhi + kb * kPnLbo + r * 16) = hi;
                *reinterpret_cast<uint4*>(smem + kOffAlo + kb * kPnLbo + r * 16) = lo;
            }
            // ---- per-query conv1 matrices M_q = W1 T_q [o][j] (written by stn_fc_tc_kernel) -> K-major operand (row o, k = j)
#pragma unroll 2
            for (int t = 0; t < 4; ++t) {
                const int e = et + 256 * t;
                const int ql = e >> 9, i = (e >> 3) & 63, kb = e & 7;
                long long q = G == 1 ? 2 * tile + ql : (2 * tile + ql) / G;  // the query of half-tile ql
                q = q < nq ? q : nq - 1;
                const float4* src = reinterpret_cast<const float4*>(tmat + q * 4096 + i * 64) + 2 * kb;
                const float4 u0 = src[0], u1 = src[1];
                const float v[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                uint4 hi, lo;
                split8(v, hi, lo);
                *reinterpret_cast<uint4*>(smem + kOffT + (2 * ql) * kTBytes + kb * kTLbo + i * 16) = hi;
                *reinterpret_cast<uint4*>(smem + kOffT + (2 * ql + 1) * kTBytes + kb * kTLbo + i * 16) = lo;
            }
            warp_arrive(bar_aready, lane);
            // the next tile's a1 rows and transforms come from HBM (written by the kernels before): pull them into the L2 now, one
            // 32-byte sector per load instruction of the loops above
            {
                const long long nt = tile + gridDim.x;
                if (nt < ntiles) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) prefetch_l2(a1 + ((nt * 8 + ew) * 128 + i * 32 + lane) * 8);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int e = et + 256 * t;
                        const int ql = e >> 9, i = (e >> 3) & 63, kb = e & 7;
                        long long q = G == 1 ? 2 * nt + ql : (2 * nt + ql) / G;
                        q = q < nq ? q : nq - 1;
                        prefetch_l2(tmat + q * 4096 + i * 64 + 8 * kb);
                    }
                }
            }

            // ---- conv1 (with the feature transform folded into its per-query matrix): bias + ReLU -> operand tile.  A row reads the
            // accumulator block of its own query
            mbar_wait(bar_accum, accum_phase);
            accum_phase ^= 1;
            tc_fence_after();
            {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(lane_grp * 32) << 16) + (row >> 6) * 64 + half * 32, v);
                bias_relu32(v, s_b1 + half * 32);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    float x8[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) x8[c] = v[kb * 8 + c];
                    uint4 hi, lo;
                    split8(x8, hi, lo);
                    *reinterpret_cast<uint4*>(smem + kOffAhi + (half * 4 + kb) * kPnLbo + row * 16) = hi;
                    *reinterpret_cast<uint4*>(smem + kOffAlo + (half * 4 + kb) * kPnLbo + row * 16) = lo;
                }
            }
            warp_arrive(bar_aready, lane);
            // ---- conv2: bias + ReLU; this thread's share (its 64 columns) of the attention logit of its row.  The activations stay in
            // TMEM: the pooling below reads the accumulator a second time instead of a shared-memory copy
            mbar_wait(bar_accum, accum_phase);
            accum_phase ^= 1;
            tc_fence_after();
            {
                float logit = 0.f;
#pragma unroll 1
                for (int cb = 0; cb < 2; ++cb) {
                    const int col0 = half * 64 + cb * 32;
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)(lane_grp * 32) << 16) + col0, v);
                    bias_relu32(v, s_b2 + col0);
#pragma unroll
                    for (int c = 0; c < 32; ++c) logit = fmaf(v[c], s_wq[col0 + c], logit);
                }
                s_part[half * 128 + row] = logit;
            }
            epi_barrier();
            // ---- softmax over the points of the half-tile (64 rows = the lane groups {0,1} or {2,3}): every warp reduces the 64 logits
            // on its own, lane l holds rows l and 32 + l (the query bias cancels in the softmax)
            float att;
            {
                const int r0 = (row >> 6) * 64;
                const PnRow ph = pn_row(tile, r0, G);       // first row of this half-tile: its query and first point
                const int nv = min(64, max(P - ph.p, 0));  // valid points of the half-tile
                const float la = lane < nv ? s_part[r0 + lane] + s_part[128 + r0 + lane] : -INFINITY;
                const float lb = 32 + lane < nv ? s_part[r0 + 32 + lane] + s_part[128 + r0 + 32 + lane] : -INFINITY;
                float m = fmaxf(la, lb);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                const float ea = lane < nv ? expf(la - m) : 0.f, eb = 32 + lane < nv ? expf(lb - m) : 0.f;
                float sum = ea + eb;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const float own = (lane_grp & 1) ? eb : ea;
                // one half-tile per query: normalise here.  Several: keep exp(l - m_h), the combine kernel merges the half-tiles'
                // (m_h, sum_h, pooled_h) like an online softmax
                att = nv > 0 ? (G == 1 ? own / sum : own) : 0.f;
                if (G > 1 && half == 0 && (row & 63) == 0 && ph.q < nq) {
                    float* dst = partial + (ph.q * G + ph.p / 64) * kPartialStride;
                    dst[128] = m;
                    dst[129] = sum;
                }
            }
            // ---- pooled[q, c] = sum_p att_p c2[p, c]: every thread scales its row of the accumulator, the 32 rows of a warp are summed
            // by recursive halving (lane l ends with column l of the block), the two warps of a half-tile meet in shared memory
#pragma unroll 1
            for (int cb = 0; cb < 2; ++cb) {
                const int col0 = half * 64 + cb * 32;
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(lane_grp * 32) << 16) + col0, v);
                bias_relu32(v, s_b2 + col0);
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] *= att;
#pragma unroll
                for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
                    const bool upper = (lane & off) != 0;
#pragma unroll
                    for (int i = 0; i < n; ++i) {
                        const float send = upper ? v[i] : v[i + n];
                        const float keep = upper ? v[i + n] : v[i];
                        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                s_pool[lane_grp * 128 + col0 + lane] = v[0];
            }
            tc_fence_before();
            epi_barrier();
            {
                const int ql = et >> 7, c = et & 127;
                const PnRow ph = pn_row(tile, ql * 64, G);
                if (ph.q < nq) {
                    const float val = s_pool[(2 * ql) * 128 + c] + s_pool[(2 * ql + 1) * 128 + c];
                    if (G == 1)
                        pooled[ph.q * 128 + c] = val;
                    else
                        partial[(ph.q * G + ph.p / 64) * kPartialStride + c] = val;
                }
            }
            // s_part / s_pool are next written behind mbarrier waits that need an arrival of every epilogue warp, i.e. after this point
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
    }
}

// merges the G half-tile partials of a query: pooled = sum_h e^(m_h - M) pooled_h / sum_h e^(m_h - M) sum_h,  M = max_h m_h
__global__ void pn_combine_kernel(const float* __restrict__ partial, long long nq, int G, float* __restrict__ pooled) {
    const long long q = blockIdx.x;
    const int c = threadIdx.x;  // 128 channels
    if (q >= nq) return;
    const float* base = partial + q * G * kPartialStride;
    float M = -INFINITY;
    for (int h = 0; h < G; ++h) M = fmaxf(M, base[h * kPartialStride + 128]);
    float num = 0.f, den = 0.f;
    for (int h = 0; h < G; ++h) {
        const float m = base[h * kPartialStride + 128];
        if (m == -INFINITY) continue;  // a half-tile without valid points
        const float w = expf(m - M);
        num = fmaf(w, base[h * kPartialStride + c], num);
        den = fmaf(w, base[h * kPartialStride + 129], den);
    }
    pooled[q * 128 + c] = num / den;
}

}  // namespace tc

int linear_impl(const float* x, const float* w, const float* bias, const float* residual, const int32_t* gather,
                float* y, int64_t m, int n, int k, int ldx, int ldy, int act, cudaStream_t st);
bool chain_tc_supported(const pps_decoder_weights* w);
int stn_fc_tc_impl(const pps_decoder_weights* w, const float* g, int64_t q, float* tmat, cudaStream_t st);

bool pointnet_tc_supported(const pps_decoder_weights* w) {
    // tc_stn_fc: pn_feat_kernel expects the MERGED matrices M_q = W1 T_q that only stn_fc_tc_kernel's pack produces
    return w->tc_pn_stn != nullptr && w->tc_pn_feat != nullptr && w->tc_stn_fc != nullptr && w->num_pts_local <= 256 &&
           w->stn_size == 256 && w->latent == 256;
}

// local branch on the tensor cores: patches [q,P,3] -> pooled128 [q,128]; scratch: a1 [tiles,8,128,8] (tile-major, 64 point slots per query), g [q,256], f1 [q,128],
// f2 [q,64], tmat [q,4096]
int pointnet_tc_impl(const pps_decoder_weights* w, const float* patches, int64_t q, float* a1, float* g, float* f1, float* f2,
                     float* tmat, float* pooled128, float* partial, cudaStream_t st) {
    if (q == 0) return PPS_OK;
    static unsigned char configured[kMaxDevices] = {};
    if (first_use_on_device(configured)) {
        PPS_CUDA(cudaFuncSetAttribute(tc::pn_stn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::stn::kSmemBytes));
        PPS_CUDA(cudaFuncSetAttribute(tc::pn_stn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::stn::kSmemBytes));
        PPS_CUDA(cudaFuncSetAttribute(tc::pn_feat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::feat::kSmemBytes));
        PPS_CUDA(cudaFuncSetAttribute(tc::pn_feat_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::feat::kSmemBytes));
    }
    const int P = w->num_pts_local;
    const int G = (P + 63) / 64;  // 64-row half-tiles per query
    const long long ntiles = (q * G + 1) / 2;
    // pn_stn: CTA pairs, two chains per CTA -> up to 74 clusters x 2 chains, a pair of chains takes a pair-tile of 2 tiles
    const long long npt = (ntiles + 1) / 2;
    const int clusters = (int)((npt + 1) / 2 < kNumSMs / 2 ? (npt + 1) / 2 : kNumSMs / 2);
    if (G > 1) PPS_CUDA(cudaMemsetAsync(g, 0, (size_t)q * 256 * sizeof(float), st));  // the half-tiles' maxima meet in an atomicMax
    const uint8_t* pack_stn = static_cast<const uint8_t*>(w->tc_pn_stn);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(2 * tc::kPnThreads);
        cfg.dynamicSmemBytes = tc::stn::kSmemBytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;  // CTA pairs: the two CTAs of a cluster sit on the two SMs of one TPC
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const long long nq = q;
        if (G > 1)
            PPS_CUDA(cudaLaunchKernelEx(&cfg, tc::pn_stn_kernel<true>, patches, nq, P, G, pack_stn, (const float*)w->pn0a_w,
                                        (const float*)w->pn0a_b, (const float*)w->pn0b_b, (const float*)w->stn1_b,
                                        (const float*)w->stn2_b, (const float*)w->stn3_b, a1, g));
        else
            PPS_CUDA(cudaLaunchKernelEx(&cfg, tc::pn_stn_kernel<false>, patches, nq, P, 1, pack_stn, (const float*)w->pn0a_w,
                                        (const float*)w->pn0a_b, (const float*)w->pn0b_b, (const float*)w->stn1_b,
                                        (const float*)w->stn2_b, (const float*)w->stn3_b, a1, g));
    }
    PPS_LAUNCH_CHECK();
    PPS_CHECK_ARG(chain_tc_supported(w), "tensor-core PointNet path without the STN FC pack (tc_stn_fc)");
    PPS_TRY(stn_fc_tc_impl(w, g, q, tmat, st));  // tmat = the per-query conv1 matrices M_q = W1 T_q (merged pack)
    const uint8_t* pack_feat = static_cast<const uint8_t*>(w->tc_pn_feat);
    const int grid_feat = (int)(ntiles < tc::feat::kCtasPerSm * kNumSMs ? ntiles : tc::feat::kCtasPerSm * kNumSMs);
    if (G > 1)
        tc::pn_feat_kernel<true><<<grid_feat, tc::feat::kThreads, tc::feat::kSmemBytes, st>>>(a1, tmat, q, P, G, pack_feat, w->pn1_b,
                                                                                             w->pn2_b, w->pnq_w, pooled128, partial);
    else
        tc::pn_feat_kernel<false><<<grid_feat, tc::feat::kThreads, tc::feat::kSmemBytes, st>>>(a1, tmat, q, P, 1, pack_feat, w->pn1_b,
                                                                                              w->pn2_b, w->pnq_w, pooled128, partial);
    PPS_LAUNCH_CHECK();
    if (G > 1) {
        tc::pn_combine_kernel<<<(unsigned)q, 128, 0, st>>>(partial, q, G, pooled128);
        PPS_LAUNCH_CHECK();
    }
    return PPS_OK;
}

// floats of the per-(query, half-tile) attention-pooling partials (0 when every patch fits one half-tile)
size_t pointnet_tc_partial_floats(const pps_decoder_weights* w, int64_t q) {
    const int G = (w->num_pts_local + 63) / 64;
    return G > 1 ? (size_t)q * G * tc::kPartialStride : 0;
}

}  // namespace pps

extern "C" size_t pps_decoder_tc_pn_stn_bytes(void) { return pps::tc::kPackStnBytes; }
extern "C" size_t pps_decoder_tc_pn_feat_bytes(void) { return pps::tc::kPackFeatBytes; }
