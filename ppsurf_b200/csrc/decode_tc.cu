// Tensor-core path of the decoder's global branch (path 1): ONE persistent, warp-specialised tcgen05 kernel per chunk.
//
// Replaces the per-(query,neighbour) part of InterpAttentionKHeadsNet.forward (source/poco_model.py:400-414): gather of
// the hoisted fc1 table, fc2, fc3, fc_query, softmax over the 64 neighbours, mean over the 64 heads and the
// attention-weighted pooling, for tiles of 128 rows = 2 queries x 64 neighbours.  Nothing but the pooled [q,256] vectors
// leaves the SM (the unfused path moves ~0.5 MB per query through HBM).
//
// Precision: the reference is fp32 and the contract is 1e-4 abs on the logits, so every GEMM runs as a split-fp16
// product on the 5th-gen tensor cores:  x = x_hi + x_lo (two fp16, 22 mantissa bits), W = W_hi + W_lo,
//   x.W ~= x_hi.W_hi + x_lo.W_hi + x_hi.W_lo   (fp32 accumulation in TMEM),
// i.e. three tcgen05.mma per k-step; the dropped x_lo.W_lo term is O(2^-22) relative.
//
// Roles, the CTA-pair scheme and the per-tile pipeline are described above projection_tc_kernel.
#include "tc_common.cuh"

namespace pps {
namespace tc {

constexpr int kRows = 128;                  // rows per tile
constexpr int kC = 256;                     // latent width
constexpr int kHeads = 64;
constexpr int kNbrs = 64;                   // neighbours per query
constexpr int kKB = kC / 8;                 // k8 blocks per row
constexpr int kALbo = kRows * 16 + 16;      // bytes between k8 blocks of an activation tile (padded)
constexpr int kABytes = kKB * kALbo;        // 66048
constexpr int kStageBytes = 8192;           // ring slot = what ONE CTA of the pair holds of a k16 step: 128 weight rows, hi 4 KB + lo 4 KB
                                            // (fc2 / fc3: this CTA's half of the 256 output features; fc_query as the M operand:
                                            // 64 zero rows + 64 heads, the same in both CTAs)
constexpr int kSub = 1;                      // k16 steps per ring slot.  2 was measured (one full-wait + one commit per two steps): the
                                            // one-term ablation kernel gains 20 %, the product kernel LOSES 11 % (30.9 k vs 27.7 k cycles per
                                            // tile: coarser slot release, later refills) -- the product is bound by MMA execution, not by the loop
constexpr int kSlotBytes = kSub * kStageBytes;
constexpr int kStages = 10 / kSub;
constexpr int kKSteps = kC / 16;            // 16 k16 steps per layer
// split-fp16 terms the product kernel issues, bit 3*layer + t (t = 0: x_hi w_hi, 1: x_lo w_hi, 2: x_hi w_lo): all three for every layer.
// fc_query ran with two terms for a while (on the 4096-query ablation set of profiles/r02_pass_ablation.md its weight-lo term moves the
// logits by 1e-6), but on the smoke workload a query with attention scores around 17 lost 2.1e-4 in the pooled vector and 1.2e-4 in a
// logit -- beyond the 1e-4 contract (tools/smoke_debug.py: 5.2e-5 with the term).  As an M=256 x N=64 product the term costs 16 small
// MMAs per tile.
constexpr uint32_t kProductTerms = 0x1FFu;
constexpr int kChunks = 4;                  // a layer's K range is released to the MMA warp in 4 chunks of 64 columns
constexpr int kGWarps = 8;                  // gather + fc2/fc3 epilogues
constexpr int kSWarps = 8;                  // softmax + attention pooling: two warps per TMEM lane group, 32 heads / 4 column blocks each
constexpr int kGThreads = kGWarps * 32;
constexpr int kSThreads = kSWarps * 32;
constexpr int kCtlWarps = 4;                // warpgroup 0: producer, MMA issuer / relay, two idle warps (setmaxnreg works per warpgroup)
constexpr int kThreads = kCtlWarps * 32 + kGThreads + kSThreads;  // 640 = 5 warpgroups
// Register budget per lane of an SM sub-partition (16384 registers = 512 per lane, 5 warps: one of every warpgroup): the kernel
// starts with 96 everywhere and redistributes with setmaxnreg: the control warpgroup releases 128 x 64 registers, the G group takes
// 256 x 32 (32 + 2 x 128 + 2 x 96 = 480 per lane); the S group keeps the launch allocation
constexpr int kCtlRegs = 32, kGRegs = 128;

// shared memory map (bytes from the base)
constexpr int kOffAhi = 0;
constexpr int kOffAlo = kOffAhi + kABytes;
constexpr int kOffRing = kOffAlo + kABytes;                  // 132096
constexpr int kOffBias = kOffRing + kStages * kSlotBytes;   // 214016: b2[256] b3[256] bq[64]
constexpr int kOffAttp = kOffBias + (256 + 256 + 64) * 4;    // 216320: [head half][query][64] partial head sums
constexpr int kOffPool = kOffAttp + 2 * 2 * 64 * 4;          // 217344: [lane group][256] partial pooled sums
constexpr int kOffW1 = kOffPool + 4 * 256 * 4;               // 221440: fc1 xyz weights [256][3]
constexpr int kOffVq = kOffW1 + 768 * 4;                     // 224512: W1_xyz . q for the tile's 2 queries [2][256]
constexpr int kOffBar = kOffVq + 2 * 256 * 4;                // 226560: full[10] empty[10] acc[3] chunk[4] d0free d1free
constexpr int kNumBars = 2 * kStages + 3 + kChunks + 2;
constexpr int kOffTmem = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmem + 16;

// packed weights in global memory: [fc2: 16 stages x (CTA 0 half, CTA 1 half)][fc3: likewise][fc_query: 4 slots x (CTA 0: heads 0..31,
// CTA 1: heads 32..63), a slot = the 4 k16 steps of one 64-column chunk, 2 KB each]
constexpr int kQSub = 4;                    // k16 steps of fc_query per ring slot (N = 64: 32 weight rows per CTA, 2 KB per step)
constexpr int kQStep = 2048;
constexpr size_t kPackBytes = size_t(2) * kKSteps * 2 * kStageBytes + size_t(kChunks) * 2 * kStageBytes;

__device__ __forceinline__ void g_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kGThreads) : "memory"); }
__device__ __forceinline__ void s_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(kSThreads) : "memory"); }

// Pipeline of one tile (2 queries x 64 neighbours):
//   gather -> fc2 (D0) -> E2 -> fc3 (D1) -> E3 -> fc_query (D0, columns 0..63) -> softmax / head mean / pooling from D1
// The kernel runs as CTA PAIRS (cluster of 2, tcgen05 cta_group::2): every MMA is M=256 = the two CTAs' 128-row tiles, each CTA
// holds only its half of the weight stage (B operand) and the pair shares it, which halves the shared-memory traffic and the
// L2 weight stream per row -- with one CTA per MMA the kernel is shared-memory-bandwidth bound (36 KB of operand reads + 16 KB
// of weight refill per 384 tensor cycles).  Rank 0 issues the MMAs of the pair; barriers that gate them live in rank 0 and
// collect arrivals from both CTAs; completions are multicast to both.
// Warp roles (512 threads, one persistent CTA per SM):
//   warp 0       weight producer: cp.async.bulk of this CTA's half of the pre-packed k16 weight stages into a 10-slot ring
//   warp 1       rank 0: MMA issuer (one thread); rank 1: relays "my half of the stage has landed" to rank 0
//   warps 2..9   G group: gather of the fc1 table rows (+ W1_xyz.q, ReLU, fp16 hi/lo split) and the fc2 / fc3 epilogues.  Every
//                epilogue rewrites the operand tile in place 64 columns at a time and releases each chunk through its own
//                mbarrier, so the next layer's MMAs (other accumulator) overlap it.
//   warps 10..15 S group: softmax over the neighbours, head mean and the attention pooling of tile t run while the G group and
//                the tensor pipe are already working on tile t+1 (D0 is handed back as soon as the scores are in registers,
//                D1 when the pooling has read fc3's accumulator).
// fc_query is a plain M=256 x N=64 product (heads on the accumulator columns; round 1 computed it transposed with the 64 heads
// padded to M=128 and both CTAs computing all 256 columns -- four times the MMA work of this form, and the kernel is bound by MMA
// execution).  The softmax over a query's 64 neighbours is then a reduction over TMEM lanes = over the 32 lanes of a warp (recursive
// halving: lane l ends with heads 2l, 2l+1) and over the two warps of the query (shared memory); the exponentials wait in TMEM.
// cycle counter of the instrumented build only: the product kernel (PROF = false) carries no clock reads -- the MMA issuer's loop
// is on the critical path (a stray branch in it cost 8 % of the kernel)
template <bool PROF>
__device__ __forceinline__ long long tick() {
    return PROF ? clock64() : 0ll;
}

// ABL = true (pass-ablation build, pps_decoder_tc_terms): bit 3*layer + t of term_mask enables split-fp16 term t of the layer
// (0 = x_hi w_hi, 1 = x_lo w_hi, 2 = x_hi w_lo); the product instantiation issues all three unconditionally
template <bool PROF, bool ABL>
__global__ void __launch_bounds__(kThreads, 1)
    projection_tc_kernel(const float* __restrict__ table, const float* __restrict__ queries, const int32_t* __restrict__ idx, int ks,
                         long long nq, const uint8_t* __restrict__ wpack, const float* __restrict__ b2, const float* __restrict__ b3,
                         const float* __restrict__ bq, const float* __restrict__ w1_xyz, float* __restrict__ pooled, long long* prof,
                         uint32_t term_mask) {
    extern __shared__ __align__(1024) uint8_t smem[];  // used directly: the compiler keeps the shared address space (LDS/STS)
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* s_bias = reinterpret_cast<float*>(smem + kOffBias);
    float* s_attp = reinterpret_cast<float*>(smem + kOffAttp);
    float* s_pool = reinterpret_cast<float*>(smem + kOffPool);
    float* s_w1 = reinterpret_cast<float*>(smem + kOffW1);
    float* s_vq = reinterpret_cast<float*>(smem + kOffVq);
    volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + kOffTmem);
    const uint32_t bar_full = sbase + kOffBar, bar_empty = bar_full + 8 * kStages, bar_acc = bar_empty + 8 * kStages,
                   bar_chunk = bar_acc + 8 * 3, bar_d0free = bar_chunk + 8 * kChunks, bar_d1free = bar_d0free + 8;
    const uint32_t crank = cluster_ctarank();  // 0 = leader of the pair
    // the barriers that gate the pair's MMAs, in the leader's shared memory (valid from either CTA)
    const uint32_t lead_full = map_to_cta(bar_full, 0), lead_chunk = map_to_cta(bar_chunk, 0),
                   lead_d0free = map_to_cta(bar_d0free, 0), lead_d1free = map_to_cta(bar_d1free, 0);

    for (int e = tid; e < 768; e += kThreads) {
        s_w1[e] = w1_xyz[e];
        if (e < 256) {
            s_bias[e] = b2[e];
            s_bias[256 + e] = b3[e];
        }
        if (e < 64) s_bias[512 + e] = bq[e];
    }
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar_full + 8 * i, crank == 0 ? 2 : 1);  // leader: own producer + the peer's relay
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 3; ++i) mbar_init(bar_acc + 8 * i, 1);  // accumulator of fc2 / fc3 / fc_query complete
        for (int i = 0; i < kChunks; ++i) mbar_init(bar_chunk + 8 * i, 2 * kGWarps);  // one elected arrival per G warp of the pair
        mbar_init(bar_d0free, 2 * kSWarps);  // the softmax warps of both CTAs hold the scores in registers
        mbar_init(bar_d1free, 2 * kSWarps);  // the pooling of both CTAs has read fc3's accumulator
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kOffTmem), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the barriers of both CTAs are initialised before any remote arrival or multicast commit
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    // tile = 2 queries x 64 neighbours; pair-tile = the two tiles of a CTA pair; both CTAs of a pair run the same number of
    // iterations (a tile past the end works on clamped queries and writes nothing)
    const long long npt = (nq + 3) / 4;
    const long long pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const long long iters = npt > pair ? (npt - pair + npairs - 1) / npairs : 0;
    const long long tile0 = 2 * pair + crank, tile_step = 2 * npairs;

    // registers follow the roles: the control warpgroup gives most of its allocation back, the G group takes it
    if (warp < kCtlWarps) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtlRegs));
    if (warp == 0) {
        // ---------------------------------------------------------------- weight producer (all lanes run the loop, one issues)
        {
            uint32_t slot = 0, phase = 0;
            for (long long it = 0; it < iters; ++it) {
                for (int layer = 0; layer < 3; ++layer) {
                    // fc2 / fc3: this CTA's half (128 features) of every k16 stage; fc_query: this CTA's 32 heads, one slot per chunk
                    const int nslots = layer < 2 ? kKSteps / kSub : kChunks;
                    const uint8_t* src = layer < 2 ? wpack + (size_t)layer * kKSteps * 2 * kStageBytes + crank * kStageBytes
                                                   : wpack + (size_t)2 * kKSteps * 2 * kStageBytes + crank * kStageBytes;
                    const uint32_t stride = 2 * kStageBytes;
                    for (int s = 0; s < nslots; ++s) {
                        mbar_wait(bar_empty + 8 * slot, phase ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(bar_full + 8 * slot, kSlotBytes);
#pragma unroll
                            for (int sub = 0; sub < kSub; ++sub)
                                bulk_copy(sbase + kOffRing + slot * kSlotBytes + sub * kStageBytes, src + sub * stride, kStageBytes,
                                          bar_full + 8 * slot);
                        }
                        __syncwarp();
                        src += kSub * stride;
                        if (++slot == kStages) {
                            slot = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (crank == 0) {
            // ---------------------------------------------------------------- MMA issuer of the pair: the whole warp runs the loop
            // (warp-uniform control flow), one elected lane issues the MMAs and commits (see elect_one)
            uint32_t slot = 0, phase = 0, chunk_phase = 0;
            long long t_chunk = 0, t_full = 0, t_dfree = 0, t_issue = 0, t_commit = 0, t_total = tick<PROF>();
            for (long long it = 0; it < iters; ++it) {
#pragma unroll
                for (int layer = 0; layer < 3; ++layer) {
                    const uint32_t idesc = umma_idesc2(256);
                    const uint32_t dcol = layer == 1 ? 256u : 0u;
                    if (it > 0 && layer < 2) {  // the S groups have taken what they need of the previous tiles out of this accumulator
                        const long long t0 = tick<PROF>();
                        mbar_wait_cluster(layer == 0 ? bar_d0free : bar_d1free, (uint32_t)((it - 1) & 1));
                        tc_fence_after();
                        t_dfree += tick<PROF>() - t0;
                    }
                    // product kernel: compile-time terms (the layer loop is unrolled); ablation kernel: the runtime mask
                    const uint32_t lm = ABL ? (term_mask >> (3 * layer)) & 7u : (kProductTerms >> (3 * layer)) & 7u;
                    if (layer < 2) {
                        for (int sl = 0; sl < kKSteps / kSub; ++sl) {  // one ring slot = kSub k16 steps: one full-wait and one commit per slot
                            // the weights first: they have usually landed long ago, and a wait on a completed mbarrier still costs ~90
                            // cycles that would otherwise sit between the operand chunk's release and the first MMA that reads it
                            long long t0 = tick<PROF>();
                            mbar_wait_cluster(bar_full + 8 * slot, phase);  // both halves of the weight stage have landed (TMA writes:
                                                                            // async proxy -> async proxy, no tcgen05 fence needed)
                            long long t1 = tick<PROF>();
                            if (((sl * kSub) & 3) == 0) {  // operand columns [64c, 64c+64) of BOTH tiles written by the previous stage of the pipeline
                                mbar_wait_cluster(bar_chunk + 8 * ((sl * kSub) >> 2), chunk_phase);
                                tc_fence_after();
                            }
                            const long long t2 = tick<PROF>();
                            t_full += t1 - t0;
                            t_chunk += t2 - t1;
                            if (elect_one()) {
#pragma unroll
                                for (int sub = 0; sub < kSub; ++sub) {
                                    const int s = sl * kSub + sub;
                                    const uint32_t a_off = 2 * s * kALbo;
                                    const uint64_t x_hi = umma_desc(sbase + kOffAhi + a_off, kALbo, 128);
                                    const uint64_t x_lo = umma_desc(sbase + kOffAlo + a_off, kALbo, 128);
                                    const uint32_t wst = sbase + kOffRing + slot * kSlotBytes + sub * kStageBytes;
                                    const uint64_t w_hi = umma_desc(wst, 128 * 16, 128);
                                    const uint64_t w_lo = umma_desc(wst + 4096, 128 * 16, 128);
                                    // D[256 rows, 256 features]: B = the two CTAs' 128-feature halves
                                    umma2(tmem + dcol, x_hi, w_hi, idesc, s > 0 ? 1u : 0u);
                                    if (lm & 2u) umma2(tmem + dcol, x_lo, w_hi, idesc, 1u);
                                    if (lm & 4u) umma2(tmem + dcol, x_hi, w_lo, idesc, 1u);
                                }
                                tc_commit2(bar_empty + 8 * slot);  // frees the ring slot of both CTAs when these MMAs have read it
                            }
                            __syncwarp();
                            t_issue += tick<PROF>() - t2;
                            if (++slot == kStages) {
                                slot = 0;
                                phase ^= 1;
                            }
                        }
                    } else {
                        // fc_query: scores[256 rows, 64 heads] at D0 columns 0..63, B = the two CTAs' 32-head halves; one ring slot per
                        // 64-column chunk (4 k16 steps of 2 KB)
                        const uint32_t idesc_q = umma_idesc2(kHeads);
                        for (int c = 0; c < kChunks; ++c) {
                            long long t0 = tick<PROF>();
                            mbar_wait_cluster(bar_full + 8 * slot, phase);
                            long long t1 = tick<PROF>();
                            mbar_wait_cluster(bar_chunk + 8 * c, chunk_phase);
                            tc_fence_after();
                            t_full += t1 - t0;
                            t_chunk += tick<PROF>() - t1;
                            if (elect_one()) {
#pragma unroll
                                for (int sub = 0; sub < kQSub; ++sub) {
                                    const int s = c * kQSub + sub;
                                    const uint32_t a_off = 2 * s * kALbo;
                                    const uint64_t x_hi = umma_desc(sbase + kOffAhi + a_off, kALbo, 128);
                                    const uint64_t x_lo = umma_desc(sbase + kOffAlo + a_off, kALbo, 128);
                                    const uint32_t wst = sbase + kOffRing + slot * kSlotBytes + sub * kQStep;
                                    const uint64_t w_hi = umma_desc(wst, 32 * 16, 128);
                                    const uint64_t w_lo = umma_desc(wst + kQStep / 2, 32 * 16, 128);
                                    umma2(tmem, x_hi, w_hi, idesc_q, s > 0 ? 1u : 0u);
                                    if (lm & 2u) umma2(tmem, x_lo, w_hi, idesc_q, 1u);
                                    if (lm & 4u) umma2(tmem, x_hi, w_lo, idesc_q, 1u);
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
                    if (elect_one()) tc_commit2(bar_acc + 8 * layer);  // accumulator of this layer complete, in both CTAs
                    __syncwarp();
                    chunk_phase ^= 1;
                }
            }
            if (PROF && prof && lane == 0) prof[32 + pair] = tick<PROF>() - t_total;  // every pair's total, to see the spread over the chip
            if (PROF && prof && blockIdx.x == 0 && lane == 0) {  // cycles: MMA warp total, waiting for operand chunks / weights / the S groups
                prof[0] = tick<PROF>() - t_total;
                prof[1] = t_chunk;
                prof[2] = t_full;
                prof[3] = t_dfree;
                prof[10] = t_issue;   // fc2 / fc3 slots only: the three tcgen05.mma of a k16 step and the commit that releases the slot
                prof[11] = t_commit;
            }
        } else if (lane == 0) {
            // ---------------------------------------------------------------- peer: tell the leader when my half of a stage is here
            uint32_t slot = 0, phase = 0;
            for (long long n = 0; n < iters * (2 * (kKSteps / kSub) + kChunks); ++n) {
                mbar_wait(bar_full + 8 * slot, phase);
                mbar_arrive_cluster(lead_full + 8 * slot);
                if (++slot == kStages) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
    }  // warps 2, 3: no role (they complete the control warpgroup)
    } else if (warp < kCtlWarps + kGWarps) {
        // ---------------------------------------------------------------- G group: gather + fc2 / fc3 epilogues
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kGRegs));
        const int ew = warp - kCtlWarps;       // 0..7
        const int gt = tid - kCtlWarps * 32;   // 0..255
        const int lane_grp = warp & 3;         // TMEM lanes this warp may touch: 32*lane_grp .. +31
        const int half = ew >> 2;              // which half of every 64-column chunk
        const int row = lane_grp * 32 + lane;  // accumulator row (= TMEM lane) of this thread
        const int kb = lane & 7, rs = lane >> 3;
        const int ql = (ew * 16) >> 6;         // the query whose rows this warp gathers
        long long t_gather = 0, t_wait = 0, t_epi = 0, t_mark = tick<PROF>();

        int src_next[4];
        {
            long long q = 2 * tile0 + ql;
            q = q < nq ? q : nq - 1;
#pragma unroll
            for (int i = 0; i < 4; ++i) src_next[i] = idx[q * ks + ((ew * 16 + 4 * i + rs) & 63)];
        }
        // W1_xyz . q for the two queries of a tile.  Computed for the NEXT tile while this group waits for fc3 (below): at the top of
        // the tile the query loads' latency would sit between fc_query's completion and the first chunk of the gather
        auto query_term = [&](long long tile) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                long long q = 2 * tile + t;
                q = q < nq ? q : nq - 1;
                s_vq[t * 256 + gt] = s_w1[3 * gt] * queries[3 * q] + s_w1[3 * gt + 1] * queries[3 * q + 1] +
                                     s_w1[3 * gt + 2] * queries[3 * q + 2];
            }
            g_barrier();
        };
        if (iters > 0) query_term(tile0);
        for (long long it = 0; it < iters; ++it) {
            const long long tile = tile0 + it * tile_step;
            // ---- gather: h1 = relu(U[idx] + W1_xyz.q) -> A_hi/A_lo, one 64-column chunk after the other so that fc2's first
            // k-steps start after a quarter of the gather.  Warp ew owns rows 16*ew .. +15 (all of one query); per chunk a lane
            // owns k8 block (lane & 7) of row 4*i + (lane >> 3): 8 lanes read 256 contiguous bytes of a table row
            {
                int src[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) src[i] = src_next[i];
                if (it + 1 < iters) {  // neighbour ids of the NEXT tile: their latency hides behind this tile
                    long long q = 2 * (tile + tile_step) + ql;
                    q = q < nq ? q : nq - 1;
#pragma unroll
                    for (int i = 0; i < 4; ++i) src_next[i] = idx[q * ks + ((ew * 16 + 4 * i + rs) & 63)];
                }
                float4 u[2][4][2];  // double-buffered in registers: chunk c+1 is in flight while chunk c is converted
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4* urow = reinterpret_cast<const float4*>(table + (size_t)src[i] * kC) + 2 * kb;
                    u[0][i][0] = urow[0];
                    u[0][i][1] = urow[1];
                }
                if (it > 0) {  // the operand tile is free when fc_query of the previous tile has read it
                    mbar_wait(bar_acc + 16, (uint32_t)((it - 1) & 1));
                }
                {
                    const long long now = tick<PROF>();
                    t_wait += now - t_mark;
                    t_mark = now;
                }
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    if (c + 1 < kChunks) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4* urow = reinterpret_cast<const float4*>(table + (size_t)src[i] * kC + (c + 1) * 64) + 2 * kb;
                            u[(c + 1) & 1][i][0] = urow[0];
                            u[(c + 1) & 1][i][1] = urow[1];
                        }
                    }
                    const float4 v0 = *reinterpret_cast<const float4*>(s_vq + ql * 256 + c * 64 + kb * 8);
                    const float4 v1 = *reinterpret_cast<const float4*>(s_vq + ql * 256 + c * 64 + kb * 8 + 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = ew * 16 + 4 * i + rs;
                        const float4 a = u[c & 1][i][0], b = u[c & 1][i][1];
                        const float raw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                        float x[8];
                        bias_relu8(x, raw, v0, v1);
                        uint4 hi, lo;
                        split8(x, hi, lo);
                        *reinterpret_cast<uint4*>(smem + kOffAhi + (c * 8 + kb) * kALbo + r * 16) = hi;
                        *reinterpret_cast<uint4*>(smem + kOffAlo + (c * 8 + kb) * kALbo + r * 16) = lo;
                    }
                    warp_arrive_cluster(lead_chunk + 8 * c, lane);
                }
            }
            {
                const long long now = tick<PROF>();
                t_gather += now - t_mark;
                t_mark = now;
            }

            // ---- fc2 / fc3 epilogues: D -> +bias, ReLU, split -> A in place, released chunk by chunk
            for (int layer = 0; layer < 2; ++layer) {
                // every G warp is past this tile's gather (fc2's accumulator is complete, which needs every warp's chunk arrivals),
                // so s_vq may be overwritten: the next tile's query term, in the shadow of fc3
                if (layer == 1 && it + 1 < iters) query_term(tile + tile_step);
                mbar_wait(bar_acc + 8 * layer, (uint32_t)(it & 1));
                tc_fence_after();
                {
                    const long long now = tick<PROF>();
                    t_wait += now - t_mark;
                    t_mark = now;
                }
                const float* bias = s_bias + layer * 256;
                const uint32_t dcol = layer == 1 ? 256u : 0u;
                // the TMEM load of chunk cb+1 is in flight while chunk cb is converted (register double buffer)
                uint32_t v[2][32];
                tmem_ld32_issue(tmem + ((uint32_t)(lane_grp * 32) << 16) + dcol + half * 32, v[0]);
#pragma unroll
                for (int cb = 0; cb < kChunks; ++cb) {
                    const int col0 = cb * 64 + half * 32;
                    float4 bv[8];  // the chunk's bias, fetched while the TMEM load is in flight
#pragma unroll
                    for (int c = 0; c < 8; ++c) bv[c] = *reinterpret_cast<const float4*>(bias + col0 + 4 * c);
                    tmem_ld_wait(v[cb & 1]);
                    if (cb + 1 < kChunks)
                        tmem_ld32_issue(tmem + ((uint32_t)(lane_grp * 32) << 16) + dcol + col0 + 64, v[(cb + 1) & 1]);
#pragma unroll
                    for (int k8 = 0; k8 < 4; ++k8) {
                        float x[8], raw[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) raw[c] = __uint_as_float(v[cb & 1][k8 * 8 + c]);
                        bias_relu8(x, raw, bv[2 * k8], bv[2 * k8 + 1]);
                        uint4 hi, lo;
                        split8(x, hi, lo);
                        const int kblk = (col0 >> 3) + k8;
                        *reinterpret_cast<uint4*>(smem + kOffAhi + kblk * kALbo + row * 16) = hi;
                        *reinterpret_cast<uint4*>(smem + kOffAlo + kblk * kALbo + row * 16) = lo;
                    }
                    warp_arrive_cluster(lead_chunk + 8 * cb, lane);
                }
                {
                    const long long now = tick<PROF>();
                    t_epi += now - t_mark;
                    t_mark = now;
                }
            }
        }
        if (prof && blockIdx.x == 0 && gt == 0) {  // cycles of G warp 0: gather, waiting for MMAs, E2 + E3
            prof[4] = t_gather;
            prof[5] = t_wait;
            prof[6] = t_epi;
        }
    } else {
        // ---------------------------------------------------------------- S group: softmax, head mean, attention pooling
        // Two warps per TMEM lane group (warps 12..19 = lane groups 0..3, 0..3): warp (lane group, sq) owns heads [32 sq, 32 sq + 32) of the
        // lane group's 32 rows in the softmax and column blocks [4 sq, 4 sq + 4) in the pooling.  (Round 2 had six S warps: lane groups 0 / 1
        // were served by ONE warp each, the S group took 13.8 k cycles per tile and fc3 of the next tile waited ~3 k cycles for D1.)
        const int sw = warp - kCtlWarps - kGWarps;      // 0..7
        const int st = tid - kCtlWarps * 32 - kGThreads;  // 0..255
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;
        const int sq = sw >> 2;
        const int cb0 = 4 * sq, cb1 = 4 * sq + 4;
        float* s_red = s_pool;                 // [2][4 lane groups][64 heads]: per-warp maxima, then per-warp sums (free until the pooling)
        float* s_att = s_attp;                 // [2 head halves][128]: partial attention weight of every row of the tile
        const uint32_t trow = tmem + ((uint32_t)(lane_grp * 32) << 16);
        long long t_wait = 0, t_soft = 0, t_pool = 0, t_mark = tick<PROF>();
        for (long long it = 0; it < iters; ++it) {
            const long long tile = tile0 + it * tile_step;
            mbar_wait(bar_acc + 16, (uint32_t)(it & 1));
            tc_fence_after();
            {
                const long long now = tick<PROF>();
                t_wait += now - t_mark;
                t_mark = now;
            }
            // softmax over the 64 neighbours of a query = the 64 rows of lane groups {0,1} or {2,3}; thread = row, 32 heads each (the
            // head's bias shifts every score of the head alike and cancels).  Reductions over a warp's 32 rows: recursive halving on a
            // COPY (lane l ends with head 32 sq + l), so the scores / exponentials stay in registers and D0 is handed back right after
            // the one TMEM load
            float e[32];
            {
                tmem_ld32(trow + 32 * sq, e);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(lead_d0free);  // fc2 of the next tile may overwrite D0
                float t[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = e[i];
#pragma unroll
                for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
                    const bool upper = (lane & off) != 0;
#pragma unroll
                    for (int i = 0; i < n; ++i) {
                        const float send = upper ? t[i] : t[i + n];
                        const float keep = upper ? t[i + n] : t[i];
                        t[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
                    }
                }
                s_red[lane_grp * 64 + 32 * sq + lane] = t[0];
            }
            s_barrier();
            {
                const float* m0 = s_red + lane_grp * 64 + 32 * sq;
                const float* m1 = s_red + (lane_grp ^ 1) * 64 + 32 * sq;  // the other 32 rows of the same query
#pragma unroll
                for (int h = 0; h < 32; h += 4) {
                    const float4 a = *reinterpret_cast<const float4*>(m0 + h), b = *reinterpret_cast<const float4*>(m1 + h);
                    // ex2.approx on (score - max) <= 0: relative error 2^-22 plus |x| 2^-24 from the scaling, far inside the contract;
                    // the accurate expf costs five times the issue slots
                    e[h] = __expf(e[h] - fmaxf(a.x, b.x));
                    e[h + 1] = __expf(e[h + 1] - fmaxf(a.y, b.y));
                    e[h + 2] = __expf(e[h + 2] - fmaxf(a.z, b.z));
                    e[h + 3] = __expf(e[h + 3] - fmaxf(a.w, b.w));
                }
                float t[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = e[i];
#pragma unroll
                for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
                    const bool upper = (lane & off) != 0;
#pragma unroll
                    for (int i = 0; i < n; ++i) {
                        const float send = upper ? t[i] : t[i + n];
                        const float keep = upper ? t[i + n] : t[i];
                        t[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                s_red[256 + lane_grp * 64 + 32 * sq + lane] = t[0];
            }
            s_barrier();
            {
                const float* z0 = s_red + 256 + lane_grp * 64 + 32 * sq;
                const float* z1 = s_red + 256 + (lane_grp ^ 1) * 64 + 32 * sq;
                float a = 0.f;
#pragma unroll
                for (int h = 0; h < 32; h += 4) {
                    const float4 u = *reinterpret_cast<const float4*>(z0 + h), v = *reinterpret_cast<const float4*>(z1 + h);
                    a += __fdividef(e[h], u.x + v.x);  // sums are in [1, 64]: rcp.approx + multiply, 1 ulp
                    a += __fdividef(e[h + 1], u.y + v.y);
                    a += __fdividef(e[h + 2], u.z + v.z);
                    a += __fdividef(e[h + 3], u.w + v.w);
                }
                s_att[sq * 128 + row] = a * (1.f / kHeads);  // this warp's 32 heads of the mean over the heads of softmax_k
            }
            s_barrier();
            {
                const long long now = tick<PROF>();
                t_soft += now - t_mark;
                t_mark = now;
            }
            // pooled[q, :] = sum_j att_j * h3[j, :] straight from the fp32 accumulator of fc3 (still in TMEM): every thread
            // scales its row, the 32 rows of a warp are summed by recursive halving (lane l ends with column l of the block)
            {
                const float a = s_att[row] + s_att[128 + row];
                const float* bias = s_bias + 256;
#pragma unroll 1
                for (int cb = cb0; cb < cb1; ++cb) {
                    const int col0 = cb * 32;
                    uint32_t vr[32];
                    tmem_ld32_issue(tmem + ((uint32_t)(lane_grp * 32) << 16) + 256u + col0, vr);
                    float4 bv[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) bv[c] = *reinterpret_cast<const float4*>(bias + col0 + 4 * c);
                    tmem_ld_wait(vr);
                    if (cb == cb1 - 1) {  // last read of fc3's accumulator: fc3 of the next tile may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(lead_d1free);
                    }
                    float v[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        v[4 * c + 0] = a * fmaxf(__uint_as_float(vr[4 * c + 0]) + bv[c].x, 0.f);
                        v[4 * c + 1] = a * fmaxf(__uint_as_float(vr[4 * c + 1]) + bv[c].y, 0.f);
                        v[4 * c + 2] = a * fmaxf(__uint_as_float(vr[4 * c + 2]) + bv[c].z, 0.f);
                        v[4 * c + 3] = a * fmaxf(__uint_as_float(vr[4 * c + 3]) + bv[c].w, 0.f);
                    }
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
                    s_pool[lane_grp * 256 + col0 + lane] = v[0];
                }
            }
            s_barrier();
            for (int e = st; e < 2 * kC; e += kSThreads) {
                const int t = e >> 8, c = e & 255;
                const long long q = 2 * tile + t;
                if (q < nq) pooled[q * kC + c] = s_pool[(2 * t) * 256 + c] + s_pool[(2 * t + 1) * 256 + c];
            }
            s_barrier();  // the head sums and partial pooled sums are consumed before the next tile overwrites them
            {
                const long long now = tick<PROF>();
                t_pool += now - t_mark;
                t_mark = now;
            }
        }
        if (prof && blockIdx.x == 0 && st == 0) {  // cycles of S warp 0: waiting for fc_query, softmax, pooling
            prof[7] = t_wait;
            prof[8] = t_soft;
            prof[9] = t_pool;
        }
    }

    tc_fence_before();
    __syncthreads();
    __syncwarp();
    cluster_sync_all();  // neither CTA exits (or frees its TMEM) while the peer may still arrive on its barriers or read its tile
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace tc

static long long* g_tc_prof = nullptr;
static uint32_t g_tc_terms = tc::kProductTerms;  // split-fp16 terms per layer of projection_tc_kernel (pps_decoder_tc_terms)
void set_tc_terms(uint32_t m) { g_tc_terms = m & 0x1FFu; }
uint32_t get_tc_terms() { return g_tc_terms; }  // device buffer of 128 counters, set by pps_debug_tc_profile

size_t projection_tc_workspace(const pps_decoder_weights*, int64_t) { return 256; }

int projection_tc_impl(const pps_decoder_weights* w, const float* table, const float* queries, const int32_t* idx, int k_stride,
                       int64_t q, void*, size_t, float* pooled, cudaStream_t st) {
    PPS_CHECK_ARG(w->tc_wpack != nullptr, "decoder path 1: the weights carry no tensor-core pack (tc_wpack is null)");
    PPS_CHECK_ARG(w->k == tc::kNbrs && w->latent == tc::kC && w->heads == tc::kHeads,
                  "decoder path 1 is built for k=64, latent=256, heads=64 (got %d, %d, %d)", w->k, w->latent, w->heads);
    if (q == 0) return PPS_OK;
    static unsigned char configured[kMaxDevices] = {};
    if (first_use_on_device(configured)) {
        PPS_CUDA(cudaFuncSetAttribute(tc::projection_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
        PPS_CUDA(cudaFuncSetAttribute(tc::projection_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
        PPS_CUDA(cudaFuncSetAttribute(tc::projection_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
        PPS_CUDA(cudaFuncSetAttribute(tc::projection_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
    }
    const long long npt = (q + 3) / 4;  // pair-tiles of 4 queries
    const int pairs = (int)(npt < kNumSMs / 2 ? npt : kNumSMs / 2);
    const uint8_t* wp = static_cast<const uint8_t*>(w->tc_wpack);
    const long long nq = q;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(tc::kThreads);
    cfg.dynamicSmemBytes = tc::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;  // CTA pairs: the two CTAs of a cluster sit on the two SMs of one TPC
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    profile_begin(st);
    const uint32_t mask = g_tc_terms;
    if (g_tc_prof && mask != tc::kProductTerms)  // instrumented ablation build
        PPS_CUDA(cudaLaunchKernelEx(&cfg, tc::projection_tc_kernel<true, true>, table, queries, idx, k_stride, nq, wp, w->b2, w->b3, w->bq,
                                    w->w1_xyz, pooled, g_tc_prof, mask));
    else if (g_tc_prof)  // instrumented build, tools/tc_phase_profile.py only
        PPS_CUDA(cudaLaunchKernelEx(&cfg, tc::projection_tc_kernel<true, false>, table, queries, idx, k_stride, nq, wp, w->b2, w->b3, w->bq,
                                    w->w1_xyz, pooled, g_tc_prof, mask));
    else if (mask != tc::kProductTerms)  // pass-ablation build (tools/pass_ablation.py)
        PPS_CUDA(cudaLaunchKernelEx(&cfg, tc::projection_tc_kernel<false, true>, table, queries, idx, k_stride, nq, wp, w->b2, w->b3, w->bq,
                                    w->w1_xyz, pooled, g_tc_prof, mask));
    else
        PPS_CUDA(cudaLaunchKernelEx(&cfg, tc::projection_tc_kernel<false, false>, table, queries, idx, k_stride, nq, wp, w->b2, w->b3, w->bq,
                                    w->w1_xyz, pooled, g_tc_prof, mask));
    PPS_LAUNCH_CHECK();
    profile_end(st);
    return PPS_OK;
}

}  // namespace pps

extern "C" size_t pps_decoder_tc_pack_bytes(void) { return pps::tc::kPackBytes; }
// debug: CTA 0 of projection_tc_kernel writes its per-phase cycle counters into `counters` (128 x int64, device memory);
// pass NULL to switch the instrumentation output off
extern "C" void pps_debug_tc_profile(long long* counters) { pps::g_tc_prof = counters; }
// debug: how many CTA pairs of projection_tc_kernel the device can hold at once (-1 on error)
extern "C" int pps_debug_tc_max_clusters(void) {
    cudaFuncSetAttribute(pps::tc::projection_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pps::tc::kSmemBytes);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pps::kNumSMs);
    cfg.blockDim = dim3(pps::tc::kThreads);
    cfg.dynamicSmemBytes = pps::tc::kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    if (cudaOccupancyMaxActiveClusters(&n, pps::tc::projection_tc_kernel<false, false>, &cfg) != cudaSuccess) return -1;
    return n;
}
// retired debug knob (cluster multicast / weight-stream experiments); kept so that the ABI is stable
extern "C" void pps_debug_tc_cluster(int) {}
// split-fp16 terms of the global branch's three GEMM layers: bit 3*layer + t, layer 0 = fc2, 1 = fc3, 2 = fc_query; t = 0: x_hi w_hi
// (always issued), 1: x_lo w_hi, 2: x_hi w_lo.  Default 0x1FF: three terms everywhere (0x0FF: two for fc_query).
// Returns the previous mask.
extern "C" int pps_decoder_tc_terms(int mask) {
    const int old = (int)pps::get_tc_terms();
    if (mask >= 0) pps::set_tc_terms((uint32_t)mask | 0x49u);
    return old;
}
