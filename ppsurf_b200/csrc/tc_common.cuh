// Device helpers shared by the tcgen05 kernels (decode_tc.cu, pointnet_tc.cu): mbarrier, TMA bulk copy, UMMA descriptors,
// tcgen05.mma / tcgen05.ld wrappers and the fp32 -> (fp16 hi, fp16 lo) split.  sm_100a only.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace pps {
namespace tc {

constexpr long long kSpinLimit = 4000000000ll;  // ~2 s of SM clocks: turns a protocol bug into a trap instead of a hang

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait suspends the warp in hardware until the phase completes or the time hint (ns) runs out, so a waiting warp issues a
// handful of instructions instead of spinning; the clock is only consulted every 64 wake-ups
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (done) break;
        if ((spins & 63u) == 63u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > kSpinLimit) __trap();
        }
    }
}
// One lane of a CONVERGED warp.  The single-thread instructions of the async units (tcgen05.mma / commit, cp.async.bulk) read their
// operands from uniform registers; inside an `if (lane == 0)` region the control flow is divergent for the compiler and it wraps
// every one of them in an ELECT / BRA.U.ANY loop with a scoreboard wait (~80 cycles per tcgen05.mma, measured: the MMA issuer of
// the projection kernel needed 510 cycles per k16 step of three MMAs that execute in 450).  With warp-uniform control flow
// around an `if (elect_one())` the compiler keeps descriptors and barrier addresses in uniform registers and predicates the
// instruction itself.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// multicast variant: this CTA's slice lands at the same shared-memory offset of every CTA in `mask` and completes tx bytes on
// the mbarrier at the same offset in each of them
__device__ __forceinline__ void bulk_copy_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA pair (cta_group::2) helpers ---------------------------------------------------------------------------------
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// arrival on a barrier that may live in the peer CTA.  Default semantics (release at CTA scope) on purpose: a cluster-scope
// release / acquire compiles to MEMBAR.ALL.GPU + CCTL.IVALL (L1 invalidation) and made the kernel 70 % slower.  What the pair
// needs is weaker: this CTA's shared-memory writes have been performed and made visible to the async proxy
// (fence.proxy.async.shared::cta) before the arrival leaves the SM; the consumer is the tensor core reading this SM's memory.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a LOCAL barrier whose arrivals may come from the peer CTA (CTA-scope acquire, see mbar_arrive_cluster)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (done) break;
        if ((spins & 63u) == 63u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > kSpinLimit) __trap();
        }
    }
}
// every lane has made its shared-memory writes visible to the async proxy (of either CTA of the pair), lane 0 arrives on the
// barrier at `cluster_addr`
__device__ __forceinline__ void warp_arrive_cluster(uint32_t cluster_addr, int lane) {
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(cluster_addr);
}
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all prior MMAs of the pair -> the barrier at the same offset in both CTAs
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
// instruction descriptor kind::f16 for the CTA pair: D fp32, A/B fp16, both K-major, M=256 (128 rows per CTA)
__host__ __device__ constexpr uint32_t umma_idesc2(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((256u >> 4) << 24); }

// UMMA shared-memory descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B (128 contiguous bytes);
// LBO = byte distance between the two k8 halves of a k16 step, SBO = byte distance between 8-row groups; version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);
}
// instruction descriptor kind::f16: D fp32, A/B fp16, both K-major, M=128
__host__ __device__ constexpr uint32_t umma_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Split form of tmem_ld32: issue the load, do independent work (shared-memory loads of the bias ...), then wait.  The wait
// names the 32 registers as in/out operands so that no consumer can be scheduled above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
// registers -> TMEM, 32 columns of this thread's lane
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

// one arrival per warp: every lane has made its shared-memory writes visible to the async proxy, the warp converges, lane 0
// arrives (256 single-thread arrivals on one mbarrier serialise)
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// ---- packed fp32 pairs (Blackwell FADD2 / FFMA2): one issue slot for two lanes of work; the epilogues are issue-bound
__device__ __forceinline__ void add2(float& r0, float& r1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 a, b, c;\n\t"
        "mov.b64 a, {%2, %3};\n\t"
        "mov.b64 b, {%4, %5};\n\t"
        "add.rn.f32x2 c, a, b;\n\t"
        "mov.b64 {%0, %1}, c;\n\t}"
        : "=f"(r0), "=f"(r1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void sub2(float& r0, float& r1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 a, b, c;\n\t"
        "mov.b64 a, {%2, %3};\n\t"
        "mov.b64 b, {%4, %5};\n\t"
        "sub.rn.f32x2 c, a, b;\n\t"
        "mov.b64 {%0, %1}, c;\n\t}"
        : "=f"(r0), "=f"(r1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// two fp32 -> packed fp16x2 (first argument in the low half), round to nearest, SATURATING to +-65504: the conversion itself
// clamps to the fp16 range, no separate min / max
__device__ __forceinline__ uint32_t cvt_pack_sat(float lo_half, float hi_half) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
    return r;
}

// fp32 -> fp16 hi + fp16 lo, 8 values -> two 16-byte vectors.  hi is the value TRUNCATED to fp16's 10 explicit mantissa
// bits (a mask, so hi converts exactly and x - hi is exact), lo = fp16(x - hi): hi + lo carries 21 mantissa bits.  Per pair of
// values: 2 LOP, 1 FADD2, 2 packed conversions  --  4 issue slots per element with the bias add and the ReLU.
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
    float la, lb;
    sub2(la, lb, a, b, ah, bh);
    hi = cvt_pack_sat(ah, bh);
    lo = cvt_pack_sat(la, lb);
}
// inputs of either sign; out-of-range values saturate to +-65504 in the conversion
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pair(x[2 * i], x[2 * i + 1], h[i], l[i]);
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// x[i] = relu(v[i] + b[i]) for 8 values with packed adds
__device__ __forceinline__ void bias_relu8(float (&x)[8], const float (&v)[8], const float4& b0, const float4& b1) {
    add2(x[0], x[1], v[0], v[1], b0.x, b0.y);
    add2(x[2], x[3], v[2], v[3], b0.z, b0.w);
    add2(x[4], x[5], v[4], v[5], b1.x, b1.y);
    add2(x[6], x[7], v[6], v[7], b1.z, b1.w);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
}

// v[c] = relu(v[c] + bias[c]) for 32 accumulator columns, bias 16-byte aligned in shared memory
__device__ __forceinline__ void bias_relu32(float (&v)[32], const float* bias) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 b = *reinterpret_cast<const float4*>(bias + 4 * c);
        add2(v[4 * c], v[4 * c + 1], v[4 * c], v[4 * c + 1], b.x, b.y);
        add2(v[4 * c + 2], v[4 * c + 3], v[4 * c + 2], v[4 * c + 3], b.z, b.w);
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
}

}  // namespace tc
}  // namespace pps
