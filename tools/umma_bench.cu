// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS, no-swizzle K-major operands) for several N,
// back to back on fixed shared-memory operands.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <cstdio>
#include "../ppsurf_b200/csrc/tc_common.cuh"
using namespace pps::tc;

__global__ void __launch_bounds__(128, 1) bench(int n, int iters, int distinct_a, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long s_bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
    if (tid == 0) {
        mbar_init(smem_u32(&s_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc(n);
        const int a_lbo = 2064, b_lbo = n * 16;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            // rotate through 16 k-steps of an operand tile like the real kernel does
            const int s = i & 15;
            const uint64_t a = umma_desc(sbase + (distinct_a ? 2 * s * a_lbo : 0), a_lbo, 128);
            const uint64_t b = umma_desc(sbase + 132096 + (s % 5) * 16384, b_lbo, 128);
            umma(tmem, a, b, idesc, i > 0 ? 1u : 0u);
        }
        tc_commit(smem_u32(&s_bar));
        mbar_wait(smem_u32(&s_bar), 0);
        long long t1 = clock64();
        out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int iters : {1, 3, 12, 48, 4096}) {
        for (int n : {128, 256}) {
            for (int da : {1}) {
                const int grid = 148;
                bench<<<grid, 128, 220 * 1024>>>(n, iters, da, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long h = 0;
                cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("iters %4d grid %3d  N=%3d  distinct A tiles %d: total %lld cycles, %.1f cycles per MMA (%s)  -> %.0f dense fp16 TFLOP/s per 148 SMs at 1.9 GHz\n", iters, grid, n, da, h,
                       double(h) / iters, cudaGetErrorString(e), 2.0 * 128 * n * 16 / (double(h) / iters) * 148 * 1.9e9 / 1e12);
            }
        }
    }
    return 0;
}
