// umma_bench.cu — micro-benchmark: cycles per tcgen05.mma (M=128, K=16, kind::f16) reading both operands from shared
// memory, for the no-swizzle K-major layout the convolution uses vs the 128B-swizzled layout GEMMs use, over N and the
// number of issuing warps.  Timing only (operands are zeros).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// layout 0: no swizzle, rows 16 B apart (SBO 128, LBO = lbo bytes).  layout 2: 128B swizzle, rows 128 B apart (SBO 1024).
__global__ void __launch_bounds__(160, 1) bench(int layout, int N, int nwarps, int iters, int distinct_a, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)));
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps && lane == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 96 * 1024;
        uint64_t hi, a_lo0, b_lo0;
        if (layout == 0) { hi = (uint64_t)(8u | (1u << 14)) << 32; a_lo0 = (a_base >> 4) | (600u << 16); b_lo0 = (b_base >> 4) | ((uint32_t)N << 16); }
        else { hi = ((uint64_t)(64u | (1u << 14)) << 32) | (2ull << 61); a_lo0 = (a_base >> 4) | (1u << 16); b_lo0 = (b_base >> 4) | (1u << 16); }
        const uint32_t d = tmem + warp * (N > 128 ? 256 : 128);   // own accumulator
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t aoff = distinct_a ? (uint32_t)((i % 9) * 33) : 0u;   // shifted views like the conv taps
            mma(d, hi | (a_lo0 + aoff), hi | b_lo0, idesc, i ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + warp)) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bar + warp)) : "memory");
        t1 = clock64();
        if (blockIdx.x == 0) out[warp] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4000;
    printf("layout N warps distinctA : cycles/MMA per warp | per SM\n");
    for (int layout : {0, 2})
        for (int N : {16, 32, 48, 96, 128, 256})
            for (int nw : {1, 2, 4})
                for (int da : {0, 1}) {
                    if (nw * (N > 128 ? 256 : 128) > 512) continue;
                    if (layout == 2 && da) continue;
                    bench<<<148, 160, 200 * 1024>>>(layout, N, nw, iters, da, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    printf("%d %3d %d %d : %7.1f | %7.1f\n", layout, N, nw, da, (double)out[0] / iters, (double)out[0] / iters / nw);
                }
    return 0;
}
