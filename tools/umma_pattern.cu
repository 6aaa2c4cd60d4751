// umma_pattern.cu — micro-benchmark of the ISSUER LOOP of conv3d_tc.cu's kd-folded program: per step, each of 4 issuer warps
// (warp-uniform code, one elected lane issues) sends a few tcgen05.mma of mixed N into its own accumulator, commits to an
// mbarrier and waits on barriers that completed long ago.  Which part costs what?   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_elect(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

// mode bits: 1 = commit per step, 2 = one completed-barrier wait per step (all lanes), 4 = second wait, 8 = second commit,
//            16 = single N=96 MMA per (entry) instead of the split 64 + 32 pattern, 32 = waits by lane 0 + __syncwarp
__global__ void __launch_bounds__(160, 1) bench(int mode, int nent, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + i)), "r"(i == 6 ? 4 : 1));
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (warp < 4) {
        const uint32_t idesc0 = (1u << 4) | ((128u >> 4) << 24);
        const uint32_t id32 = idesc0 | ((32u >> 3) << 17), id64 = idesc0 | ((64u >> 3) << 17), id96 = idesc0 | ((96u >> 3) << 17);
        const uint32_t a_base = smem_u32(smem) + warp * 2048, b_base = smem_u32(smem) + 96 * 1024;
        const uint64_t hi = (uint64_t)(8u | (1u << 14)) << 32;
        const uint64_t a0 = hi | ((a_base >> 4) | (600u << 16)), b0 = hi | ((b_base >> 4) | (96u << 16));
        const uint32_t d = tmem + warp * 128;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (!(mode & 64)) {
                if (mode & 2) { if (mode & 32) { if (lane == 0) mbar_wait(bar + 4, 1); __syncwarp(); } else mbar_wait(bar + 4, 1); }     // fresh barrier: parity-1 wait passes at once
                if (mode & 4) { if (mode & 32) { if (lane == 0) mbar_wait(bar + 5, 1); __syncwarp(); } else mbar_wait(bar + 5, 1); }
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int e = 0; e < nent; ++e) {
                const uint64_t a = a0 + (uint64_t)(e * 32), b = b0 + (uint64_t)(e * 192);
                if (mode & 16) mma_elect(d, a, b, id96, 1u);
                else if (e == 0) { mma_elect(d, a, b, id64, 1u); mma_elect(d + 64, a, b + 128, id32, 0u); }
                else mma_elect(d, a, b, id96, 1u);
            }
            if (mode & 64) {   // 64: the NEXT step's waits are issued behind this step's MMAs, before its commits
                if (mode & 2) mbar_wait(bar + 4, 1);
                if (mode & 4) mbar_wait(bar + 5, 1);
            }
            if (mode & 1) commit_elect(bar + warp);      // nobody waits on these: phases just flip
            if (mode & 8) commit_elect(bar + warp);
        }
        commit_elect(bar + 6);
        long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) out[warp] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) { mbar_wait(bar + 6, 0); }   // drain: all four issuers' MMAs have completed
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    printf("mode nent : cycles per step (issuer warp 0)   [mode bits: 1 commit, 2 wait, 4 wait2, 8 commit2, 16 unsplit N=96, 32 lane-0 waits, 64 waits after the MMAs]\n");
    for (int nent : {2, 6})
        for (int mode : {0, 16, 1, 3, 15, 31, 79, 95, 9}) {
            bench<<<148, 160, 200 * 1024>>>(mode, nent, iters, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("%2d %d : %8.1f\n", mode, nent, (double)out[0] / iters);
        }
    return 0;
}
