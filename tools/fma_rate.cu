// fma_rate.cu — issue rate per SM of the arithmetic the sweep kernel is made of: FFMA, packed FFMA2, HFMA2, and the mixed-precision
// FHFMA / FHADD (fp16 operands promoted inside the instruction, fp32 accumulate).   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed) {
    float a[8];
    __half2 h[8];
    float2 p[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + i; h[i] = __floats2half2_rn(seed + i, seed - i); p[i] = make_float2(seed + i, seed - i); }
    const float m = 1.0001f;
    const __half2 hm = __floats2half2_rn(1.001f, 0.999f);
    const unsigned short hs = __half_as_ushort(__float2half(1.001f));
    const float2 pm = make_float2(1.0001f, 0.9999f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = fmaf(a[i], m, 0.5f);
            if (OP == 1) p[i] = __ffma2_rn(p[i], pm, pm);
            if (OP == 2) h[i] = __hfma2(h[i], hm, hm);
            if (OP == 3) asm volatile("fma.rn.f32.f16 %0, %1, %1, %0;" : "+f"(a[i]) : "h"(hs));
            if (OP == 4) asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(a[i]) : "h"(hs));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += a[i] + __low2float(h[i]) + p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out)[1 << 20] = t1 - t0;
}

int main() {
    float* out;
    cudaMalloc(&out, (1 << 23) + 64);
    const int iters = 4096;
    const char* names[5] = {"FFMA", "FFMA2 (2 x fp32)", "HFMA2 (2 x fp16)", "FHFMA (f16 x f16 + f32)", "FHADD (f16 + f32)"};
    printf("op : warp-instructions per clock per SM (512 threads = 4 warps per scheduler, 8 independent chains each)\n");
    for (int op = 0; op < 5; ++op) {
        long long cyc = 0;
        for (int rep = 0; rep < 2; ++rep) {
            if (op == 0) k<0><<<148, 512>>>(out, iters, 1.f);
            if (op == 1) k<1><<<148, 512>>>(out, iters, 1.f);
            if (op == 2) k<2><<<148, 512>>>(out, iters, 1.f);
            if (op == 3) k<3><<<148, 512>>>(out, iters, 1.f);
            if (op == 4) k<4><<<148, 512>>>(out, iters, 1.f);
            cudaDeviceSynchronize();
            cudaMemcpy(&cyc, reinterpret_cast<long long*>(out) + (1 << 20), 8, cudaMemcpyDeviceToHost);
        }
        printf("%-24s : %.2f\n", names[op], 16.0 * 8 * iters / (double)cyc);
    }
    return 0;
}
