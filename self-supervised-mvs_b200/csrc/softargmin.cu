// softargmin.cu — softmax over the depth axis + expected depth + expected index + 4-plane confidence,
// one pass structure per pixel column (thread = pixel, lanes along w so every plane read is coalesced).
// Reference: jdacs/models/mvsnet.py:141-151, module.py:145-148; jdacs-ms/models/modules.py:324-331,
// network.py:147-148, 173-189.  The three reference passes over the [B,D,H,W] volume (softmax, two
// regressions, pad+avg_pool+gather) become two sweeps of the column that stay in L2.
#include "mvs_rt.h"

__global__ void __launch_bounds__(128)
softargmin_fwd_kernel(const float* __restrict__ cost, const float* __restrict__ depth, int per_pixel,
                      float* __restrict__ depth_out, int64_t* __restrict__ index_out, float* __restrict__ conf_out,
                      float* __restrict__ prob_out, int B, int D, int HW) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HW) return;
    const int p = (int)(i % HW);
    const int b = (int)(i / HW);
    const float* col = cost + (int64_t)b * D * HW + p;
    // sweep 1: max and sum of exponentials (as ATen's softmax: exp(x - max) / sum)
    // only B*H*W threads exist (20 k at the headline size), so latency is hidden with memory-level parallelism:
    // every sweep is unrolled to keep 16 independent loads in flight; the accumulation ORDER stays sequential in d.
    float mx = -INFINITY;
#pragma unroll 16
    for (int d = 0; d < D; ++d) mx = fmaxf(mx, __ldg(col + (int64_t)d * HW));
    float sum = 0.f;
#pragma unroll 16
    for (int d = 0; d < D; ++d) sum += expf(__ldg(col + (int64_t)d * HW) - mx);
    // sweep 2: expectations, ascending d, fp32 (hazard H12: the index is a truncated float sum)
    float e_depth = 0.f, e_index = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
        const float pd = expf(__ldg(col + (int64_t)d * HW) - mx) / sum;
        const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
        e_depth += pd * dv;
        e_index += pd * (float)d;
        if (prob_out) prob_out[((int64_t)b * D + d) * HW + p] = pd;
    }
    if (depth_out) depth_out[i] = e_depth;
    if (index_out || conf_out) {
        const long long idx = (long long)e_index;  // trunc toward zero, like .long()
        if (index_out) index_out[i] = (int64_t)idx;
        if (conf_out) {
            // 4 * avg_pool3d over planes idx-1 .. idx+2 of the zero-padded volume = windowed sum, then /4*4
            float s = 0.f;
            for (int k = -1; k <= 2; ++k) {
                const long long d = idx + k;
                if (d >= 0 && d < D) s += expf(__ldg(col + (int64_t)d * HW) - mx) / sum;
            }
            conf_out[i] = 4.f * (s / 4.f);
        }
    }
}

#ifndef MVS_CPU_EMU
// Same arithmetic, four times the memory-level parallelism: a block = 32 pixel columns x 4 quarters of the depth axis.  The
// quarters read the column once from global memory into shared memory (coalesced along w) and exponentiate it there in
// parallel (both are order-free); one warp then walks the column sequentially in shared memory for the two order-sensitive
// sums (sum of exponentials, expectations: ascending plane order is part of the parity contract, hazard H12).  The column is
// read from DRAM once instead of three times and expf runs once per element instead of twice, with identical results.
__global__ void __launch_bounds__(128)
softargmin_fwd_smem_kernel(const float* __restrict__ cost, const float* __restrict__ depth, int per_pixel,
                           float* __restrict__ depth_out, int64_t* __restrict__ index_out, float* __restrict__ conf_out,
                           float* __restrict__ prob_out, int B, int D, int HW) {
    extern __shared__ float xs[];                 // [D][32]
    __shared__ float red[4][32];
    const int c = threadIdx.x & 31, qd = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + c;
    const bool live = i < (int64_t)B * HW;
    const int p = live ? (int)(i % HW) : 0;
    const int b = live ? (int)(i / HW) : 0;
    const float* col = cost + (int64_t)b * D * HW + p;
    const int dq = (D + 3) / 4, d0 = qd * dq, d1 = min(D, d0 + dq);
    float mx = -INFINITY;
#pragma unroll 16
    for (int d = d0; d < d1; ++d) {
        const float x = live ? __ldg(col + (int64_t)d * HW) : 0.f;
        xs[d * 32 + c] = x;
        mx = fmaxf(mx, x);
    }
    red[qd][c] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0][c], red[1][c]), fmaxf(red[2][c], red[3][c]));
#pragma unroll 8
    for (int d = d0; d < d1; ++d) xs[d * 32 + c] = expf(xs[d * 32 + c] - mx);
    __syncthreads();
    // sum of exponentials: sequential in d (one warp), then the normalisation in parallel (an IEEE divide per element is the
    // expensive part and is order-free), then the two expectations sequentially again
    if (qd == 0) {
        float sum = 0.f;
#pragma unroll 8
        for (int d = 0; d < D; ++d) sum += xs[d * 32 + c];
        red[0][c] = sum;
    }
    __syncthreads();
    const float sum = red[0][c];
#pragma unroll 8
    for (int d = d0; d < d1; ++d) {
        const float pd = xs[d * 32 + c] / sum;
        xs[d * 32 + c] = pd;
        if (prob_out && live) prob_out[((int64_t)b * D + d) * HW + p] = pd;
    }
    __syncthreads();
    if (qd != 0 || !live) return;
    float e_depth = 0.f, e_index = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
        const float pd = xs[d * 32 + c];
        const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
        e_depth += pd * dv;
        e_index += pd * (float)d;
    }
    if (depth_out) depth_out[i] = e_depth;
    if (index_out || conf_out) {
        const long long idx = (long long)e_index;  // trunc toward zero, like .long()
        if (index_out) index_out[i] = (int64_t)idx;
        if (conf_out) {
            float s = 0.f;
            for (int k = -1; k <= 2; ++k) {
                const long long d = idx + k;
                if (d >= 0 && d < D) s += xs[(int)d * 32 + c];
            }
            conf_out[i] = 4.f * (s / 4.f);
        }
    }
}
#endif

// d depth / d cost_d = p_d (depth_d - E[depth])
__global__ void __launch_bounds__(128)
softargmin_bwd_kernel(const float* __restrict__ cost, const float* __restrict__ depth, int per_pixel,
                      const float* __restrict__ grad_depth, float* __restrict__ grad_cost, int B, int D, int HW) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HW) return;
    const int p = (int)(i % HW);
    const int b = (int)(i / HW);
    const float* col = cost + (int64_t)b * D * HW + p;
    float mx = -INFINITY;
    for (int d = 0; d < D; ++d) mx = fmaxf(mx, __ldg(col + (int64_t)d * HW));
    float sum = 0.f;
    for (int d = 0; d < D; ++d) sum += expf(__ldg(col + (int64_t)d * HW) - mx);
    float e_depth = 0.f;
    for (int d = 0; d < D; ++d) {
        const float pd = expf(__ldg(col + (int64_t)d * HW) - mx) / sum;
        const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
        e_depth += pd * dv;
    }
    const float g = __ldg(grad_depth + i);
    for (int d = 0; d < D; ++d) {
        const float pd = expf(__ldg(col + (int64_t)d * HW) - mx) / sum;
        const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
        grad_cost[((int64_t)b * D + d) * HW + p] = g * pd * (dv - e_depth);
    }
}

extern "C" int mvs_softargmin_fwd(const float* cost, const float* depth, int per_pixel, float* depth_out,
                                  int64_t* index_out, float* conf_out, float* prob_out, int B, int D, int H, int W,
                                  void* stream) {
    MVS_REQUIRE(cost && depth, MVS_E_ARG, "mvs_softargmin_fwd: null pointer");
    MVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, MVS_E_SHAPE, "mvs_softargmin_fwd: bad dims");
    const int64_t total = (int64_t)B * H * W;
#ifndef MVS_CPU_EMU
    const size_t smem = (size_t)D * 32 * sizeof(float);
    if (smem <= 96 * 1024) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(softargmin_fwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        softargmin_fwd_smem_kernel<<<mvs_cdiv(total, 32), 128, smem, (cudaStream_t)stream>>>(cost, depth, per_pixel, depth_out, index_out,
                                                                                           conf_out, prob_out, B, D, H * W);
        return MVS_CHECK_LAUNCH("mvs_softargmin_fwd");
    }
#endif
    const int bs = total < 148 * 1024 ? 32 : 128;   // few pixel columns: one warp per block spreads them over all SMs
    MVS_LAUNCH(softargmin_fwd_kernel, dim3(mvs_cdiv(total, bs)), dim3(bs), stream, cost, depth, per_pixel, depth_out,
               index_out, conf_out, prob_out, B, D, H * W);
    return MVS_CHECK_LAUNCH("mvs_softargmin_fwd");
}

extern "C" int mvs_softargmin_bwd(const float* cost, const float* depth, int per_pixel, const float* grad_depth,
                                  float* grad_cost, int B, int D, int H, int W, void* stream) {
    MVS_REQUIRE(cost && depth && grad_depth && grad_cost, MVS_E_ARG, "mvs_softargmin_bwd: null pointer");
    MVS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, MVS_E_SHAPE, "mvs_softargmin_bwd: bad dims");
    const int64_t total = (int64_t)B * H * W;
    MVS_LAUNCH(softargmin_bwd_kernel, dim3(mvs_cdiv(total, 128)), dim3(128), stream, cost, depth, per_pixel, grad_depth,
               grad_cost, B, D, H * W);
    return MVS_CHECK_LAUNCH("mvs_softargmin_bwd");
}
