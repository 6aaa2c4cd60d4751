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
