// invwarp.cu — photometric inverse warp of the self-supervised loss (a10) and CVP depth hypotheses (a9).
//
// inverse_warping (jdacs/losses/homography.py:186-238, sampler :292-374): back-project every reference pixel
// with the estimated depth, project into the source view, bilinear-sample the source image with clamped taps
// and emit a validity mask.  The ~40 small ATen launches per view of the reference become one launch; the
// backward sends the gradient to DEPTH through the sampling coordinates (hazard H13) and optionally to the image.
// Quirks kept on purpose: the reference intrinsics are used for both cameras (H6); the mask tests
// y0 <= H-1 while the weights come from the clamped x1 / y1 (H7).
#include "mvs_rt.h"
#include "linalg.h"
#include "invwarp_dev.h"

__global__ void invwarp_cam_kernel(const float* __restrict__ left, const float* __restrict__ right, float* __restrict__ ws, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    invwarp_compose_cam(left + (int64_t)b * 32, right + (int64_t)b * 32, ws + (int64_t)b * 24);
}

__global__ void __launch_bounds__(128)
invwarp_fwd_kernel(const float* __restrict__ img, const float* __restrict__ depth, const float* __restrict__ ws,
                   float* __restrict__ warped, float* __restrict__ mask, int B, int H, int W, int C) {
    const int HW = H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HW) return;
    const int p = (int)(i % HW), b = (int)(i / HW);
    InvTaps t;
    invwarp_taps(ws + (int64_t)b * 24, __ldg(depth + i), p % W, p / W, H, W, t);
    const float wa = t.fx * t.fy, wb = t.fx * (1.f - t.fy), wc = (1.f - t.fx) * t.fy, wd = (1.f - t.fx) * (1.f - t.fy);
    const float* im = img + (int64_t)b * HW * C;
    for (int c = 0; c < C; ++c)
        warped[i * C + c] = wa * __ldg(im + (int64_t)t.ia * C + c) + wb * __ldg(im + (int64_t)t.ib * C + c) +
                            wc * __ldg(im + (int64_t)t.ic * C + c) + wd * __ldg(im + (int64_t)t.id * C + c);
    mask[i] = t.mask;
}

__global__ void __launch_bounds__(128)
invwarp_bwd_kernel(const float* __restrict__ img, const float* __restrict__ depth, const float* __restrict__ ws,
                   const float* __restrict__ gw, float* __restrict__ gdepth, float* __restrict__ gimg, int B, int H, int W, int C) {
    const int HW = H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HW) return;
    const int p = (int)(i % HW), b = (int)(i / HW);
    InvTaps t;
    invwarp_taps(ws + (int64_t)b * 24, __ldg(depth + i), p % W, p / W, H, W, t);
    const float wa = t.fx * t.fy, wb = t.fx * (1.f - t.fy), wc = (1.f - t.fx) * t.fy, wd = (1.f - t.fx) * (1.f - t.fy);
    const float* im = img + (int64_t)b * HW * C;
    float gfx = 0.f, gfy = 0.f;  // d loss / d fx, d loss / d fy
    for (int c = 0; c < C; ++c) {
        const float g = __ldg(gw + i * C + c);
        const float a = __ldg(im + (int64_t)t.ia * C + c), bb = __ldg(im + (int64_t)t.ib * C + c);
        const float cc = __ldg(im + (int64_t)t.ic * C + c), dd = __ldg(im + (int64_t)t.id * C + c);
        gfx += g * (t.fy * a + (1.f - t.fy) * bb - t.fy * cc - (1.f - t.fy) * dd);
        gfy += g * (t.fx * a - t.fx * bb + (1.f - t.fx) * cc - (1.f - t.fx) * dd);
        if (gimg) {
            float* gi = gimg + (int64_t)b * HW * C;
            atomicAdd(gi + (int64_t)t.ia * C + c, g * wa);
            atomicAdd(gi + (int64_t)t.ib * C + c, g * wb);
            atomicAdd(gi + (int64_t)t.ic * C + c, g * wc);
            atomicAdd(gi + (int64_t)t.id * C + c, g * wd);
        }
    }
    // fx = x1 - x, fy = y1 - y  ->  d/dx = -d/dfx ; the normalise / un-normalise pair has unit slope
    gdepth[i] = -(gfx * t.dx_dd + gfy * t.dy_dd);
}

static int check_invwarp(const char* who, const void* a, const void* b, const void* c, const void* d, const void* e, int B, int H, int W, int C) {
    MVS_REQUIRE(a && b && c && d && e, MVS_E_ARG, "%s: null pointer", who);
    MVS_REQUIRE(B > 0 && H > 1 && W > 1 && C > 0, MVS_E_SHAPE, "%s: bad dims", who);
    return MVS_OK;
}

extern "C" int mvs_invwarp_fwd(const float* img, const float* left_cam, const float* right_cam, const float* depth,
                               float* warped, float* mask, float* cam_ws, int B, int H, int W, int C, void* stream) {
    int rc = check_invwarp("mvs_invwarp_fwd", img, left_cam, right_cam, depth, cam_ws, B, H, W, C);
    if (rc) return rc;
    MVS_REQUIRE(warped && mask, MVS_E_ARG, "mvs_invwarp_fwd: null output");
    MVS_LAUNCH(invwarp_cam_kernel, dim3(mvs_cdiv(B, 32)), dim3(32), stream, left_cam, right_cam, cam_ws, B);
    MVS_LAUNCH(invwarp_fwd_kernel, dim3(mvs_cdiv((int64_t)B * H * W, 128)), dim3(128), stream, img, depth, cam_ws, warped, mask, B, H, W, C);
    return MVS_CHECK_LAUNCH("mvs_invwarp_fwd");
}

extern "C" int mvs_invwarp_bwd(const float* img, const float* left_cam, const float* right_cam, const float* depth,
                               const float* grad_warped, float* grad_depth, float* grad_img, float* cam_ws, int B, int H,
                               int W, int C, void* stream) {
    int rc = check_invwarp("mvs_invwarp_bwd", img, left_cam, right_cam, depth, cam_ws, B, H, W, C);
    if (rc) return rc;
    MVS_REQUIRE(grad_warped && grad_depth, MVS_E_ARG, "mvs_invwarp_bwd: null pointer");
    MVS_LAUNCH(invwarp_cam_kernel, dim3(mvs_cdiv(B, 32)), dim3(32), stream, left_cam, right_cam, cam_ws, B);
    MVS_LAUNCH(invwarp_bwd_kernel, dim3(mvs_cdiv((int64_t)B * H * W, 128)), dim3(128), stream, img, depth, cam_ws, grad_warped, grad_depth, grad_img, B, H, W, C);
    return MVS_CHECK_LAUNCH("mvs_invwarp_bwd");
}

// ------------------------------------------------------------------------------------------------ a9: CVP hypotheses
// calDepthHypo (jdacs-ms/models/modules.py:107-206): per pixel, the depth step that moves the projection into
// source view 0 by one pixel along the epipolar line (2x2 solve in fp64); interval_b = mean |step|; the hypotheses
// are depth_up + k * interval_b, k = -half .. half-1.
struct HypoCam { double Kr_inv[9], Er_inv[16], Ks[9], Es[16], A[9]; };

__device__ static void hypo_cam(const float* ref_in, const float* src_in, const float* ref_ex, const float* src_ex, HypoCam& c) {
    double Kr[9], Er[16], M[9], Minv[9];
    for (int i = 0; i < 9; ++i) { Kr[i] = ref_in[i]; c.Ks[i] = src_in[i]; }
    for (int i = 0; i < 16; ++i) { Er[i] = ref_ex[i]; c.Es[i] = src_ex[i]; }
    inv3(Kr, c.Kr_inv);
    inv4(Er, c.Er_inv);
    // A = Kr Rr (Ks Rs)^-1
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += c.Ks[r * 3 + k] * c.Es[k * 4 + cc];
        M[r * 3 + cc] = s;
    }
    inv3(M, Minv);
    double KR[9];
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += Kr[r * 3 + k] * Er[k * 4 + cc];
        KR[r * 3 + cc] = s;
    }
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += KR[r * 3 + k] * Minv[k * 3 + cc];
        c.A[r * 3 + cc] = s;
    }
}

__device__ static void hypo_project(const HypoCam& c, double x, double y, double d, double (&p)[3], double& z) {
    const double px = x * d, py = y * d, pz = d;
    double cam[3], wld[3], sc[3];
    for (int r = 0; r < 3; ++r) cam[r] = c.Kr_inv[r * 3] * px + c.Kr_inv[r * 3 + 1] * py + c.Kr_inv[r * 3 + 2] * pz;
    for (int r = 0; r < 3; ++r) wld[r] = c.Er_inv[r * 4] * cam[0] + c.Er_inv[r * 4 + 1] * cam[1] + c.Er_inv[r * 4 + 2] * cam[2] + c.Er_inv[r * 4 + 3];
    for (int r = 0; r < 3; ++r) sc[r] = c.Es[r * 4] * wld[0] + c.Es[r * 4 + 1] * wld[1] + c.Es[r * 4 + 2] * wld[2] + c.Es[r * 4 + 3];
    for (int r = 0; r < 3; ++r) p[r] = c.Ks[r * 3] * sc[0] + c.Ks[r * 3 + 1] * sc[1] + c.Ks[r * 3 + 2] * sc[2];
    z = p[2];
    p[0] /= z; p[1] /= z; p[2] = 1.0;
}

__global__ void __launch_bounds__(128)
hypo_interval_kernel(const float* __restrict__ depth_up, const float* __restrict__ ref_in, const float* __restrict__ src_in0,
                     const float* __restrict__ ref_ex, const float* __restrict__ src_ex0, double* __restrict__ ws, int H, int W) {
    const int b = blockIdx.y;
    const int HW = H * W;
    HypoCam c;
    hypo_cam(ref_in + (int64_t)b * 9, src_in0 + (int64_t)b * 9, ref_ex + (int64_t)b * 16, src_ex0 + (int64_t)b * 16, c);
    double acc = 0.0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const double x = (double)(p % W), y = (double)(p / W);
        const double d1 = (double)depth_up[(int64_t)b * HW + p];
        double x1[3], x2[3], z1, z2;
        hypo_project(c, x, y, d1, x1, z1);
        hypo_project(c, x, y, d1 + 1.0, x2, z2);
        const double theta = atan((x2[1] - x1[1]) / (x2[0] - x1[0]));
        const double x3[3] = {x1[0] + cos(theta), x1[1] + sin(theta), x1[2]};
        double t1[3], t2[3];
        for (int r = 0; r < 3; ++r) {
            t1[r] = z1 * (c.A[r * 3] * x1[0] + c.A[r * 3 + 1] * x1[1] + c.A[r * 3 + 2] * x1[2]);
            t2[r] = c.A[r * 3] * x3[0] + c.A[r * 3 + 1] * x3[1] + c.A[r * 3 + 2] * x3[2];
        }
        // [[y, t2y], [1, t2z]] (a, b)^T = (t1y, t1z)
        const double det = y * t2[2] - t2[1];
        acc += fabs((t1[1] * t2[2] - t2[1] * t1[2]) / det);
    }
#ifndef MVS_CPU_EMU
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) != 0) return;
#endif
    atomicAdd(ws + b, acc);
}

__global__ void __launch_bounds__(256)
hypo_fill_kernel(const float* __restrict__ depth_up, const double* __restrict__ ws, float* __restrict__ hypos, int B, int HW, int half) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HW) return;
    const int b = (int)(i / HW), p = (int)(i % HW);
    const double interval = ws[b] / (double)HW;
    const double d = (double)depth_up[i];
    for (int k = -half; k < half; ++k) hypos[((int64_t)b * 2 * half + (k + half)) * HW + p] = (float)(d + (double)k * interval);
}

extern "C" int mvs_depth_hypo_refine(const float* depth_up, const float* ref_in, const float* src_in0, const float* ref_ex,
                                     const float* src_ex0, float* hypos, double* ws, int B, int H, int W, int half, void* stream) {
    MVS_REQUIRE(depth_up && ref_in && src_in0 && ref_ex && src_ex0 && hypos && ws, MVS_E_ARG, "mvs_depth_hypo_refine: null pointer");
    MVS_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && half > 0, MVS_E_SHAPE, "mvs_depth_hypo_refine: bad dims");
    unsigned bx = mvs_cdiv((int64_t)H * W, 128 * 4);
    if (bx > 296) bx = 296;
    MVS_LAUNCH(hypo_interval_kernel, dim3(bx, (unsigned)B), dim3(128), stream, depth_up, ref_in, src_in0, ref_ex, src_ex0, ws, H, W);
    MVS_LAUNCH(hypo_fill_kernel, dim3(mvs_cdiv((int64_t)B * H * W, 256)), dim3(256), stream, depth_up, ws, hypos, B, H * W, half);
    return MVS_CHECK_LAUNCH("mvs_depth_hypo_refine");
}
