// invwarp_dev.h — device code shared by the stand-alone photometric warp (invwarp.cu) and the fused self-supervised loss
// (loss.cu): camera composition, the reference's pixel grid and the clamped 4-tap sampler of inverse_warping
// (jdacs/losses/homography.py:186-238, 241-257, 292-374).
#pragma once
#include "mvs_rt.h"
#include "linalg.h"

// cam_ws[b] = { Kinv[9], P[12] (= K_hom @ [R_rel | t_rel], top 3 rows), pad[3] }
// L, R = one item's [2][4][4] camera blocks ([0] = extrinsic, [1][:3][:3] = intrinsic); o = 24 floats
__device__ __forceinline__ void invwarp_compose_cam(const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ o) {
    double K[9], Kinv[9], Rl[9], Rr[9], tl[3], tr[3];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) { K[r * 3 + c] = L[16 + r * 4 + c]; Rl[r * 3 + c] = L[r * 4 + c]; Rr[r * 3 + c] = R[r * 4 + c]; }
        tl[r] = L[r * 4 + 3]; tr[r] = R[r * 4 + 3];
    }
    inv3(K, Kinv);
    double Rrel[9], trel[3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {  // R_right @ R_left^T
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += Rr[r * 3 + k] * Rl[c * 3 + k];
            Rrel[r * 3 + c] = s;
        }
    for (int r = 0; r < 3; ++r) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += Rrel[r * 3 + k] * tl[k];
        trel[r] = tr[r] - s;
    }
    for (int i = 0; i < 9; ++i) o[i] = (float)Kinv[i];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += K[r * 3 + k] * Rrel[k * 3 + c];
            o[9 + r * 4 + c] = (float)s;
        }
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += K[r * 3 + k] * trel[k];
        o[9 + r * 4 + 3] = (float)s;
    }
    o[21] = o[22] = o[23] = 0.f;
}

struct InvTaps {
    int ia, ib, ic, id;   // pixel indices (y0,x0) (y1,x0) (y0,x1) (y1,x1), clamped
    float fx, fy;         // x1 - x, y1 - y with the CLAMPED x1, y1
    float mask;
    float dx_dd, dy_dd;   // d x / d depth, d y / d depth
};

// reference pixel grid: _meshgrid_abs builds it from linspace(-1, 1, n) (homography.py:241-257, hazard H8)
__device__ __forceinline__ float meshgrid_abs(int i, int n) {
    const float step = 2.0f / (float)(n - 1);
    // torch.linspace fills symmetrically: start + i*step for the first half, end - (n-1-i)*step for the second
    const float lin = (i < n / 2) ? (-1.0f + step * (float)i) : (1.0f - step * (float)(n - 1 - i));
    return (lin + 1.0f) * 0.5f * (float)(n - 1);
}

__device__ __forceinline__ void invwarp_taps(const float* __restrict__ cw, float depth, int px, int py, int H, int W, InvTaps& t) {
    const float gx = meshgrid_abs(px, W), gy = meshgrid_abs(py, H);
    // cam = (Kinv @ (gx, gy, 1)) * depth
    const float r0 = cw[0] * gx + cw[1] * gy + cw[2];
    const float r1 = cw[3] * gx + cw[4] * gy + cw[5];
    const float r2 = cw[6] * gx + cw[7] * gy + cw[8];
    const float X = r0 * depth, Y = r1 * depth, Z = r2 * depth;
    const float* P = cw + 9;
    const float qx = P[0] * X + P[1] * Y + P[2] * Z + P[3];
    const float qy = P[4] * X + P[5] * Y + P[6] * Z + P[7];
    const float qz = P[8] * X + P[9] * Y + P[10] * Z + P[11];
    const float zz = qz + 1e-10f;
    const float u = qx / zz, v = qy / zz;
    // derivative of (u, v) w.r.t. depth: q = a * depth + t with a = P[:, :3] @ r
    const float ax = P[0] * r0 + P[1] * r1 + P[2] * r2;
    const float ay = P[4] * r0 + P[5] * r1 + P[6] * r2;
    const float az = P[8] * r0 + P[9] * r1 + P[10] * r2;
    t.dx_dd = (ax - u * az) / zz;
    t.dy_dd = (ay - v * az) / zz;
    // normalise (_spatial_transformer) then un-normalise (_bilinear_sample)
    const float x = ((u / (float)(W - 1) * 2.0f - 1.0f) + 1.0f) * ((float)W - 1.0f) / 2.0f;
    const float y = ((v / (float)(H - 1) * 2.0f - 1.0f) + 1.0f) * ((float)H - 1.0f) / 2.0f;
    // floor -> int32 as torch .int() does; coordinates beyond int range (or NaN) are pushed far outside instead
    const float xf = floorf(x), yf = floorf(y);
    const bool fin = (xf > -1.0e9f) && (xf < 1.0e9f) && (yf > -1.0e9f) && (yf < 1.0e9f);
    int x0 = fin ? (int)xf : -1000000, y0 = fin ? (int)yf : -1000000;
    int x1 = x0 + 1, y1 = y0 + 1;
    t.mask = (x0 >= 0 && x1 <= W - 1 && y0 >= 0 && y0 <= H - 1) ? 1.f : 0.f;
    x0 = max(0, min(x0, W - 1)); x1 = max(0, min(x1, W - 1));
    y0 = max(0, min(y0, H - 1)); y1 = max(0, min(y1, H - 1));
    t.ia = y0 * W + x0; t.ib = y1 * W + x0; t.ic = y0 * W + x1; t.id = y1 * W + x1;
    t.fx = (float)x1 - x;
    t.fy = (float)y1 - y;
}

