// loss.cu — the self-supervised photometric loss (UnSupLoss) as five small launches forward and one backward.
//
// Reference: jdacs/losses/unsup_loss.py:24-83 (jdacs-ms twin :23-86) on top of losses/modules.py:17-90 and the photometric warp
// of losses/homography.py:186-238.  Per source view the reference runs x0.25 bilinear resize, ~40 launches of inverse_warping,
// 6 avg-pools of SSIM, two image gradients, three smooth-L1 means, then a stack / top-k over the views: ~250 launches forward
// and as many backward, all on 128x160 maps (launch-bound on a B200).  Here:
//
//   loss_setup_kernel   per (view, item) camera composition (fp64) and zeroed accumulators
//   loss_prep_kernel    [B,N,3,Hi,Wi] -> [N][B,H,W,3] NHWC views (the x0.25 bilinear resize = mean of the 2x2 centre pixels)
//   loss_warp_kernel    inverse_warping of every source view: warped [V][B,H,W,3], mask [V][B,H,W]
//   loss_terms_kernel   per (view, pixel): smooth-L1 of the masked photo / dx / dy residuals, 3x3 SSIM (views 1, 2) with the
//                       coefficients its backward needs, edge-aware depth smoothness -> fp64 sums
//   loss_topk_kernel    per pixel the three smallest of (reconstr_v + 1e4 (1 - mask_v)), masked at 1e4 -> fp64 sum and per-view
//                       selection counts (the gradient of the mean towards each view's scalar)
//   loss_final_kernel   out = {12 rec + 6 ssim + w smooth, rec, ssim, smooth}
//   loss_bwd_kernel     one thread per reference pixel: d/d depth of everything above (smoothness directly; photo / dx / dy /
//                       SSIM through the warped colours and the sampling coordinates), no atomics, fixed order
//
// Sums are accumulated in fp64 (warp shuffle, one atomic per warp and quantity), so the scalars are stable run to run to fp32
// rounding.  The quirks of the reference are kept: reference intrinsics on both sides (H6), mask from the unclamped x1 (H7),
// reconstr_v is ONE scalar per view broadcast over the pixels before the top-3 selection, SSIM on the unmasked warped image.
#include "mvs_rt.h"
#include "invwarp_dev.h"

namespace {

constexpr int kAccView = 4;                       // per view: photo, dx, dy, ssim
constexpr int kAccSmx = kAccView * MVS_MAX_SRC;   // 32
constexpr int kAccSmy = kAccSmx + 1;
constexpr int kAccTop = kAccSmx + 2;
constexpr int kAccCnt = kAccSmx + 4;              // [MVS_MAX_SRC] selection counts
constexpr int kAccN = kAccCnt + MVS_MAX_SRC;      // 44 (MVS_LOSS_ACC_DOUBLES = 48)

__device__ __forceinline__ float sl1(float z) { const float a = fabsf(z); return a < 1.f ? 0.5f * z * z : a - 0.5f; }
__device__ __forceinline__ float sl1_grad(float z) { return fabsf(z) < 1.f ? z : (z > 0.f ? 1.f : -1.f); }
__device__ __forceinline__ float sgn(float z) { return z > 0.f ? 1.f : (z < 0.f ? -1.f : 0.f); }

// one fp64 sum per quantity: shuffle-reduce the warp, lane 0 adds (serial emulation: every thread adds)
__device__ __forceinline__ void acc_add(double* a, double v) {
#ifndef MVS_CPU_EMU
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(a, v);
#else
    if (v != 0.0) atomicAdd(a, v);
#endif
}

__global__ void loss_setup_kernel(const float* __restrict__ cams, float* __restrict__ cam_ws, double* __restrict__ acc, int B, int N) {
    const int V = N - 1;
    for (int i = threadIdx.x; i < V * B; i += blockDim.x) {
        const int v = i / B, b = i - v * B;
        const float* base = cams + (int64_t)b * N * 32;
        invwarp_compose_cam(base, base + (int64_t)(v + 1) * 32, cam_ws + (int64_t)i * 24);
    }
    for (int i = threadIdx.x; i < MVS_LOSS_ACC_DOUBLES; i += blockDim.x) acc[i] = 0.0;
}

// F.interpolate(scale_factor=0.25, mode='bilinear') samples at 4 d + 1.5: the mean of source pixels 4d+1, 4d+2 in each axis
__global__ void __launch_bounds__(256)
loss_prep_kernel(const float* __restrict__ imgs, float* __restrict__ small, int B, int N, int Hi, int Wi, int H, int W, int scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * B * H * W) return;
    const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)((i / ((int64_t)W * H)) % B), n = (int)(i / ((int64_t)W * H * B));
    const float* im = imgs + ((int64_t)b * N + n) * 3 * Hi * Wi;
    for (int c = 0; c < 3; ++c) {
        const float* pl = im + (int64_t)c * Hi * Wi;
        float v;
        if (scale == 1) v = __ldg(pl + (int64_t)y * Wi + x);
        else {
            const float* q = pl + (int64_t)(4 * y + 1) * Wi + 4 * x + 1;
            v = 0.5f * (0.5f * __ldg(q) + 0.5f * __ldg(q + 1)) + 0.5f * (0.5f * __ldg(q + Wi) + 0.5f * __ldg(q + Wi + 1));
        }
        small[i * 3 + c] = v;
    }
}

__global__ void __launch_bounds__(128)
loss_warp_kernel(const float* __restrict__ small, const float* __restrict__ depth, const float* __restrict__ cam_ws,
                 float* __restrict__ warped, float* __restrict__ mask, int B, int V, int H, int W) {
    const int HW = H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)V * B * HW) return;
    const int p = (int)(i % HW), b = (int)((i / HW) % B), v = (int)(i / ((int64_t)HW * B));
    InvTaps t;
    invwarp_taps(cam_ws + ((int64_t)v * B + b) * 24, __ldg(depth + (int64_t)b * HW + p), p % W, p / W, H, W, t);
    const float wa = t.fx * t.fy, wb = t.fx * (1.f - t.fy), wc = (1.f - t.fx) * t.fy, wd = (1.f - t.fx) * (1.f - t.fy);
    const float* im = small + ((int64_t)(v + 1) * B + b) * HW * 3;
    for (int c = 0; c < 3; ++c)
        warped[i * 3 + c] = wa * __ldg(im + (int64_t)t.ia * 3 + c) + wb * __ldg(im + (int64_t)t.ib * 3 + c) +
                            wc * __ldg(im + (int64_t)t.ic * 3 + c) + wd * __ldg(im + (int64_t)t.id * 3 + c);
    mask[i] = t.mask;
}

// grid.y = view (0..V-1) or V = the smoothness term.  coef [min(V,2)][B,H,W,3][3]: d ssim_c / d warped_q = k0 + k1 ref_q + k2 warped_q
__global__ void __launch_bounds__(128)
loss_terms_kernel(const float* __restrict__ small, const float* __restrict__ depth, const float* __restrict__ warped,
                  const float* __restrict__ mask, float* __restrict__ coef, double* __restrict__ acc, int B, int V, int H, int W,
                  float lambda) {
    const int HW = H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < (int64_t)B * HW;
    const int p = live ? (int)(i % HW) : 0, b = live ? (int)(i / HW) : 0;
    const int y = p / W, x = p - y * W;
    const int v = blockIdx.y;
    const float* ref = small + (int64_t)b * HW * 3;          // view 0
    if (v == V) {
        // ---- depth_smoothness (modules.py:67-77): mean |(d[x] - d[x+1]) exp(-lambda mean_c |I[x] - I[x+1]|)| + the same along y
        double sx = 0.0, sy = 0.0;
        if (live) {
            const float* dp = depth + (int64_t)b * HW;
            const float d0 = __ldg(dp + p);
            for (int dir = 0; dir < 2; ++dir) {
                const int q = dir == 0 ? p + 1 : p + W;
                if (dir == 0 ? x < W - 1 : y < H - 1) {
                    float g = 0.f;
                    for (int c = 0; c < 3; ++c) g += fabsf(__ldg(ref + (int64_t)p * 3 + c) - __ldg(ref + (int64_t)q * 3 + c));
                    const float wgt = expf(-(lambda * (g / 3.f)));
                    const float s = fabsf((d0 - __ldg(dp + q)) * wgt);
                    if (dir == 0) sx = s; else sy = s;
                }
            }
        }
        acc_add(acc + kAccSmx, sx);
        acc_add(acc + kAccSmy, sy);
        return;
    }
    double photo = 0.0, gdx = 0.0, gdy = 0.0, ss = 0.0;
    if (live) {
        const float* wv = warped + ((int64_t)v * B + b) * HW * 3;
        const float* mv = mask + ((int64_t)v * B + b) * HW;
        const float m0 = __ldg(mv + p);
        float wm[3], rm[3];
        for (int c = 0; c < 3; ++c) { wm[c] = __ldg(wv + (int64_t)p * 3 + c) * m0; rm[c] = __ldg(ref + (int64_t)p * 3 + c) * m0; }
        float s = 0.f;
        for (int c = 0; c < 3; ++c) s += sl1(wm[c] - rm[c]);
        photo = s;
        for (int dir = 0; dir < 2; ++dir) {
            if (dir == 0 ? x < W - 1 : y < H - 1) {
                const int q = dir == 0 ? p + 1 : p + W;
                const float m1 = __ldg(mv + q);
                float t = 0.f;
                for (int c = 0; c < 3; ++c) {
                    const float a = __ldg(wv + (int64_t)q * 3 + c) * m1 - wm[c], r = __ldg(ref + (int64_t)q * 3 + c) * m1 - rm[c];
                    t += sl1(a - r);
                }
                if (dir == 0) gdx = t; else gdy = t;
            }
        }
        if (v < 2) {
            // ---- SSIM (modules.py:17-52) at window centre (y, x): x = reference, y = warped (unmasked), mask pooled
            float k[3][3];
            for (int c = 0; c < 3; ++c) k[c][0] = k[c][1] = k[c][2] = 0.f;
            if (y >= 1 && y <= H - 2 && x >= 1 && x <= W - 2) {
                float sm = 0.f, sx[3] = {0.f, 0.f, 0.f}, sy[3] = {0.f, 0.f, 0.f}, sxx[3] = {0.f, 0.f, 0.f}, syy[3] = {0.f, 0.f, 0.f}, sxy[3] = {0.f, 0.f, 0.f};
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int q = p + dy * W + dx;
                        sm += __ldg(mv + q);
                        for (int c = 0; c < 3; ++c) {
                            const float a = __ldg(ref + (int64_t)q * 3 + c), bq = __ldg(wv + (int64_t)q * 3 + c);
                            sx[c] += a; sy[c] += bq; sxx[c] += a * a; syy[c] += bq * bq; sxy[c] += a * bq;
                        }
                    }
                const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
                const float mp = sm / 9.f;
                float tot = 0.f;
                for (int c = 0; c < 3; ++c) {
                    const float mx = sx[c] / 9.f, my = sy[c] / 9.f;
                    const float vx = sxx[c] / 9.f - mx * mx, vy = syy[c] / 9.f - my * my, cxy = sxy[c] / 9.f - mx * my;
                    const float a1 = 2.f * mx * my + C1, a2 = 2.f * cxy + C2, b1 = mx * mx + my * my + C1, b2 = vx + vy + C2;
                    const float n = a1 * a2, d = b1 * b2;
                    const float sv = (1.f - n / d) * 0.5f;
                    tot += mp * fminf(fmaxf(sv, 0.f), 1.f);
                    if (sv >= 0.f && sv <= 1.f) {
                        // d n / d y_q = n0 + n1 x_q ; d d / d y_q = d0 + d1 y_q  (y_q = warped, x_q = reference at a window pixel q)
                        const float n0 = (2.f / 9.f) * (mx * a2 - a1 * mx), n1 = (2.f / 9.f) * a1;
                        const float d0 = (2.f / 9.f) * (my * b2 - b1 * my), d1 = (2.f / 9.f) * b1;
                        const float f = -0.5f * mp / (d * d);
                        k[c][0] = f * (d * n0 - n * d0); k[c][1] = f * d * n1; k[c][2] = -f * n * d1;
                    }
                }
                ss = tot;
            }
            float* kc = coef + (((int64_t)v * B + b) * HW + p) * 9;
            for (int c = 0; c < 3; ++c) { kc[c * 3] = k[c][0]; kc[c * 3 + 1] = k[c][1]; kc[c * 3 + 2] = k[c][2]; }
        }
    }
    acc_add(acc + v * kAccView + 0, photo);
    acc_add(acc + v * kAccView + 1, gdx);
    acc_add(acc + v * kAccView + 2, gdy);
    if (v < 2) acc_add(acc + v * kAccView + 3, ss);
}

// compute_reconstr_loss (modules.py:80-90) of view v from the sums: 0.5 photo + 0.5 (dx + dy), each a mean
__device__ __forceinline__ float view_scalar(const double* __restrict__ acc, int v, int B, int H, int W) {
    const double P = (double)B * H * W;
    const float photo = (float)(acc[v * kAccView] / (3.0 * P));
    const float gx = (float)(acc[v * kAccView + 1] / (3.0 * (double)B * H * (W - 1)));
    const float gy = (float)(acc[v * kAccView + 2] / (3.0 * (double)B * (H - 1) * W));
    return 0.5f * photo + 0.5f * (gx + gy);
}

template <int MAXV>
__global__ void __launch_bounds__(128)
loss_topk_kernel(const float* __restrict__ mask, double* __restrict__ acc, int B, int V, int H, int W) {
    const int HW = H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < (int64_t)B * HW;
    float val[MAXV];
    bool taken[MAXV], sel[MAXV];      // among the three smallest / and below 1e4 (a valid view: it carries gradient)
    double top = 0.0;
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
        taken[v] = sel[v] = false;
        val[v] = (live && v < V) ? view_scalar(acc, v, B, H, W) + 1e4f * (1.f - __ldg(mask + (int64_t)v * B * HW + i)) : 3.0e38f;
    }
    if (live) {
        for (int r = 0; r < 3; ++r) {           // torch.topk(-vol, 3): the three smallest (unsorted), then zero those >= 1e4
            float best = 3.4e38f;
            int bi = -1;
#pragma unroll
            for (int v = 0; v < MAXV; ++v) if (v < V && !taken[v] && val[v] < best) { best = val[v]; bi = v; }
#pragma unroll
            for (int v = 0; v < MAXV; ++v) if (v == bi) { taken[v] = true; if (val[v] < 1e4f) { top += (double)val[v]; sel[v] = true; } }
        }
    }
    acc_add(acc + kAccTop, top);
#pragma unroll
    for (int v = 0; v < MAXV; ++v) if (v < V) acc_add(acc + kAccCnt + v, (live && sel[v]) ? 1.0 : 0.0);
}

__global__ void loss_final_kernel(const double* __restrict__ acc, float* __restrict__ out, int B, int V, int H, int W, float w_smooth) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double P = (double)B * H * W;
    const float rec = (float)(acc[kAccTop] / P);
    float ssim = 0.f;
    for (int v = 0; v < V && v < 2; ++v) ssim += (float)(acc[v * kAccView + 3] / (3.0 * (double)B * (H - 2) * (W - 2)));
    const float smooth = (float)(acc[kAccSmx] / ((double)B * H * (W - 1))) + (float)(acc[kAccSmy] / ((double)B * (H - 1) * W));
    out[1] = rec; out[2] = ssim; out[3] = smooth;
    out[0] = 12.f * rec + 6.f * ssim + w_smooth * smooth;
}

__global__ void __launch_bounds__(128)
loss_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ small, const float* __restrict__ depth,
                const float* __restrict__ warped, const float* __restrict__ mask, const float* __restrict__ coef,
                const float* __restrict__ cam_ws, const double* __restrict__ acc, float* __restrict__ gdepth, int B, int V, int H,
                int W, float lambda, float w_smooth) {
    const int HW = H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HW) return;
    const int p = (int)(i % HW), b = (int)(i / HW);
    const int y = p / W, x = p - y * W;
    const float g0 = __ldg(gout);
    const float g_rec = __ldg(gout + 1) + 12.f * g0, g_ssim = __ldg(gout + 2) + 6.f * g0, g_sm = __ldg(gout + 3) + w_smooth * g0;
    const float* ref = small + (int64_t)b * HW * 3;
    const float* dp = depth + (int64_t)b * HW;
    const float d0 = __ldg(dp + p);
    float gd = 0.f;
    // ---- smoothness: pairs (p, p+1), (p-1, p) along x and the same along y
    {
        const float inv_mx = 1.f / ((float)B * H * (W - 1)), inv_my = 1.f / ((float)B * (H - 1) * W);
        for (int e = 0; e < 4; ++e) {
            const int dir = e >> 1, back = e & 1;                 // back: p is the second pixel of the pair
            const int st = dir == 0 ? 1 : W;
            const bool ok = dir == 0 ? (back ? x > 0 : x < W - 1) : (back ? y > 0 : y < H - 1);
            if (!ok) continue;
            const int a = back ? p - st : p, q = a + st;
            float g = 0.f;
            for (int c = 0; c < 3; ++c) g += fabsf(__ldg(ref + (int64_t)a * 3 + c) - __ldg(ref + (int64_t)q * 3 + c));
            const float wgt = expf(-(lambda * (g / 3.f)));
            const float t = (__ldg(dp + a) - __ldg(dp + q)) * wgt;
            gd += (back ? -1.f : 1.f) * sgn(t) * wgt * (dir == 0 ? inv_mx : inv_my);
        }
        gd *= g_sm;
    }
    const float P = (float)B * H * W;
    const float c_photo = 0.5f / (3.f * P), c_dx = 0.5f / (3.f * (float)B * H * (W - 1)), c_dy = 0.5f / (3.f * (float)B * (H - 1) * W);
    const float c_ssim = g_ssim / (3.f * (float)B * (H - 2) * (W - 2));
    for (int v = 0; v < V; ++v) {
        const float* wv = warped + ((int64_t)v * B + b) * HW * 3;
        const float* mv = mask + ((int64_t)v * B + b) * HW;
        const float m0 = __ldg(mv + p);
        const float c_r = g_rec * (float)(acc[kAccCnt + v] / (double)P);     // d loss / d reconstr_v
        float gw[3] = {0.f, 0.f, 0.f};
        float w0[3], r0[3];
        for (int c = 0; c < 3; ++c) { w0[c] = __ldg(wv + (int64_t)p * 3 + c); r0[c] = __ldg(ref + (int64_t)p * 3 + c); }
        if (m0 != 0.f && c_r != 0.f) {
            for (int c = 0; c < 3; ++c) gw[c] += c_photo * sl1_grad(w0[c] * m0 - r0[c] * m0);
            for (int e = 0; e < 4; ++e) {
                const int dir = e >> 1, back = e & 1;
                const int st = dir == 0 ? 1 : W;
                const bool ok = dir == 0 ? (back ? x > 0 : x < W - 1) : (back ? y > 0 : y < H - 1);
                if (!ok) continue;
                const int q = back ? p - st : p + st;
                const float m1 = __ldg(mv + q);
                for (int c = 0; c < 3; ++c) {
                    const float wq = __ldg(wv + (int64_t)q * 3 + c) * m1, rq = __ldg(ref + (int64_t)q * 3 + c) * m1;
                    // residual of the pair: (second - first) of warped minus the same of the reference
                    const float z = back ? ((w0[c] * m0 - wq) - (r0[c] * m0 - rq)) : ((wq - w0[c] * m0) - (rq - r0[c] * m0));
                    gw[c] += (back ? 1.f : -1.f) * sl1_grad(z) * (dir == 0 ? c_dx : c_dy);
                }
            }
            for (int c = 0; c < 3; ++c) gw[c] *= c_r * m0;
        }
        if (v < 2 && c_ssim != 0.f) {
            const float* kv = coef + ((int64_t)v * B + b) * HW * 9;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int yy = y + dy, xx = x + dx;
                    if (yy < 1 || yy > H - 2 || xx < 1 || xx > W - 2) continue;
                    const float* kc = kv + (int64_t)(yy * W + xx) * 9;
                    for (int c = 0; c < 3; ++c) gw[c] += c_ssim * (__ldg(kc + c * 3) + __ldg(kc + c * 3 + 1) * r0[c] + __ldg(kc + c * 3 + 2) * w0[c]);
                }
        }
        if (gw[0] == 0.f && gw[1] == 0.f && gw[2] == 0.f) continue;
        // ---- through the sampler to depth (as invwarp_bwd_kernel)
        InvTaps t;
        invwarp_taps(cam_ws + ((int64_t)v * B + b) * 24, d0, x, y, H, W, t);
        const float* im = small + ((int64_t)(v + 1) * B + b) * HW * 3;
        float gfx = 0.f, gfy = 0.f;
        for (int c = 0; c < 3; ++c) {
            const float a = __ldg(im + (int64_t)t.ia * 3 + c), bb = __ldg(im + (int64_t)t.ib * 3 + c);
            const float cc = __ldg(im + (int64_t)t.ic * 3 + c), dd = __ldg(im + (int64_t)t.id * 3 + c);
            gfx += gw[c] * (t.fy * a + (1.f - t.fy) * bb - t.fy * cc - (1.f - t.fy) * dd);
            gfy += gw[c] * (t.fx * a - t.fx * bb + (1.f - t.fx) * cc - (1.f - t.fx) * dd);
        }
        gd += -(gfx * t.dx_dd + gfy * t.dy_dd);
    }
    gdepth[i] = gd;
}

int check_loss(const char* who, int B, int N, int Hi, int Wi, int H, int W) {
    MVS_REQUIRE(B > 0 && H > 2 && W > 2, MVS_E_SHAPE, "%s: bad dims B=%d H=%d W=%d", who, B, H, W);
    MVS_REQUIRE(N - 1 >= 3, MVS_E_SHAPE, "%s: the top-3 view selection needs at least 3 source views (got %d views; selected index k out of range)", who, N);
    MVS_REQUIRE(N - 1 <= MVS_MAX_SRC, MVS_E_SHAPE, "%s: at most %d source views", who, MVS_MAX_SRC);
    MVS_REQUIRE((Hi == H && Wi == W) || (Hi / 4 == H && Wi / 4 == W), MVS_E_SHAPE,
                "%s: images must be at the depth map's size or 4x it (images %dx%d, depth %dx%d)", who, Hi, Wi, H, W);
    return MVS_OK;
}

}  // namespace

extern "C" int mvs_unsup_loss_fwd(const float* imgs, const float* cams, const float* depth, int B, int N, int Hi, int Wi, int H, int W,
                                  float smooth_lambda, float smooth_weight, float* small, float* warped, float* mask, float* coef,
                                  float* cam_ws, double* acc, float* out, void* stream) {
    MVS_REQUIRE(imgs && cams && depth && small && warped && mask && coef && cam_ws && acc && out, MVS_E_ARG, "mvs_unsup_loss_fwd: null pointer");
    int rc = check_loss("mvs_unsup_loss_fwd", B, N, Hi, Wi, H, W);
    if (rc) return rc;
    const int V = N - 1;
    const int64_t P = (int64_t)B * H * W;
    MVS_LAUNCH(loss_setup_kernel, dim3(1), dim3(64), stream, cams, cam_ws, acc, B, N);
    MVS_LAUNCH(loss_prep_kernel, dim3(mvs_cdiv(P * N, 256)), dim3(256), stream, imgs, small, B, N, Hi, Wi, H, W, Hi == H ? 1 : 4);
    MVS_LAUNCH(loss_warp_kernel, dim3(mvs_cdiv(P * V, 128)), dim3(128), stream, small, depth, cam_ws, warped, mask, B, V, H, W);
    MVS_LAUNCH(loss_terms_kernel, dim3(mvs_cdiv(P, 128), (unsigned)(V + 1)), dim3(128), stream, small, depth, warped, mask, coef, acc, B, V, H, W, smooth_lambda);
    MVS_LAUNCH(loss_topk_kernel<MVS_MAX_SRC>, dim3(mvs_cdiv(P, 128)), dim3(128), stream, mask, acc, B, V, H, W);
    MVS_LAUNCH(loss_final_kernel, dim3(1), dim3(32), stream, acc, out, B, V, H, W, smooth_weight);
    return MVS_CHECK_LAUNCH("mvs_unsup_loss_fwd");
}

extern "C" int mvs_unsup_loss_bwd(const float* grad_out, const float* small, const float* depth, const float* warped, const float* mask,
                                  const float* coef, const float* cam_ws, const double* acc, float* grad_depth, int B, int N, int H, int W,
                                  float smooth_lambda, float smooth_weight, void* stream) {
    MVS_REQUIRE(grad_out && small && depth && warped && mask && coef && cam_ws && acc && grad_depth, MVS_E_ARG, "mvs_unsup_loss_bwd: null pointer");
    int rc = check_loss("mvs_unsup_loss_bwd", B, N, H, W, H, W);
    if (rc) return rc;
    MVS_LAUNCH(loss_bwd_kernel, dim3(mvs_cdiv((int64_t)B * H * W, 128)), dim3(128), stream, grad_out, small, depth, warped, mask, coef,
               cam_ws, acc, grad_depth, B, N - 1, H, W, smooth_lambda, smooth_weight);
    return MVS_CHECK_LAUNCH("mvs_unsup_loss_bwd");
}
