// output.cu — the inference output side: what happens to depth / confidence maps after the network.
//
//   mvs_upsample_nearest     F.interpolate(depth.unsqueeze(1), size=(1200, 1600)) of jdacs/eval_dense.py:150-153 (mode 'nearest',
//                            ATen's source index floor(dst * (float)in / out)), optionally writing the rows bottom-up, the order
//                            save_pfm stores them in (datasets/data_io.py:53-80 flips the image before tofile), so that the
//                            device buffer, copied to the host, IS the body of the .pfm file
//   mvs_depth_preview_u8     write_depth_img's 8-bit preview ((depth - 500) / 2 -> mode "L": clamp to [0, 255], truncate;
//                            eval_dense.py:110-121)
//   mvs_geo_consistency      reproject_with_depth + check_geometric_consistency (eval_dense.py:177-232): project every reference
//                            pixel into a source view with its depth, sample the source depth map there (cv2.remap, bilinear,
//                            constant-0 border), lift that sample back into the reference view and keep the pixel when it lands
//                            within `dist_thresh` pixels and `rel_thresh` relative depth of where it started.
//
// The reference does the last one in NumPy float64 on the host per (reference, source) pair of 1600x1200 maps; here it is one
// launch per batch of pairs, one thread per pixel, fp64 per-pixel algebra (same promotion points: the pixel grid is integer,
// depth float32, the products float64, the remap coordinates and the outputs float32).  cv2.remap's INTER_LINEAR on float
// coordinates quantises them to 1/32 pixel (cvRound(x * 32), round-half-even) and blends in float32 with the table weights
// (1 - fy)(1 - fx), (1 - fy) fx, fy (1 - fx), fy fx in that order: reproduced exactly, unfused.
#include "mvs_rt.h"

namespace {

// one thread = 4 consecutive output pixels of a row (one 16-byte store when the row length allows: the kernel is a pure write
// stream, 4-byte stores ran at 0.7 TB/s)
__global__ void __launch_bounds__(256)
upsample_nearest_kernel(const float* __restrict__ src, float* __restrict__ dst, int M, int H, int W, int Ho, int Wo, int flip_rows) {
    const int Wq = (Wo + 3) / 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * Ho * Wq) return;
    const int xq = (int)(i % Wq), y = (int)((i / Wq) % Ho), m = (int)(i / ((int64_t)Wq * Ho));
    // ATen nearest: scale = (float) in / out; src = min((int) floorf(dst * scale), in - 1)
    const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
    const int ys = min((int)floorf((float)y * sh), H - 1);
    const int yo = flip_rows ? Ho - 1 - y : y;
    const float* row = src + ((int64_t)m * H + ys) * W;
    float* o = dst + ((int64_t)m * Ho + yo) * Wo + 4 * xq;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int x = min(4 * xq + e, Wo - 1);
        v[e] = __ldg(row + min((int)floorf((float)x * sw), W - 1));
    }
#ifndef MVS_CPU_EMU
    if ((Wo & 3) == 0) { *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]); return; }
#endif
    for (int e = 0; e < 4 && 4 * xq + e < Wo; ++e) o[e] = v[e];
}

__global__ void __launch_bounds__(256)
depth_preview_u8_kernel(const float* __restrict__ depth, uint8_t* __restrict__ out, int64_t n, float offset, float scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = (__ldg(depth + i) - offset) / scale;       // (depth - 500) / 2
    out[i] = v <= 0.f ? (uint8_t)0 : (v >= 255.f ? (uint8_t)255 : (uint8_t)v);   // Pillow F -> L: clamp, truncate (NaN -> 0 as in C)
}

// cams [pair][60] doubles, prepared by the caller with the float32 matrix algebra of the reference (np.linalg.inv / np.matmul of
// float32 arrays) and widened: inv(K_ref) [0,9) | (E_src inv(E_ref))[:3] [9,21) | K_src [21,30) | inv(K_src) [30,39) |
// (E_ref inv(E_src))[:3] [39,51) | K_ref [51,60)

__device__ __forceinline__ int round_half_even_to_int(float v) {
#ifdef MVS_CPU_EMU
    return (int)lrintf(v);          // default rounding mode: to nearest, ties to even (what cvRound does)
#else
    return __float2int_rn(v);
#endif
}
__device__ __forceinline__ float mul_rn(float a, float b) {
#ifdef MVS_CPU_EMU
    volatile float r = a * b; return r;
#else
    return __fmul_rn(a, b);
#endif
}
__device__ __forceinline__ float add_rn(float a, float b) {
#ifdef MVS_CPU_EMU
    volatile float r = a + b; return r;
#else
    return __fadd_rn(a, b);
#endif
}

// cv2.remap(src, x, y, INTER_LINEAR) at one point, float32 single-channel source, BORDER_CONSTANT 0
__device__ __forceinline__ float remap_bilinear(const float* __restrict__ src, int H, int W, float x, float y) {
    // cvRound(x * INTER_TAB_SIZE); non-finite / huge values saturate to short range like saturate_cast<short>(sx >> 5)
    const float fx32 = x * 32.f, fy32 = y * 32.f;
    int sx, sy;
    if (!(fx32 > -2.0e9f && fx32 < 2.0e9f)) sx = INT32_MIN; else sx = round_half_even_to_int(fx32);
    if (!(fy32 > -2.0e9f && fy32 < 2.0e9f)) sy = INT32_MIN; else sy = round_half_even_to_int(fy32);
    const int ax = sx & 31, ay = sy & 31;
    int ix = sx >> 5, iy = sy >> 5;
    ix = max(-32768, min(32767, ix)); iy = max(-32768, min(32767, iy));
    const float tx = (float)ax * (1.f / 32.f), ty = (float)ay * (1.f / 32.f);
    const float w0 = (1.f - ty) * (1.f - tx), w1 = (1.f - ty) * tx, w2 = ty * (1.f - tx), w3 = ty * tx;
    if (ix >= W || ix + 1 < 0 || iy >= H || iy + 1 < 0) return 0.f;
    const bool x0 = ix >= 0, x1 = ix + 1 < W, y0 = iy >= 0, y1 = iy + 1 < H;
    const float v0 = (x0 && y0) ? __ldg(src + (int64_t)iy * W + ix) : 0.f;
    const float v1 = (x1 && y0) ? __ldg(src + (int64_t)iy * W + ix + 1) : 0.f;
    const float v2 = (x0 && y1) ? __ldg(src + (int64_t)(iy + 1) * W + ix) : 0.f;
    const float v3 = (x1 && y1) ? __ldg(src + (int64_t)(iy + 1) * W + ix + 1) : 0.f;
    return add_rn(add_rn(add_rn(mul_rn(v0, w0), mul_rn(v1, w1)), mul_rn(v2, w2)), mul_rn(v3, w3));
}

__device__ __forceinline__ void mat3_vec(const double* m, double a, double b, double c, double (&o)[3]) {
    for (int r = 0; r < 3; ++r) o[r] = m[r * 3] * a + m[r * 3 + 1] * b + m[r * 3 + 2] * c;
}
__device__ __forceinline__ void mat34_vec(const double* m, const double (&v)[3], double (&o)[3]) {
    for (int r = 0; r < 3; ++r) o[r] = m[r * 4] * v[0] + m[r * 4 + 1] * v[1] + m[r * 4 + 2] * v[2] + m[r * 4 + 3];
}

__global__ void __launch_bounds__(128)
geo_consistency_kernel(const float* __restrict__ depth_ref, const float* __restrict__ depth_src, const double* __restrict__ cams,
                       uint8_t* __restrict__ mask, float* __restrict__ depth_reproj, float* __restrict__ x_src_out,
                       float* __restrict__ y_src_out, float* __restrict__ x_rep_out, float* __restrict__ y_rep_out, int B, int H, int W,
                       float dist_thresh, float rel_thresh, int apply_mask) {
    const int HW = H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HW) return;
    const int p = (int)(i % HW), b = (int)(i / HW);
    const int px = p % W, py = p / W;
    const double* c = cams + (int64_t)b * 60;
    const float dref = __ldg(depth_ref + i);
    // step 1 (eval_dense.py:179-192): reference pixel -> reference camera space -> source camera space -> source pixel
    double xyz_ref[3], xyz_src[3], kx[3];
    mat3_vec(c, (double)px * (double)dref, (double)py * (double)dref, (double)dref, xyz_ref);   // inv(K_ref) @ ((x, y, 1) * depth)
    mat34_vec(c + 9, xyz_ref, xyz_src);
    mat3_vec(c + 21, xyz_src[0], xyz_src[1], xyz_src[2], kx);
    const double xs_d = kx[0] / kx[2], ys_d = kx[1] / kx[2];
    const float xs = (float)xs_d, ys = (float)ys_d;
    // step 2 (:194-214): sample the source depth there, lift it back into the reference view
    const float dsrc = remap_bilinear(depth_src + (int64_t)b * HW, H, W, xs, ys);
    double xyz_s2[3], xyz_r2[3], kr[3];
    mat3_vec(c + 30, xs_d * (double)dsrc, ys_d * (double)dsrc, (double)dsrc, xyz_s2);         // inv(K_src) @ ((x_src, y_src, 1) * sampled)
    mat34_vec(c + 39, xyz_s2, xyz_r2);
    const float drep = (float)xyz_r2[2];
    mat3_vec(c + 51, xyz_r2[0], xyz_r2[1], xyz_r2[2], kr);
    const float xr = (float)(kr[0] / kr[2]), yr = (float)(kr[1] / kr[2]);
    // check_geometric_consistency (:217-232): float32 arrays against the int64 pixel grid -> float64 differences
    const double dx = (double)xr - (double)px, dy = (double)yr - (double)py;
    const double dist = sqrt(dx * dx + dy * dy);
    const float ddiff = fabsf(drep - dref);
    const float rel = ddiff / dref;
    const bool ok = (dist < (double)dist_thresh) && (rel < rel_thresh);
    if (mask) mask[i] = ok ? 1 : 0;
    if (depth_reproj) depth_reproj[i] = (apply_mask && !ok) ? 0.f : drep;
    if (x_src_out) x_src_out[i] = xs;
    if (y_src_out) y_src_out[i] = ys;
    if (x_rep_out) x_rep_out[i] = xr;
    if (y_rep_out) y_rep_out[i] = yr;
}

}  // namespace

extern "C" int mvs_upsample_nearest(const float* src, float* dst, int M, int H, int W, int Ho, int Wo, int flip_rows, void* stream) {
    MVS_REQUIRE(src && dst, MVS_E_ARG, "mvs_upsample_nearest: null pointer");
    MVS_REQUIRE(M > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, MVS_E_SHAPE, "mvs_upsample_nearest: bad dims");
    MVS_LAUNCH(upsample_nearest_kernel, dim3(mvs_cdiv((int64_t)M * Ho * ((Wo + 3) / 4), 256)), dim3(256), stream, src, dst, M, H, W, Ho, Wo, flip_rows);
    return MVS_CHECK_LAUNCH("mvs_upsample_nearest");
}

extern "C" int mvs_depth_preview_u8(const float* depth, uint8_t* out, int64_t n, float offset, float scale, void* stream) {
    MVS_REQUIRE(depth && out, MVS_E_ARG, "mvs_depth_preview_u8: null pointer");
    MVS_REQUIRE(n > 0 && scale != 0.f, MVS_E_SHAPE, "mvs_depth_preview_u8: empty map or zero scale");
    MVS_LAUNCH(depth_preview_u8_kernel, dim3(mvs_cdiv(n, 256)), dim3(256), stream, depth, out, n, offset, scale);
    return MVS_CHECK_LAUNCH("mvs_depth_preview_u8");
}

extern "C" int mvs_geo_consistency(const float* depth_ref, const float* depth_src, const double* cams, uint8_t* mask,
                                   float* depth_reprojected, float* x_src, float* y_src, float* x_reprojected, float* y_reprojected,
                                   int B, int H, int W, float dist_thresh, float rel_thresh, int apply_mask, void* stream) {
    MVS_REQUIRE(depth_ref && depth_src && cams, MVS_E_ARG, "mvs_geo_consistency: null pointer");
    MVS_REQUIRE(B > 0 && H > 0 && W > 0 && H <= 32767 && W <= 32767, MVS_E_SHAPE, "mvs_geo_consistency: bad dims (maps up to 32767 x 32767, as cv2.remap)");
    MVS_LAUNCH(geo_consistency_kernel, dim3(mvs_cdiv((int64_t)B * H * W, 128)), dim3(128), stream, depth_ref, depth_src, cams, mask,
               depth_reprojected, x_src, y_src, x_reprojected, y_reprojected, B, H, W, dist_thresh, rel_thresh, apply_mask);
    return MVS_CHECK_LAUNCH("mvs_geo_consistency");
}
