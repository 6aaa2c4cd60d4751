// warp.cu — plane-sweep homography warp.
//   * mvs_warp_var_fwd/bwd : fused warp + bilinear gather + running sum / sum-of-squares -> variance volume.
//     The per-source warped volumes, the sampling grid and S1/S2 never exist in memory: one pass reads every
//     feature map (L2 resident, 1.3-2.6 MB each) and writes the variance volume once.
//   * mvs_homo_warp_fwd/bwd: the stand-alone homo_warping() API (NCHW in, NCDHW out) for drop-in parity.
// Reference arithmetic: jdacs/models/module.py:105-140, mvsnet.py:120-136; jdacs-ms/models/modules.py:62-104,
// 209-261, network.py:114-137.  Sampler semantics: ATen grid_sampler_2d bilinear / zeros padding.
#include "mvs_rt.h"
#include <stdlib.h>
#include <string.h>
#ifndef MVS_CPU_EMU
#include <cuda.h>
#endif

struct SrcPtrs { const void* p[MVS_MAX_SRC]; };
struct GradPtrs { float* p[MVS_MAX_SRC]; };

// ray = rot @ (x, y, 1): depth-independent part of the homography (module.py:125).
__device__ __forceinline__ void pixel_ray(const float* __restrict__ rt, float x, float y, float (&ray)[3]) {
    ray[0] = __ldg(rt + 0) * x + __ldg(rt + 1) * y + __ldg(rt + 2);
    ray[1] = __ldg(rt + 3) * x + __ldg(rt + 4) * y + __ldg(rt + 5);
    ray[2] = __ldg(rt + 6) * x + __ldg(rt + 7) * y + __ldg(rt + 8);
}

// (ray * depth + trans) -> projective divide -> normalise as the reference does -> un-normalise as
// grid_sample does (module.py:126-134 then ATen grid_sampler_unnormalize; hazard H1).
__device__ __forceinline__ void source_coord(const float (&ray)[3], const float* __restrict__ rt, float depth, int H,
                                             int W, int align_corners, float& ix, float& iy) {
    const float px = ray[0] * depth + __ldg(rt + 9);
    const float py = ray[1] * depth + __ldg(rt + 10);
    const float pz = ray[2] * depth + __ldg(rt + 11);
    const float u = px / pz, v = py / pz;
    const float xn = u / ((float)(W - 1) * 0.5f) - 1.f;
    const float yn = v / ((float)(H - 1) * 0.5f) - 1.f;
    if (align_corners) {
        ix = (xn + 1.f) / 2.f * (float)(W - 1);
        iy = (yn + 1.f) / 2.f * (float)(H - 1);
    } else {
        ix = ((xn + 1.f) * (float)W - 1.f) / 2.f;
        iy = ((yn + 1.f) * (float)H - 1.f) / 2.f;
    }
}

// The four bilinear taps of (ix, iy): pixel offset y*W+x (or -1 when the tap is outside the map; NaN and
// inf coordinates fail every comparison and drop all taps) and weight, in ATen order nw, ne, sw, se.
struct Taps { int off[4]; float w[4]; };
__device__ __forceinline__ void bilinear_taps(float ix, float iy, int H, int W, Taps& t, int pitch = 0) {
    if (pitch == 0) pitch = W;
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float x1 = x0 + 1.f, y1 = y0 + 1.f;
    const bool vx0 = (x0 >= 0.f) && (x0 <= (float)(W - 1)), vx1 = (x1 >= 0.f) && (x1 <= (float)(W - 1));
    const bool vy0 = (y0 >= 0.f) && (y0 <= (float)(H - 1)), vy1 = (y1 >= 0.f) && (y1 <= (float)(H - 1));
    const float wx0 = x1 - ix, wx1 = ix - x0, wy0 = y1 - iy, wy1 = iy - y0;
    const int xi = vx0 ? (int)x0 : (vx1 ? (int)x1 - 1 : 0);
    const int yi = vy0 ? (int)y0 : (vy1 ? (int)y1 - 1 : 0);
    const int base = yi * pitch + xi;
    t.off[0] = (vx0 && vy0) ? base : -1;              t.w[0] = wx0 * wy0;
    t.off[1] = (vx1 && vy0) ? base + 1 : -1;          t.w[1] = wx1 * wy0;
    t.off[2] = (vx0 && vy1) ? base + pitch : -1;      t.w[2] = wx0 * wy1;
    t.off[3] = (vx1 && vy1) ? base + pitch + 1 : -1;  t.w[3] = wx1 * wy1;
}

template <typename T>
__device__ __forceinline__ void gather8(const T* __restrict__ map, const Taps& t, float (&out)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (t.off[i] >= 0) {
            float v[8];
            V8<T>::load(map + (int64_t)t.off[i] * 8, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) out[k] += v[k] * t.w[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------ fused forward
// thread = (pixel, channel-block, depth chunk).  Lanes run along w, so the gathers of a warp touch
// neighbouring source pixels and each plane is stored as 32 consecutive 16/32-byte vectors.
template <typename TI, typename TO>
__global__ void __launch_bounds__(128)
warp_var_fwd_kernel(const TI* __restrict__ ref, SrcPtrs srcs, int nsrc, const float* __restrict__ rt,
                    const float* __restrict__ depth, int per_pixel, TO* __restrict__ var, int B, int CB, int D, int H,
                    int W, int dper, int align_corners, int ref_sq_in_sum, int pad) {
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int nchunk = (D + dper - 1) / dper;
    int y = blockIdx.y;
    const int dc = y % nchunk; y /= nchunk;
    const int cb = y % CB;
    const int b = y / CB;
    const float fx = (float)(p % W), fy = (float)(p / W);
    // zero-bordered maps (pad): rows of W + 2 pixels, H + 3 rows, pixel (0, 0) at row 1, column 1
    const int pitch = pad ? W + 2 : W, org = pad ? pitch + 1 : 0;
    const int64_t map_off = (((int64_t)b * CB + cb) * (pad ? (int64_t)(H + 3) * pitch : (int64_t)HW) + org) * 8;
    const int pp = (p / W) * pitch + (p % W);

    float r[8], r2[8];
    V8<TI>::load(ref + map_off + (int64_t)pp * 8, r);
#pragma unroll
    for (int k = 0; k < 8; ++k) r2[k] = r[k] * r[k];

    float ray[MVS_MAX_SRC][3];
#pragma unroll
    for (int s = 0; s < MVS_MAX_SRC; ++s)
        if (s < nsrc) pixel_ray(rt + ((int64_t)s * B + b) * 12, fx, fy, ray[s]);

    const float n = (float)(nsrc + 1);
    const int d_end = min(D, (dc + 1) * dper);
    for (int d = dc * dper; d < d_end; ++d) {
        const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
        float s1[8], s2[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { s1[k] = ref_sq_in_sum ? r2[k] : r[k]; s2[k] = r2[k]; }
#pragma unroll
        for (int s = 0; s < MVS_MAX_SRC; ++s) {
            if (s < nsrc) {
                float ix, iy;
                source_coord(ray[s], rt + ((int64_t)s * B + b) * 12, dv, H, W, align_corners, ix, iy);
                Taps t;
                bilinear_taps(ix, iy, H, W, t, pitch);
                float wv[8];
                gather8<TI>(reinterpret_cast<const TI*>(srcs.p[s]) + map_off, t, wv);
#pragma unroll
                for (int k = 0; k < 8; ++k) { s1[k] += wv[k]; s2[k] += wv[k] * wv[k]; }
            }
        }
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float m = s1[k] / n; o[k] = s2[k] / n - m * m; }
        V8<TO>::store(var + ((((int64_t)b * CB + cb) * D + d) * HW + p) * 8, o);
    }
}

// ------------------------------------------------------------------------------------------------ fused backward
// d var / d w_i = 2 w_i / N - 2 S1 / N^2 ;  d var / d ref = 2 r / N - 2 S1 / N^2  (x 2r on the S1 term when the
// reference feature entered S1 squared).  S1 is recomputed; each source is then re-sampled and its gradient
// scattered to the 4 taps with fp32 atomics (the grid carries no gradient: module.py:115, hazard H13).
template <typename TI, typename TO>
__global__ void __launch_bounds__(128)
warp_var_bwd_kernel(const TO* __restrict__ gvar, const TI* __restrict__ ref, SrcPtrs srcs, int nsrc,
                    const float* __restrict__ rt, const float* __restrict__ depth, int per_pixel,
                    float* __restrict__ gref, GradPtrs gsrcs, int B, int CB, int D, int H, int W, int dper,
                    int align_corners, int ref_sq_in_sum, int pad) {
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int nchunk = (D + dper - 1) / dper;
    int y = blockIdx.y;
    const int dc = y % nchunk; y /= nchunk;
    const int cb = y % CB;
    const int b = y / CB;
    const float fx = (float)(p % W), fy = (float)(p / W);
    // feature maps may be zero-bordered (pad, see the forward kernel); gradient maps never are
    const int pitch = pad ? W + 2 : W, org = pad ? pitch + 1 : 0;
    const int64_t map_off = (((int64_t)b * CB + cb) * (pad ? (int64_t)(H + 3) * pitch : (int64_t)HW) + org) * 8;
    const int64_t grad_off = ((int64_t)b * CB + cb) * HW * 8;
    const int pp = (p / W) * pitch + (p % W);

    float r[8];
    V8<TI>::load(ref + map_off + (int64_t)pp * 8, r);
    float ray[MVS_MAX_SRC][3];
#pragma unroll
    for (int s = 0; s < MVS_MAX_SRC; ++s)
        if (s < nsrc) pixel_ray(rt + ((int64_t)s * B + b) * 12, fx, fy, ray[s]);

    const float n = (float)(nsrc + 1);
    const float inv_n = 1.f / n;
    float gr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) gr[k] = 0.f;

    const int d_end = min(D, (dc + 1) * dper);
    for (int d = dc * dper; d < d_end; ++d) {
        const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
        float g[8];
        V8<TO>::load(gvar + ((((int64_t)b * CB + cb) * D + d) * HW + p) * 8, g);
        float s1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s1[k] = ref_sq_in_sum ? r[k] * r[k] : r[k];
#pragma unroll
        for (int s = 0; s < MVS_MAX_SRC; ++s) {
            if (s < nsrc) {
                float ix, iy;
                source_coord(ray[s], rt + ((int64_t)s * B + b) * 12, dv, H, W, align_corners, ix, iy);
                Taps t;
                bilinear_taps(ix, iy, H, W, t, pitch);
                float wv[8];
                gather8<TI>(reinterpret_cast<const TI*>(srcs.p[s]) + map_off, t, wv);
#pragma unroll
                for (int k = 0; k < 8; ++k) s1[k] += wv[k];
            }
        }
        float m2[8];  // 2 * S1 / N^2
#pragma unroll
        for (int k = 0; k < 8; ++k) m2[k] = 2.f * s1[k] * inv_n * inv_n;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float dr = ref_sq_in_sum ? (2.f * r[k] * inv_n - m2[k] * 2.f * r[k]) : (2.f * r[k] * inv_n - m2[k]);
            gr[k] += g[k] * dr;
        }
#pragma unroll
        for (int s = 0; s < MVS_MAX_SRC; ++s) {
            if (s < nsrc && gsrcs.p[s] != nullptr) {
                float ix, iy;
                source_coord(ray[s], rt + ((int64_t)s * B + b) * 12, dv, H, W, align_corners, ix, iy);
                Taps t, tg;
                bilinear_taps(ix, iy, H, W, t, pitch);    // gather offsets (feature maps)
                bilinear_taps(ix, iy, H, W, tg, W);       // scatter offsets (gradient maps)
                float wv[8];
                gather8<TI>(reinterpret_cast<const TI*>(srcs.p[s]) + map_off, t, wv);
                float c[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) c[k] = g[k] * (2.f * wv[k] * inv_n - m2[k]);
                float* gs = gsrcs.p[s] + grad_off;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (tg.off[i] >= 0) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) atomicAdd(gs + (int64_t)tg.off[i] * 8 + k, c[k] * tg.w[i]);
                    }
                }
            }
        }
    }
    if (gref != nullptr) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(gref + grad_off + (int64_t)p * 8 + k, gr[k]);
    }
}

#ifndef MVS_CPU_EMU
// ------------------------------------------------------------------------------------------------ fused forward, 16-bit storage
// Same arithmetic as warp_var_fwd_kernel, arranged for instruction throughput (the kernel is issue-bound long before
// it is HBM-bound): a thread owns CPT channel blocks of its pixel so the homography runs once per CPT*8 channels;
// the four taps are blended with packed half2 / bfloat162 FMAs (one rounding to the storage type per FMA, the same
// size as the storage rounding of the inputs); running sum / sum of squares / variance stay fp32 but use the packed
// f32x2 pipe (FFMA2); projective divide by reciprocal; the reference's normalise / un-normalise pair is folded into one FMA.
template <typename T> struct Pack2;
template <> struct Pack2<__half> {
    typedef __half2 type;
    static __device__ __forceinline__ __half2 splat(float w) { return __float2half2_rn(w); }
    static __device__ __forceinline__ float2 to_f2(__half2 v) { return __half22float2(v); }
    static __device__ __forceinline__ __half2 from_f2(float2 v) { return __float22half2_rn(v); }
    static __device__ __forceinline__ __half2 lo(__half2 v) { return __low2half2(v); }
    static __device__ __forceinline__ __half2 hi(__half2 v) { return __high2half2(v); }
    // s1 += v, s2 += v * v with the fp16 pair promoted inside the instruction (FHADD / FHFMA): same values as converting first,
    // without the two conversion instructions per pair and without the double-issue packed-fp32 forms
    static __device__ __forceinline__ void accumulate(__half2 v, float2& s1, float2& s2) {
        const unsigned short l = __half_as_ushort(__low2half(v)), h = __half_as_ushort(__high2half(v));
        asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(s1.x) : "h"(l));
        asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(s1.y) : "h"(h));
        asm("fma.rn.f32.f16 %0, %1, %1, %0;" : "+f"(s2.x) : "h"(l));
        asm("fma.rn.f32.f16 %0, %1, %1, %0;" : "+f"(s2.y) : "h"(h));
    }
    static __device__ __forceinline__ void start(__half2 v, float2& s1, float2& s2) {
        s1 = __half22float2(v);
        const unsigned short l = __half_as_ushort(__low2half(v)), h = __half_as_ushort(__high2half(v));
        asm("fma.rn.f32.f16 %0, %1, %1, %2;" : "=f"(s2.x) : "h"(l), "f"(0.f));
        asm("fma.rn.f32.f16 %0, %1, %1, %2;" : "=f"(s2.y) : "h"(h), "f"(0.f));
    }
};
template <> struct Pack2<__nv_bfloat16> {
    typedef __nv_bfloat162 type;
    static __device__ __forceinline__ __nv_bfloat162 splat(float w) { return __float2bfloat162_rn(w); }
    static __device__ __forceinline__ float2 to_f2(__nv_bfloat162 v) { return __bfloat1622float2(v); }
    static __device__ __forceinline__ __nv_bfloat162 from_f2(float2 v) { return __float22bfloat162_rn(v); }
    static __device__ __forceinline__ __nv_bfloat162 lo(__nv_bfloat162 v) { return __low2bfloat162(v); }
    static __device__ __forceinline__ __nv_bfloat162 hi(__nv_bfloat162 v) { return __high2bfloat162(v); }
    static __device__ __forceinline__ void accumulate(__nv_bfloat162 v, float2& s1, float2& s2) {
        const unsigned short l = __bfloat16_as_ushort(__low2bfloat16(v)), h = __bfloat16_as_ushort(__high2bfloat16(v));
        asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(s1.x) : "h"(l));
        asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(s1.y) : "h"(h));
        asm("fma.rn.f32.bf16 %0, %1, %1, %0;" : "+f"(s2.x) : "h"(l));
        asm("fma.rn.f32.bf16 %0, %1, %1, %0;" : "+f"(s2.y) : "h"(h));
    }
    static __device__ __forceinline__ void start(__nv_bfloat162 v, float2& s1, float2& s2) {
        s1 = __bfloat1622float2(v);
        s2 = __fmul2_rn(s1, s1);
    }
};

template <typename T, int CPT, int NS, int MINB>   // NS = compile-time bound on the source count (register arrays are sized by it)
__global__ void __launch_bounds__(128, MINB)
warp_var_fwd_fast_kernel(const T* __restrict__ ref, SrcPtrs srcs, int nsrc, const float* __restrict__ rt,
                         const float* __restrict__ depth, int per_pixel, T* __restrict__ var, int B, int CB, int D, int H,
                         int W, int dper, int align_corners, int ref_sq_in_sum) {
    typedef typename Pack2<T>::type T2;
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int nchunk = (D + dper - 1) / dper;
    const int CG = CB / CPT;
    int y = blockIdx.y;
    const int dc = y % nchunk; y /= nchunk;
    const int cg = y % CG;
    const int b = y / CG;
    const float fx = (float)(p % W), fy = (float)(p / W);
    const int64_t plane = (int64_t)HW * 8;                       // elements per channel block of a map
    const int64_t map_off = ((int64_t)b * CB + (int64_t)cg * CPT) * plane;

    float2 r[CPT][4];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(ref + map_off + c * plane + (int64_t)p * 8));
        const T2* h = reinterpret_cast<const T2*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) r[c][j] = Pack2<T>::to_f2(h[j]);
    }
    float ray[NS][3], tr[NS][3];
#pragma unroll
    for (int s = 0; s < NS; ++s)
        if (s < nsrc) {
            const float* m = rt + ((int64_t)s * B + b) * 12;
            pixel_ray(m, fx, fy, ray[s]);
            tr[s][0] = __ldg(m + 9); tr[s][1] = __ldg(m + 10); tr[s][2] = __ldg(m + 11);
        }
    // ix = u * sx + ox : align_corners ? u : u * W/(W-1) - 0.5
    const float sx = align_corners ? 1.f : (float)W / (float)(W - 1), sy = align_corners ? 1.f : (float)H / (float)(H - 1);
    const float oxy = align_corners ? 0.f : -0.5f;
    const float fW = (float)W, fH = (float)H;
    const float inv_n = 1.f / (float)(nsrc + 1);
    const float2 inv_n2 = make_float2(inv_n, inv_n);

    const int d_end = min(D, (dc + 1) * dper);
    for (int d = dc * dper; d < d_end; ++d) {
        const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
        float2 s1[CPT][4], s2[CPT][4];
#pragma unroll
        for (int c = 0; c < CPT; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s2[c][j] = __fmul2_rn(r[c][j], r[c][j]); s1[c][j] = ref_sq_in_sum ? s2[c][j] : r[c][j]; }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (s < nsrc) {
                const float pz = ray[s][2] * dv + tr[s][2];
                float iz;                                        // pz == 0 -> inf -> NaN/inf coordinates -> every tap rejected
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(pz));
                // coordinates clamped to [-1, W] x [-1, H]: outside that range every tap is out of the map anyway, and the
                // clamp (fmaxf drops a NaN) keeps the magic-number floor below exact
                const float ix = fminf(fmaxf((ray[s][0] * dv + tr[s][0]) * iz * sx + oxy, -1.f), fW);
                const float iy = fminf(fmaxf((ray[s][1] * dv + tr[s][1]) * iz * sy + oxy, -1.f), fH);
                // floor and float->int without the conversion (XU) pipe: adding 1.5 * 2^23 with round-down leaves floor(ix)
                // in the low mantissa bits
                const float kMagic = 12582912.f;
                const float tx = __fadd_rd(ix, kMagic), ty = __fadd_rd(iy, kMagic);
                const int xi = __float_as_int(tx) - 0x4B400000, yi = __float_as_int(ty) - 0x4B400000;
                const float wx1 = ix - (tx - kMagic), wy1 = iy - (ty - kMagic);     // exact
                const float wx0 = 1.f - wx1, wy0 = 1.f - wy1;                       // = (x0 + 1) - ix, correctly rounded
                // taps outside the map get weight exactly 0 and read a clamped (in-map) pixel, so there is no divergent load
                const float mx0 = (unsigned)xi < (unsigned)W ? wx0 : 0.f, mx1 = (unsigned)(xi + 1) < (unsigned)W ? wx1 : 0.f;
                const float my0 = (unsigned)yi < (unsigned)H ? wy0 : 0.f, my1 = (unsigned)(yi + 1) < (unsigned)H ? wy1 : 0.f;
                const int xa = min(max(xi, 0), W - 1), xb = min(xi + 1, W - 1), ya = min(max(yi, 0), H - 1), yb = min(yi + 1, H - 1);
                const int o0 = ya * W + xa, o1 = ya * W + xb, o2 = yb * W + xa, o3 = yb * W + xb;
                const T2 w0 = Pack2<T>::splat(mx0 * my0), w1 = Pack2<T>::splat(mx1 * my0);
                const T2 w2 = Pack2<T>::splat(mx0 * my1), w3 = Pack2<T>::splat(mx1 * my1);
                const T* base = reinterpret_cast<const T*>(srcs.p[s]) + map_off;
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const T* m = base + c * plane;
                    const uint4 a = __ldg(reinterpret_cast<const uint4*>(m + (int64_t)o0 * 8));
                    const uint4 bq = __ldg(reinterpret_cast<const uint4*>(m + (int64_t)o1 * 8));
                    const uint4 cq = __ldg(reinterpret_cast<const uint4*>(m + (int64_t)o2 * 8));
                    const uint4 dq = __ldg(reinterpret_cast<const uint4*>(m + (int64_t)o3 * 8));
                    const T2* ha = reinterpret_cast<const T2*>(&a);
                    const T2* hb = reinterpret_cast<const T2*>(&bq);
                    const T2* hc = reinterpret_cast<const T2*>(&cq);
                    const T2* hd = reinterpret_cast<const T2*>(&dq);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        T2 v = __hmul2(ha[j], w0);
                        v = __hfma2(hb[j], w1, v);
                        v = __hfma2(hc[j], w2, v);
                        v = __hfma2(hd[j], w3, v);
                        const float2 f = Pack2<T>::to_f2(v);
                        s1[c][j] = __fadd2_rn(s1[c][j], f);
                        s2[c][j] = __ffma2_rn(f, f, s2[c][j]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            uint4 out;
            T2* ho = reinterpret_cast<T2*>(&out);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 m = __fmul2_rn(s1[c][j], inv_n2);
                const float2 q = __fmul2_rn(s2[c][j], inv_n2);
                ho[j] = Pack2<T>::from_f2(__ffma2_rn(make_float2(-m.x, -m.y), m, q));
            }
            *reinterpret_cast<uint4*>(var + ((((int64_t)b * CB + cg * CPT + c) * D + d) * HW + p) * 8) = out;
        }
    }
}

// ------------------------------------------------------------------------------------------------ fused forward, 16-bit, zero-bordered maps
// The production kernel.  Maps are "C8P": [M][C/8][H + 3][W + 2][8] with pixel (y, x) at row y + 1, column x + 1 and zeros
// around it (mvs_pack_c8_padded).  With sampling coordinates clamped to the border, grid_sample's zero padding needs no tap
// validity tests, no index clamps and no per-tap offsets: the four taps are base, +16 B, +row, +row + 16 B, and every tap
// outside the map reads a stored zero.  A thread owns ALL channels of its pixel (CPT channel blocks), so the homography,
// floor and weights are computed once per (voxel, source) instead of once per 16 channels; that halves the instruction count
// of this issue-bound kernel (SASS: 147 -> ~95 instructions per source and 16 channels).
template <typename T, int CPT, int NS, int MINB, bool REFSQ>
__global__ void __launch_bounds__(128, MINB)
warp_var_fwd_pad_kernel(const T* __restrict__ ref, SrcPtrs srcs, int nsrc, const float* __restrict__ rt,
                        const float* __restrict__ depth, int per_pixel, T* __restrict__ var, int B, int CB, int D, int H,
                        int W, int dper, int align_corners, int dzl) {
    typedef typename Pack2<T>::type T2;
    const int HW = H * W;
    // A warp covers (32 >> dzl) consecutive pixels x (1 << dzl) consecutive depth planes: neighbouring planes sample
    // neighbouring source pixels (the sweep moves a fraction of a pixel per plane), so the lanes of one gather instruction
    // fall into 2-3 cache lines instead of 5, and the kernel is bound by exactly those L1 wavefronts.  Stores stay whole
    // lines (8 pixels x 16 B = 128 B per plane at dzl = 2).
    const int lane = threadIdx.x & 31, pxw = 32 >> dzl;
    const int p = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * pxw + (lane & (pxw - 1));
    const int dz = lane >> (5 - dzl), dstep = 1 << dzl;
    if (p >= HW) return;
    const int nchunk = (D + dper - 1) / dper;
    const int CG = CB / CPT;
    int y = blockIdx.y;
    const int dc = y % nchunk; y /= nchunk;
    const int cg = y % CG;
    const int b = y / CG;
    const int py = p / W, px = p - py * W;
    const int pitch = W + 2;
    const uint32_t row_b = (uint32_t)pitch * 16u;                 // bytes per map row
    const uint32_t plane_b = (uint32_t)(H + 3) * row_b;           // bytes per channel block of a map
    const int64_t map_b = ((int64_t)b * CB + (int64_t)cg * CPT) * plane_b;

    uint4 rq[CPT];                                                // reference feature, kept packed
    {
        const char* rp = reinterpret_cast<const char*>(ref) + map_b + (uint32_t)((py + 1) * pitch + px + 1) * 16u;
#pragma unroll
        for (int c = 0; c < CPT; ++c) rq[c] = __ldg(reinterpret_cast<const uint4*>(rp + (size_t)c * plane_b));
    }
    // ix = u * sx + ox with u = (ray0 d + t0) / (ray2 d + t2): align_corners ? u : u * W/(W-1) - 0.5, plus 1 for the border;
    // the scale is folded into the ray and the translation
    const float sx = align_corners ? 1.f : (float)W / (float)(W - 1), sy = align_corners ? 1.f : (float)H / (float)(H - 1);
    const float oxy = (align_corners ? 0.f : -0.5f) + 1.f;
    float rx[NS], ry[NS], rz[NS], tx[NS], ty[NS], tz[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s)
        if (s < nsrc) {
            const float* m = rt + ((int64_t)s * B + b) * 12;
            float ray[3];
            pixel_ray(m, (float)px, (float)py, ray);
            rx[s] = ray[0] * sx; ry[s] = ray[1] * sy; rz[s] = ray[2];
            tx[s] = __ldg(m + 9) * sx; ty[s] = __ldg(m + 10) * sy; tz[s] = __ldg(m + 11);
        }
    const float xmax = (float)(W + 1), ymax = (float)(H + 1);
    const float inv_n = 1.f / (float)(nsrc + 1);
    const float2 inv_n2 = make_float2(inv_n, inv_n);

    const int d_end = min(D, (dc + 1) * dper);
    for (int d = dc * dper + dz; d < d_end; d += dstep) {
        const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
        // running sums over the SOURCES only (source 0 initialises them); the reference feature joins in the epilogue
        float2 s1[CPT][4], s2[CPT][4];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (s < nsrc) {
                const float pz = fmaf(rz[s], dv, tz[s]);
                float iz;                                        // pz == 0 -> inf -> NaN / inf coordinates -> clamped to the zero border
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(pz));
                // clamp to the border ring (fmaxf drops a NaN): beyond it every tap is outside the map and reads zeros anyway
                const float ix = fminf(fmaxf(fmaf(fmaf(rx[s], dv, tx[s]), iz, oxy), 0.f), xmax);
                const float iy = fminf(fmaxf(fmaf(fmaf(ry[s], dv, ty[s]), iz, oxy), 0.f), ymax);
                // floor + float->int without the conversion pipe: adding 1.5 * 2^23 rounding down leaves floor() in the mantissa
                const float kMagic = 12582912.f;
                const float fxm = __fadd_rd(ix, kMagic), fym = __fadd_rd(iy, kMagic);
                const int xi = __float_as_int(fxm) - 0x4B400000, yi = __float_as_int(fym) - 0x4B400000;
                const float wx = ix - (fxm - kMagic), wy = iy - (fym - kMagic);      // exact fractional parts
                const float w11 = wx * wy, w10 = wx - w11, w01 = wy - w11, w00 = (1.f - wx) - w01;
                const T2 wa = Pack2<T>::from_f2(make_float2(w00, w10)), wb = Pack2<T>::from_f2(make_float2(w01, w11));
                const T2 w0 = Pack2<T>::lo(wa), w1 = Pack2<T>::hi(wa), w2 = Pack2<T>::lo(wb), w3 = Pack2<T>::hi(wb);
                const char* base = reinterpret_cast<const char*>(srcs.p[s]) + map_b + (uint32_t)(yi * pitch + xi) * 16u;
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const char* m = base + (size_t)c * plane_b;
                    const uint4 a = __ldg(reinterpret_cast<const uint4*>(m));
                    const uint4 bq = __ldg(reinterpret_cast<const uint4*>(m + 16));
                    const uint4 cq = __ldg(reinterpret_cast<const uint4*>(m + row_b));
                    const uint4 dq = __ldg(reinterpret_cast<const uint4*>(m + row_b + 16));
                    const T2* ha = reinterpret_cast<const T2*>(&a);
                    const T2* hb = reinterpret_cast<const T2*>(&bq);
                    const T2* hc = reinterpret_cast<const T2*>(&cq);
                    const T2* hd = reinterpret_cast<const T2*>(&dq);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        T2 v = __hmul2(ha[j], w0);
                        v = __hfma2(hb[j], w1, v);
                        v = __hfma2(hc[j], w2, v);
                        v = __hfma2(hd[j], w3, v);
                        if (s == 0) Pack2<T>::start(v, s1[c][j], s2[c][j]);
                        else Pack2<T>::accumulate(v, s1[c][j], s2[c][j]);
                    }
                }
            }
        }
        T* vp = var + ((((int64_t)b * CB + cg * CPT) * D + d) * HW + p) * 8;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            uint4 out;
            T2* ho = reinterpret_cast<T2*>(&out);
#pragma unroll
            const T2* h = reinterpret_cast<const T2*>(&rq[c]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // var = (S2 + r^2) / N - ((S1 + r) / N)^2 = ((S2 + r^2) - (S1 + r)^2 / N) / N     (S1 + r^2 when REFSQ, hazard H2)
                const float2 r = Pack2<T>::to_f2(h[j]);
                float2 t, u;
                if (REFSQ) { const float2 r2 = __fmul2_rn(r, r); t = __fadd2_rn(s1[c][j], r2); u = __fadd2_rn(s2[c][j], r2); }
                else { t = __fadd2_rn(s1[c][j], r); u = __ffma2_rn(r, r, s2[c][j]); }
                const float2 v = __ffma2_rn(__fmul2_rn(t, t), make_float2(-inv_n, -inv_n), u);
                ho[j] = Pack2<T>::from_f2(__fmul2_rn(v, inv_n2));
            }
            *reinterpret_cast<uint4*>(vp + (int64_t)c * D * HW * 8) = out;
        }
    }
}

// ------------------------------------------------------------------------------------------------ fused forward, 16-bit, TMA-staged source tiles
// A block owns an 8 x 16 pixel tile of the reference frustum x a chunk of DC depth planes x all channels.  The pixels of every
// source map that the tile can touch over the chunk form a small window (the sweep moves a fraction of a pixel per plane, the
// homography is close to a translation over 16 pixels): its bounding box is found from the tile corners x the planes of the
// chunk, and when it fits a kBH x kBW box the window of ALL channel blocks is fetched by one TMA per source into shared
// memory ([cb][row][pixel][8], so a quarter-warp's gather is 128 contiguous bytes: conflict-free, exactly 4 wavefronts per
// instruction where the L1 path pays ~5 for line straddling, 32-bit addresses, no L2 misses).  A source whose window does not
// fit (wide baseline, plane through the camera, non-finite depth) is gathered from global memory exactly like
// warp_var_fwd_pad_kernel, per source and per block, so the result never depends on the staging decision.
// The maps are the zero-bordered C8P maps; TMA zero-fills whatever lies beyond them.
constexpr int kBH = 12, kBW = 24;          // staged window (pixels); 18 KB per source at 32 channels
constexpr int kTileH = 8, kTileW = 16;     // reference pixels per block

struct WvMaps { CUtensorMap m[MVS_MAX_SRC]; };

__device__ __forceinline__ uint32_t wv_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 wv_lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

template <typename T, int NS, bool REFSQ, int MINB>
__global__ void __launch_bounds__(256, MINB)
warp_var_fwd_tma_kernel(const __grid_constant__ WvMaps maps, const T* __restrict__ ref, SrcPtrs srcs, int nsrc,
                        const float* __restrict__ rt, const float* __restrict__ depth, T* __restrict__ var, int B, int CB, int D,
                        int H, int W, int DC, int align_corners) {
    typedef typename Pack2<T>::type T2;
    extern __shared__ __align__(128) uint8_t wv_smem[];
    __shared__ uint64_t bar;
    __shared__ int s_min[NS][2], s_max[NS][2], s_bad[NS];     // clamped sampling coordinates of the block (float bits; all >= 0)
    __shared__ int s_org[NS][2], s_staged[NS];

    const int HW = H * W;
    const int tiles_x = (W + kTileW - 1) / kTileW;
    const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
    const int nchunk = (D + DC - 1) / DC;
    const int b = blockIdx.y / nchunk, dc = blockIdx.y - b * nchunk;
    const int d_begin = dc * DC, d_end = min(D, d_begin + DC);
    const int q = threadIdx.x & 127, cp = threadIdx.x >> 7;          // pixel of the tile, channel-block pair
    const int py = tile_y * kTileH + (q >> 4), px = tile_x * kTileW + (q & 15);
    const bool active = py < H && px < W;
    const int pitch = W + 2;
    const uint32_t row_b = (uint32_t)pitch * 16u, plane_b = (uint32_t)(H + 3) * row_b;
    const uint32_t win_plane = kBH * kBW * 16u, win_src = (uint32_t)CB * win_plane;     // bytes per channel block / per source in smem

    const float sx = align_corners ? 1.f : (float)W / (float)(W - 1), sy = align_corners ? 1.f : (float)H / (float)(H - 1);
    const float oxy = (align_corners ? 0.f : -0.5f) + 1.f;
    const float xmax = (float)(W + 1), ymax = (float)(H + 1);

    if (threadIdx.x < NS) {
        s_min[threadIdx.x][0] = s_min[threadIdx.x][1] = 0x7f800000; s_max[threadIdx.x][0] = s_max[threadIdx.x][1] = 0; s_bad[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(wv_smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // ---- window of every source: the 4 tile corners x every plane of the chunk, with exactly the arithmetic of the gather below
    {
        const int nd = d_end - d_begin;
        const int x0 = tile_x * kTileW, x1 = min(x0 + kTileW - 1, W - 1), y0 = tile_y * kTileH, y1 = min(y0 + kTileH - 1, H - 1);
        for (int e = threadIdx.x; e < nsrc * 4 * nd; e += blockDim.x) {
            const int s = e / (4 * nd), r = e - s * 4 * nd, corner = r / nd, d = d_begin + (r - corner * nd);
            const float* m = rt + ((int64_t)s * B + b) * 12;
            float ray[3];
            pixel_ray(m, (float)((corner & 1) ? x1 : x0), (float)((corner & 2) ? y1 : y0), ray);
            const float dv = __ldg(depth + (int64_t)b * D + d);
            const float pz = fmaf(ray[2], dv, __ldg(m + 11));
            float iz;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(pz));
            const float ix = fminf(fmaxf(fmaf(fmaf(ray[0] * sx, dv, __ldg(m + 9) * sx), iz, oxy), 0.f), xmax);
            const float iy = fminf(fmaxf(fmaf(fmaf(ray[1] * sy, dv, __ldg(m + 10) * sy), iz, oxy), 0.f), ymax);
            const int bad = (!(pz > 0.f) || !(fabsf(dv) < 3.0e38f)) ? 1 : 0;       // the map is not monotone across pz <= 0: no window
            int lox = __float_as_int(ix), hix = lox, loy = __float_as_int(iy), hiy = loy;   // non-negative floats order like their bits
            if ((4 * nd) % 32 == 0 && e - (int)(threadIdx.x & 31) + 31 < nsrc * 4 * nd) {
                // the whole warp evaluates one source: reduce in the warp (redux), one shared-memory atomic per warp and quantity
                lox = __reduce_min_sync(0xffffffffu, lox); hix = __reduce_max_sync(0xffffffffu, hix);
                loy = __reduce_min_sync(0xffffffffu, loy); hiy = __reduce_max_sync(0xffffffffu, hiy);
                const int anybad = __reduce_max_sync(0xffffffffu, bad);
                if ((threadIdx.x & 31) == 0) {
                    if (anybad) atomicOr(&s_bad[s], 1);
                    atomicMin(&s_min[s][0], lox); atomicMax(&s_max[s][0], hix); atomicMin(&s_min[s][1], loy); atomicMax(&s_max[s][1], hiy);
                }
            } else {
                if (bad) atomicOr(&s_bad[s], 1);
                atomicMin(&s_min[s][0], lox); atomicMax(&s_max[s][0], hix); atomicMin(&s_min[s][1], loy); atomicMax(&s_max[s][1], hiy);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < nsrc) {
        const int s = threadIdx.x;
        // one pixel of slack on the low side and two on the high side (the +1 tap and the rounding of interior pixels, whose
        // coordinates are fused differently from the corners' by at most an ulp)
        // (slack grows with the coordinate magnitude: the interior pixels' fused arithmetic and rcp.approx agree with the corners'
        // to a few ulps of the coordinate, 2^-22 relative -- 0.01 px covers maps up to ~4000 px, beyond that the term takes over)
        const float slack = 0.01f + 4.0e-6f * fmaxf(xmax, ymax);
        const int bx = max((int)floorf(__int_as_float(s_min[s][0]) - slack), 0), by = max((int)floorf(__int_as_float(s_min[s][1]) - slack), 0);
        const int ex = (int)floorf(__int_as_float(s_max[s][0]) + slack) + 1, ey = (int)floorf(__int_as_float(s_max[s][1]) + slack) + 1;
        s_org[s][0] = bx; s_org[s][1] = by;
        s_staged[s] = (!s_bad[s] && ex - bx < kBW && ey - by < kBH) ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t bytes = 0;
        for (int s = 0; s < nsrc; ++s) bytes += s_staged[s] ? win_src : 0u;
        if (bytes) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wv_smem_u32(&bar)), "r"(bytes) : "memory");
            for (int s = 0; s < nsrc; ++s)
                if (s_staged[s])
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                 ::"r"(wv_smem_u32(wv_smem + (size_t)s * win_src)), "l"(&maps.m[s]), "r"(wv_smem_u32(&bar)),
                                   "r"(s_org[s][0] * 8), "r"(s_org[s][1]), "r"(0), "r"(b) : "memory");
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(wv_smem_u32(&bar)) : "memory");
        }
    }

    // ---- per-thread setup while the windows land
    const int pyc = min(py, H - 1), pxc = min(px, W - 1);            // inactive threads of a ragged tile compute on a clamped pixel, store nothing
    const int64_t map_b = ((int64_t)b * CB + (int64_t)cp * 2) * plane_b;
    uint4 rq[2];
    {
        const char* rp = reinterpret_cast<const char*>(ref) + map_b + (uint32_t)((pyc + 1) * pitch + pxc + 1) * 16u;
#pragma unroll
        for (int c = 0; c < 2; ++c) rq[c] = __ldg(reinterpret_cast<const uint4*>(rp + (size_t)c * plane_b));
    }
    float rx[NS], ry[NS], rz[NS], tx[NS], ty[NS], tz[NS];
    uint32_t sbase[NS];                                                // smem address of (cb pair, window origin) minus the origin offset
    bool stg[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s)
        if (s < nsrc) {
            const float* m = rt + ((int64_t)s * B + b) * 12;
            float ray[3];
            pixel_ray(m, (float)pxc, (float)pyc, ray);
            rx[s] = ray[0] * sx; ry[s] = ray[1] * sy; rz[s] = ray[2];
            tx[s] = __ldg(m + 9) * sx; ty[s] = __ldg(m + 10) * sy; tz[s] = __ldg(m + 11);
            stg[s] = s_staged[s] != 0;
            sbase[s] = wv_smem_u32(wv_smem) + (uint32_t)s * win_src + (uint32_t)cp * 2u * win_plane - (uint32_t)(s_org[s][1] * kBW + s_org[s][0]) * 16u;
        }
    const float inv_n = 1.f / (float)(nsrc + 1);
    const float2 inv_n2 = make_float2(inv_n, inv_n);
    {   // wait for the windows (phase 0)
        asm volatile("{\n\t.reg .pred p;\n\tWV_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra WV_DONE_%=;\n\tbra WV_WAIT_%=;\n\tWV_DONE_%=:\n\t}"
                     ::"r"(wv_smem_u32(&bar)) : "memory");
    }

    const int p = pyc * W + pxc;
    for (int d = d_begin; d < d_end; ++d) {
        const float dv = __ldg(depth + (int64_t)b * D + d);
        float2 s1[2][4], s2[2][4];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (s < nsrc) {
                const float pz = fmaf(rz[s], dv, tz[s]);
                float iz;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(pz));
                const float ix = fminf(fmaxf(fmaf(fmaf(rx[s], dv, tx[s]), iz, oxy), 0.f), xmax);
                const float iy = fminf(fmaxf(fmaf(fmaf(ry[s], dv, ty[s]), iz, oxy), 0.f), ymax);
                const float kMagic = 12582912.f;
                const float fxm = __fadd_rd(ix, kMagic), fym = __fadd_rd(iy, kMagic);
                const int xi = __float_as_int(fxm) - 0x4B400000, yi = __float_as_int(fym) - 0x4B400000;
                const float wx = ix - (fxm - kMagic), wy = iy - (fym - kMagic);
                const float w11 = wx * wy, w10 = wx - w11, w01 = wy - w11, w00 = (1.f - wx) - w01;
                const T2 wa = Pack2<T>::from_f2(make_float2(w00, w10)), wb = Pack2<T>::from_f2(make_float2(w01, w11));
                const T2 w0 = Pack2<T>::lo(wa), w1 = Pack2<T>::hi(wa), w2 = Pack2<T>::lo(wb), w3 = Pack2<T>::hi(wb);
                uint4 tap[2][4];
                if (stg[s]) {
                    const uint32_t a = sbase[s] + (uint32_t)(yi * kBW + xi) * 16u;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        tap[c][0] = wv_lds128(a + c * win_plane);
                        tap[c][1] = wv_lds128(a + c * win_plane + 16u);
                        tap[c][2] = wv_lds128(a + c * win_plane + kBW * 16u);
                        tap[c][3] = wv_lds128(a + c * win_plane + kBW * 16u + 16u);
                    }
                } else {
                    const char* base = reinterpret_cast<const char*>(srcs.p[s]) + map_b + (uint32_t)(yi * pitch + xi) * 16u;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const char* m = base + (size_t)c * plane_b;
                        tap[c][0] = __ldg(reinterpret_cast<const uint4*>(m));
                        tap[c][1] = __ldg(reinterpret_cast<const uint4*>(m + 16));
                        tap[c][2] = __ldg(reinterpret_cast<const uint4*>(m + row_b));
                        tap[c][3] = __ldg(reinterpret_cast<const uint4*>(m + row_b + 16));
                    }
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const T2* ha = reinterpret_cast<const T2*>(&tap[c][0]);
                    const T2* hb = reinterpret_cast<const T2*>(&tap[c][1]);
                    const T2* hc = reinterpret_cast<const T2*>(&tap[c][2]);
                    const T2* hd = reinterpret_cast<const T2*>(&tap[c][3]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        T2 v = __hmul2(ha[j], w0);
                        v = __hfma2(hb[j], w1, v);
                        v = __hfma2(hc[j], w2, v);
                        v = __hfma2(hd[j], w3, v);
                        if (s == 0) Pack2<T>::start(v, s1[c][j], s2[c][j]);
                        else Pack2<T>::accumulate(v, s1[c][j], s2[c][j]);
                    }
                }
            }
        }
        T* vp = var + ((((int64_t)b * CB + cp * 2) * D + d) * HW + p) * 8;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint4 out;
            T2* ho = reinterpret_cast<T2*>(&out);
            const T2* h = reinterpret_cast<const T2*>(&rq[c]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 r = Pack2<T>::to_f2(h[j]);
                float2 t, u;
                if (REFSQ) { const float2 r2 = __fmul2_rn(r, r); t = __fadd2_rn(s1[c][j], r2); u = __fadd2_rn(s2[c][j], r2); }
                else { t = __fadd2_rn(s1[c][j], r); u = __ffma2_rn(r, r, s2[c][j]); }
                const float2 v = __ffma2_rn(__fmul2_rn(t, t), make_float2(-inv_n, -inv_n), u);
                ho[j] = Pack2<T>::from_f2(__fmul2_rn(v, inv_n2));
            }
            if (active) *reinterpret_cast<uint4*>(vp + (int64_t)c * D * HW * 8) = out;
        }
    }
}

// ------------------------------------------------------------------------------------------------ fused backward, 16-bit storage
// Training path (zero-bordered C8P maps, 16-bit volumes).  Same gradient as warp_var_bwd_kernel; what differs is the scatter:
//   * a thread owns (pixel, 8 channels) and walks the planes of its chunk.  Consecutive planes sample the SAME 2x2 source cell
//     for several steps (the sweep moves a fraction of a pixel per plane), so the four tap contributions are accumulated in
//     registers while the cell is unchanged and flushed with 8 vector reductions (red.global.add.v4.f32) when it changes:
//     ~3x fewer reductions than per plane, and 4x fewer instructions than scalar atomics (32 of them per voxel and source in
//     warp_var_bwd_kernel).  Shared-memory privatisation was rejected: fp32 atomicAdd on shared memory is a CAS loop on sm_100
//     (ATOMS.CAST.SPIN, 2 cycles per lane), slower than REDG to L2 (1.3 cycles per lane for 4 floats).
//   * 64 accumulator registers hold two sources, so sources are processed in groups of two, each group re-gathering all sources
//     for S1 (the maps are L1 / L2 resident); taps are blended in fp32 with exact weights.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename T, int NS, bool REFSQ>
__global__ void __launch_bounds__(128, 2)
warp_var_bwd16_kernel(const T* __restrict__ gvar, const T* __restrict__ ref, SrcPtrs srcs, int nsrc, const float* __restrict__ rt,
                      const float* __restrict__ depth, int per_pixel, float* __restrict__ gref, GradPtrs gsrcs, int B, int CB, int D,
                      int H, int W, int dper, int align_corners) {
    const int HW = H * W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int nchunk = (D + dper - 1) / dper;
    int y = blockIdx.y;
    const int dc = y % nchunk; y /= nchunk;
    const int cb = y % CB;
    const int b = y / CB;
    const int py = p / W, px = p - py * W;
    const int pitch = W + 2;
    const uint32_t row_b = (uint32_t)pitch * 16u, plane_b = (uint32_t)(H + 3) * row_b;
    const int64_t map_b = ((int64_t)b * CB + cb) * plane_b;
    const int64_t grad_off = ((int64_t)b * CB + cb) * HW * 8;

    float r[8];
    V8<T>::load(reinterpret_cast<const T*>(reinterpret_cast<const char*>(ref) + map_b + (uint32_t)((py + 1) * pitch + px + 1) * 16u), r);
    const float sx = align_corners ? 1.f : (float)W / (float)(W - 1), sy = align_corners ? 1.f : (float)H / (float)(H - 1);
    const float oxy = (align_corners ? 0.f : -0.5f) + 1.f;
    float rx[NS], ry[NS], rz[NS], tx[NS], ty[NS], tz[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s)
        if (s < nsrc) {
            const float* m = rt + ((int64_t)s * B + b) * 12;
            float ray[3];
            pixel_ray(m, (float)px, (float)py, ray);
            rx[s] = ray[0] * sx; ry[s] = ray[1] * sy; rz[s] = ray[2];
            tx[s] = __ldg(m + 9) * sx; ty[s] = __ldg(m + 10) * sy; tz[s] = __ldg(m + 11);
        }
    const float xmax = (float)(W + 1), ymax = (float)(H + 1);
    const float inv_n = 1.f / (float)(nsrc + 1);
    float gr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) gr[k] = 0.f;
    const int d_begin = dc * dper, d_end = min(D, d_begin + dper);

    for (int g0 = 0; g0 < nsrc; g0 += 2) {
        int cell[2] = {-1, -1};                  // padded-map index (yi * pitch + xi) of the cell the accumulators belong to
        float acc[2][4][8];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[q][t][k] = 0.f;
        auto flush = [&](int q) {
            if (cell[q] < 0) return;
            float* gs = gsrcs.p[g0 + q] + grad_off;
            const int yi = cell[q] / pitch, xi = cell[q] - yi * pitch;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int yy = yi + (t >> 1) - 1, xx = xi + (t & 1) - 1;       // pixel of the un-bordered gradient map
                if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) {
                    float* o = gs + (int64_t)(yy * W + xx) * 8;
                    red_add_v4(o, acc[q][t][0], acc[q][t][1], acc[q][t][2], acc[q][t][3]);
                    red_add_v4(o + 4, acc[q][t][4], acc[q][t][5], acc[q][t][6], acc[q][t][7]);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[q][t][k] = 0.f;
            }
        };
        const bool on0 = gsrcs.p[g0] != nullptr, on1 = g0 + 1 < nsrc && gsrcs.p[g0 + 1] != nullptr;
        for (int d = d_begin; d < d_end; ++d) {
            const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
            float g[8];
            V8<T>::load(gvar + ((((int64_t)b * CB + cb) * D + d) * HW + p) * 8, g);
            float s1[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) s1[k] = REFSQ ? r[k] * r[k] : r[k];
            float ws[2][8], wt[2][4];
            int cid[2] = {-1, -1};
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (s < nsrc) {
                    const float pz = fmaf(rz[s], dv, tz[s]);
                    float iz;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(pz));
                    const float ix = fminf(fmaxf(fmaf(fmaf(rx[s], dv, tx[s]), iz, oxy), 0.f), xmax);
                    const float iy = fminf(fmaxf(fmaf(fmaf(ry[s], dv, ty[s]), iz, oxy), 0.f), ymax);
                    const float fxm = floorf(ix), fym = floorf(iy);
                    const int xi = (int)fxm, yi = (int)fym;
                    const float wx = ix - fxm, wy = iy - fym;
                    const float w11 = wx * wy, w10 = wx - w11, w01 = wy - w11, w00 = (1.f - wx) - w01;
                    const char* base = reinterpret_cast<const char*>(srcs.p[s]) + map_b + (uint32_t)(yi * pitch + xi) * 16u;
                    float a0[8], a1[8], a2[8], a3[8], v[8];
                    V8<T>::load(reinterpret_cast<const T*>(base), a0);
                    V8<T>::load(reinterpret_cast<const T*>(base + 16), a1);
                    V8<T>::load(reinterpret_cast<const T*>(base + row_b), a2);
                    V8<T>::load(reinterpret_cast<const T*>(base + row_b + 16), a3);
#pragma unroll
                    for (int k = 0; k < 8; ++k) { v[k] = fmaf(a3[k], w11, fmaf(a2[k], w01, fmaf(a1[k], w10, a0[k] * w00))); s1[k] += v[k]; }
                    if (s == g0 || s == g0 + 1) {
                        const int q = s - g0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) ws[q][k] = v[k];
                        wt[q][0] = w00; wt[q][1] = w10; wt[q][2] = w01; wt[q][3] = w11;
                        cid[q] = yi * pitch + xi;
                    }
                }
            }
            float m2[8];      // 2 S1 / N^2
#pragma unroll
            for (int k = 0; k < 8; ++k) m2[k] = 2.f * s1[k] * inv_n * inv_n;
            if (g0 == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float dr = REFSQ ? (2.f * r[k] * inv_n - m2[k] * 2.f * r[k]) : (2.f * r[k] * inv_n - m2[k]);
                    gr[k] += g[k] * dr;
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q == 0 ? on0 : on1) {
                    if (cid[q] != cell[q]) { flush(q); cell[q] = cid[q]; }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float c = g[k] * (2.f * ws[q][k] * inv_n - m2[k]);
#pragma unroll
                        for (int t = 0; t < 4; ++t) acc[q][t][k] = fmaf(wt[q][t], c, acc[q][t][k]);
                    }
                }
            }
        }
        if (on0) flush(0);
        if (on1) flush(1);
    }
    if (gref != nullptr) {
        float* o = gref + grad_off + (int64_t)p * 8;
        red_add_v4(o, gr[0], gr[1], gr[2], gr[3]);
        red_add_v4(o + 4, gr[4], gr[5], gr[6], gr[7]);
    }
}

// The same gradient with ONE THREAD PER (pixel, channel block, SOURCE): G = 2 / 4 / 8 adjacent lanes own the sources of a pixel.
// Each lane gathers and blends only its own source (the kernel above blends every source once per pair of sources, because 64
// accumulator registers hold two sources at most), the lanes of a pixel sum their warped values with log2(G) butterfly shuffles
// to get S1, and every lane keeps the run-length accumulators of its own source only (32 registers): ~1.8x fewer instructions
// per voxel, twice the resident warps.  Arithmetic per element is unchanged (fp32 blend with exact weights, fp32 coefficients).
template <typename T, int G, bool REFSQ>
__global__ void __launch_bounds__(128, 4)
warp_var_bwd16s_kernel(const T* __restrict__ gvar, const T* __restrict__ ref, SrcPtrs srcs, int nsrc, const float* __restrict__ rt,
                       const float* __restrict__ depth, int per_pixel, float* __restrict__ gref, GradPtrs gsrcs, int B, int CB, int D,
                       int H, int W, int dper, int align_corners) {
    const int HW = H * W;
    const int sl = threadIdx.x % G;                                    // source owned by this lane
    const int pr = blockIdx.x * (128 / G) + threadIdx.x / G;
    const bool active = pr < HW;
    const int p = active ? pr : HW - 1;                                // lanes past the map compute on a clamped pixel (shuffles stay converged)
    const int nchunk = (D + dper - 1) / dper;
    int y = blockIdx.y;
    const int dc = y % nchunk; y /= nchunk;
    const int cb = y % CB;
    const int b = y / CB;
    const int py = p / W, px = p - py * W;
    const int pitch = W + 2;
    const uint32_t row_b = (uint32_t)pitch * 16u, plane_b = (uint32_t)(H + 3) * row_b;
    const int64_t map_b = ((int64_t)b * CB + cb) * plane_b;
    const int64_t grad_off = ((int64_t)b * CB + cb) * HW * 8;
    const bool has_src = sl < nsrc;
    const int s = has_src ? sl : 0;

    float r[8];                                                         // reference features: lane 0 of the pixel carries them into S1
    V8<T>::load(reinterpret_cast<const T*>(reinterpret_cast<const char*>(ref) + map_b + (uint32_t)((py + 1) * pitch + px + 1) * 16u), r);
    const float sx = align_corners ? 1.f : (float)W / (float)(W - 1), sy = align_corners ? 1.f : (float)H / (float)(H - 1);
    const float oxy = (align_corners ? 0.f : -0.5f) + 1.f;
    float rx, ry, rz, tx, ty, tz;
    {
        const float* m = rt + ((int64_t)s * B + b) * 12;
        float ray[3];
        pixel_ray(m, (float)px, (float)py, ray);
        rx = ray[0] * sx; ry = ray[1] * sy; rz = ray[2];
        tx = __ldg(m + 9) * sx; ty = __ldg(m + 10) * sy; tz = __ldg(m + 11);
    }
    const char* smap = reinterpret_cast<const char*>(srcs.p[s]) + map_b;
    float* gs = gsrcs.p[s];
    const bool on = has_src && active && gs != nullptr;
    if (gs) gs += grad_off;
    const float xmax = (float)(W + 1), ymax = (float)(H + 1);
    const float inv_n = 1.f / (float)(nsrc + 1);
    float gr[8], acc[4][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { gr[k] = 0.f; acc[0][k] = acc[1][k] = acc[2][k] = acc[3][k] = 0.f; }
    int cell = -1;                                                      // padded-map index (yi * pitch + xi) the accumulators belong to
    auto flush = [&]() {
        if (cell < 0) return;
        const int yi = cell / pitch, xi = cell - yi * pitch;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int yy = yi + (t >> 1) - 1, xx = xi + (t & 1) - 1;   // pixel of the un-bordered gradient map
            if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) {
                float* o = gs + (int64_t)(yy * W + xx) * 8;
                red_add_v4(o, acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
                red_add_v4(o + 4, acc[t][4], acc[t][5], acc[t][6], acc[t][7]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
        }
    };
    const int d_begin = dc * dper, d_end = min(D, d_begin + dper);
    for (int d = d_begin; d < d_end; ++d) {
        const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
        float g[8];
        V8<T>::load(gvar + ((((int64_t)b * CB + cb) * D + d) * HW + p) * 8, g);
        const float pz = fmaf(rz, dv, tz);
        float iz;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(pz));
        const float ix = fminf(fmaxf(fmaf(fmaf(rx, dv, tx), iz, oxy), 0.f), xmax);
        const float iy = fminf(fmaxf(fmaf(fmaf(ry, dv, ty), iz, oxy), 0.f), ymax);
        const float fxm = floorf(ix), fym = floorf(iy);
        const int xi = (int)fxm, yi = (int)fym;
        const float wx = ix - fxm, wy = iy - fym;
        const float w11 = wx * wy, w10 = wx - w11, w01 = wy - w11, w00 = (1.f - wx) - w01;
        const char* base = smap + (uint32_t)(yi * pitch + xi) * 16u;
        float a0[8], a1[8], a2[8], a3[8], v[8], s1[8];
        V8<T>::load(reinterpret_cast<const T*>(base), a0);
        V8<T>::load(reinterpret_cast<const T*>(base + 16), a1);
        V8<T>::load(reinterpret_cast<const T*>(base + row_b), a2);
        V8<T>::load(reinterpret_cast<const T*>(base + row_b + 16), a3);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = fmaf(a3[k], w11, fmaf(a2[k], w01, fmaf(a1[k], w10, a0[k] * w00)));
            s1[k] = (has_src ? v[k] : 0.f) + (sl == 0 ? (REFSQ ? r[k] * r[k] : r[k]) : 0.f);
        }
#pragma unroll
        for (int o = 1; o < G; o <<= 1)
#pragma unroll
            for (int k = 0; k < 8; ++k) s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
        float m2[8];      // 2 S1 / N^2
#pragma unroll
        for (int k = 0; k < 8; ++k) m2[k] = 2.f * s1[k] * inv_n * inv_n;
        if (sl == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float dr = REFSQ ? (2.f * r[k] * inv_n - m2[k] * 2.f * r[k]) : (2.f * r[k] * inv_n - m2[k]);
                gr[k] += g[k] * dr;
            }
        }
        if (on) {
            const int cid = yi * pitch + xi;
            if (cid != cell) { flush(); cell = cid; }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float c = g[k] * (2.f * v[k] * inv_n - m2[k]);
                acc[0][k] = fmaf(w00, c, acc[0][k]); acc[1][k] = fmaf(w10, c, acc[1][k]);
                acc[2][k] = fmaf(w01, c, acc[2][k]); acc[3][k] = fmaf(w11, c, acc[3][k]);
            }
        }
    }
    if (on) flush();
    if (gref != nullptr && sl == 0 && active) {
        float* o = gref + grad_off + (int64_t)p * 8;
        red_add_v4(o, gr[0], gr[1], gr[2], gr[3]);
        red_add_v4(o + 4, gr[4], gr[5], gr[6], gr[7]);
    }
}

typedef CUresult (*WvEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WvEncodeTiledFn wv_encode_tiled() {
    static WvEncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<WvEncodeTiledFn>(p);
    }
    return fn;
}
#endif  // !MVS_CPU_EMU

static int depth_chunk(int D, int HW, int B, int CB) {
    // enough (pixel-tile x chunk) blocks for >= ~4 waves of 148 SMs x 8 resident blocks, chunks of >= 8 planes
    int dper = D;
    const int64_t tiles = (int64_t)mvs_cdiv(HW, 128) * B * CB;
    while (dper > 8 && tiles * ((D + dper - 1) / dper) < 148 * 8 * 4) dper = (dper + 1) / 2;
    return dper;
}

static int check_warp_var(const void* ref, const void* const* srcs, int nsrc, const float* rt, const float* depth,
                          int B, int C, int D, int H, int W) {
    MVS_REQUIRE(ref && srcs && rt && depth, MVS_E_ARG, "warp_var: null pointer");
    MVS_REQUIRE(nsrc >= 1 && nsrc <= MVS_MAX_SRC, MVS_E_SHAPE, "warp_var: need 1 <= nsrc <= %d, got %d", MVS_MAX_SRC, nsrc);
    for (int s = 0; s < nsrc; ++s) MVS_REQUIRE(srcs[s], MVS_E_ARG, "warp_var: source %d is null", s);
    MVS_REQUIRE(B > 0 && D > 0 && H > 1 && W > 1, MVS_E_SHAPE, "warp_var: bad dims B=%d D=%d H=%d W=%d", B, D, H, W);
    MVS_REQUIRE(C > 0 && C % 8 == 0, MVS_E_SHAPE, "warp_var: C=%d must be a multiple of 8 (C8 layout)", C);
    MVS_REQUIRE((int64_t)B * (C / 8) * D <= 65535 * 8, MVS_E_SHAPE, "warp_var: B*C/8*D too large for the launch grid");
    return MVS_OK;
}

extern "C" int mvs_warp_var_fwd(const void* ref, const void* const* srcs, int nsrc, const float* rt, const float* depth,
                                int per_pixel, void* var, int B, int C, int D, int H, int W, int dtype_in, int dtype_out,
                                int align_corners, int ref_sq_in_sum, int pad, void* stream) {
    int rc = check_warp_var(ref, srcs, nsrc, rt, depth, B, C, D, H, W);
    if (rc) return rc;
    MVS_REQUIRE(var, MVS_E_ARG, "mvs_warp_var_fwd: null output");
    SrcPtrs sp;
    for (int s = 0; s < MVS_MAX_SRC; ++s) sp.p[s] = s < nsrc ? srcs[s] : nullptr;
    const int CB = C / 8, HW = H * W;
#ifndef MVS_CPU_EMU
    const int tma_knob = mvs_knob(MVS_KNOB_WARP_TMA, 1);    // test knob (mvs_set_knob): 0 = gather every source from global memory
    if (dtype_in == dtype_out && dtype_in != MVS_F32 && pad && !per_pixel && (CB == 2 || CB == 4) && tma_knob) {
        // 16-bit storage, zero-bordered maps, plane hypotheses shared by the pixels of an item: TMA-staged source windows
        const size_t smem = (size_t)nsrc * CB * kBH * kBW * 16;
        WvEncodeTiledFn enc = wv_encode_tiled();
        if (enc && smem <= 200 * 1024) {
            WvMaps maps;
            memset(&maps, 0, sizeof(maps));
            const CUtensorMapDataType dt = dtype_in == MVS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
            const cuuint64_t Wp = W + 2, Hp = H + 3;
            const cuuint64_t gdim[4] = {Wp * 8, Hp, (cuuint64_t)CB, (cuuint64_t)B};
            const cuuint64_t gstr[3] = {Wp * 16, Hp * Wp * 16, (cuuint64_t)CB * Hp * Wp * 16};
            const cuuint32_t box[4] = {kBW * 8, kBH, (cuuint32_t)CB, 1}, estr[4] = {1, 1, 1, 1};
            for (int s = 0; s < nsrc; ++s) {
                const CUresult cr = enc(&maps.m[s], dt, 4, const_cast<void*>(srcs[s]), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                MVS_REQUIRE(cr == CUDA_SUCCESS, MVS_E_LAUNCH, "mvs_warp_var_fwd: cuTensorMapEncodeTiled failed (%d)", (int)cr);
            }
            const int dc_knob = mvs_knob(MVS_KNOB_WARP_DC, 16);   // tuning knob: planes per block (window size grows with it)
            const int DC = (dc_knob >= 4 && dc_knob <= 64) ? dc_knob : 16;
            const int tiles = (int)(mvs_cdiv(H, kTileH) * mvs_cdiv(W, kTileW));
            MVS_REQUIRE((int64_t)B * mvs_cdiv(D, DC) <= 65535, MVS_E_SHAPE, "mvs_warp_var_fwd: B*D/16 too large for the launch grid");
            const dim3 gridt((unsigned)tiles, (unsigned)(B * mvs_cdiv(D, DC)));
            const unsigned nthr = 128u * (unsigned)(CB / 2);
#define MVS_WT_ARGS(T) maps, (const T*)ref, sp, nsrc, rt, depth, (T*)var, B, CB, D, H, W, DC, align_corners
#define MVS_WT_LAUNCH1(T, NS, RS, MB) do { cudaFuncSetAttribute(warp_var_fwd_tma_kernel<T, NS, RS, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                warp_var_fwd_tma_kernel<T, NS, RS, MB><<<gridt, nthr, smem, (cudaStream_t)stream>>>(MVS_WT_ARGS(T)); } while (0)
#define MVS_WT_LAUNCH(T, NS, MB) do { if (ref_sq_in_sum) MVS_WT_LAUNCH1(T, NS, true, MB); else MVS_WT_LAUNCH1(T, NS, false, MB); } while (0)
            const int tma_minb = mvs_knob(MVS_KNOB_WARP_TMA_MINB, 2);
#define MVS_WT_BY_NS(T) do { if (nsrc <= 2) MVS_WT_LAUNCH(T, 2, 3); else if (nsrc <= 4) { if (tma_minb == 3) MVS_WT_LAUNCH(T, 4, 3); else MVS_WT_LAUNCH(T, 4, 2); } \
                             else if (nsrc <= 6) MVS_WT_LAUNCH(T, 6, 1); else MVS_WT_LAUNCH(T, 8, 1); } while (0)
            if (dtype_in == MVS_F16) MVS_WT_BY_NS(__half); else MVS_WT_BY_NS(__nv_bfloat16);
#undef MVS_WT_BY_NS
#undef MVS_WT_LAUNCH
#undef MVS_WT_LAUNCH1
#undef MVS_WT_ARGS
            return MVS_CHECK_LAUNCH("mvs_warp_var_fwd");
        }
    }
    if (dtype_in == dtype_out && dtype_in != MVS_F32 && pad && CB % 2 == 0) {
        // 16-bit storage, zero-bordered maps: packed-math kernel, all channels of a pixel in one thread when C % 32 == 0
        MVS_REQUIRE((int64_t)(H + 3) * (W + 2) * 16 < (1ll << 31), MVS_E_SHAPE, "mvs_warp_var_fwd: maps too large");
        const int cpt_knob = mvs_knob(MVS_KNOB_WARP_CPT, 2);     // tuning knobs; measured at N=5, C=32, D=192, 128x160 x4 items:
        // CPT 2 / 4 blocks per SM 0.76 ms, CPT 4 / 3 blocks 0.91 ms (fewer instructions but too few warps to hide the gathers)
        const int minb4 = mvs_knob(MVS_KNOB_WARP_MINB, 3);
        const int dz_knob = mvs_knob(MVS_KNOB_WARP_DZ, 1);     // measured: 0 -> 0.795, 1 -> 0.759, 2 -> 0.768, 3 -> 0.813 ms
        const int cpt = (CB % 4 == 0 && cpt_knob == 4) ? 4 : 2;
        int dzl = dz_knob < 0 ? 0 : (dz_knob > 3 ? 3 : dz_knob);
        while (dzl > 0 && (1 << dzl) > D) --dzl;
        // a block = 4 warps = (128 >> dzl) pixels x (1 << dzl) planes per step; chunks of >= 8 steps, >= ~4 waves of blocks
        const int ptiles = (int)mvs_cdiv(HW, 128 >> dzl);
        int dperf = D;
        while (dperf > (8 << dzl) && (int64_t)ptiles * B * (CB / cpt) * ((D + dperf - 1) / dperf) < 148 * 8 * 4) dperf = (dperf + 1) / 2;
        dperf = (dperf + (1 << dzl) - 1) >> dzl << dzl;
        const dim3 gridf((unsigned)ptiles, (unsigned)(B * (CB / cpt) * ((D + dperf - 1) / dperf)));
#define MVS_WP_ARGS(T) (const T*)ref, sp, nsrc, rt, depth, per_pixel, (T*)var, B, CB, D, H, W, dperf, align_corners, dzl
#define MVS_WP_LAUNCH(T, CPT, NS, MB) do { if (ref_sq_in_sum) warp_var_fwd_pad_kernel<T, CPT, NS, MB, true><<<gridf, 128, 0, (cudaStream_t)stream>>>(MVS_WP_ARGS(T)); \
                                           else warp_var_fwd_pad_kernel<T, CPT, NS, MB, false><<<gridf, 128, 0, (cudaStream_t)stream>>>(MVS_WP_ARGS(T)); } while (0)
#define MVS_WP_BY_NS(T, CPT, MB) do { if (nsrc <= 2) MVS_WP_LAUNCH(T, CPT, 2, MB); else if (nsrc <= 4) MVS_WP_LAUNCH(T, CPT, 4, MB); \
                                      else if (nsrc <= 6) MVS_WP_LAUNCH(T, CPT, 6, MB); else MVS_WP_LAUNCH(T, CPT, 8, MB); } while (0)
#define MVS_WP_BY_CPT(T) do { if (cpt == 4 && minb4 == 2) MVS_WP_BY_NS(T, 4, 2); else if (cpt == 4) MVS_WP_BY_NS(T, 4, 3); else MVS_WP_BY_NS(T, 2, 4); } while (0)
        if (dtype_in == MVS_F16) MVS_WP_BY_CPT(__half); else MVS_WP_BY_CPT(__nv_bfloat16);
#undef MVS_WP_BY_CPT
#undef MVS_WP_BY_NS
#undef MVS_WP_LAUNCH
#undef MVS_WP_ARGS
        return MVS_CHECK_LAUNCH("mvs_warp_var_fwd");
    }
    if (dtype_in == dtype_out && dtype_in != MVS_F32 && CB % 2 == 0 && !pad) {
        // 16-bit storage, plain C8 maps: packed-math kernel, 2 channel blocks per thread, source-count bound in {2,4,6,8}
        const int dperf = depth_chunk(D, HW, B, CB / 2);
        const dim3 gridf(mvs_cdiv(HW, 128), (unsigned)(B * (CB / 2) * ((D + dperf - 1) / dperf)));
        const int mb4 = mvs_knob(MVS_KNOB_WARP_MINB, 4);   // tuning knob: 4 blocks/SM (122 registers, no spills) measured faster than 5 (96, spills)
#define MVS_WV_LAUNCH(T, NS, MB) warp_var_fwd_fast_kernel<T, 2, NS, MB><<<gridf, 128, 0, (cudaStream_t)stream>>>( \
            (const T*)ref, sp, nsrc, rt, depth, per_pixel, (T*)var, B, CB, D, H, W, dperf, align_corners, ref_sq_in_sum)
#define MVS_WV_BY_NS(T) do { if (nsrc <= 2) MVS_WV_LAUNCH(T, 2, 5); else if (nsrc <= 4) { if (mb4 == 4) MVS_WV_LAUNCH(T, 4, 4); else MVS_WV_LAUNCH(T, 4, 5); } \
                             else if (nsrc <= 6) MVS_WV_LAUNCH(T, 6, 3); else MVS_WV_LAUNCH(T, 8, 3); } while (0)
        if (dtype_in == MVS_F16) MVS_WV_BY_NS(__half); else MVS_WV_BY_NS(__nv_bfloat16);
#undef MVS_WV_BY_NS
#undef MVS_WV_LAUNCH
        return MVS_CHECK_LAUNCH("mvs_warp_var_fwd");
    }
#endif
    const int dper = depth_chunk(D, HW, B, CB);
    const dim3 grid(mvs_cdiv(HW, 128), (unsigned)(B * CB * ((D + dper - 1) / dper)));
    MVS_DISPATCH_DTYPE(dtype_in, TI, MVS_DISPATCH_DTYPE(dtype_out, TO,
        MVS_LAUNCH((warp_var_fwd_kernel<TI, TO>), grid, dim3(128), stream, (const TI*)ref, sp, nsrc, rt, depth, per_pixel,
                   (TO*)var, B, CB, D, H, W, dper, align_corners, ref_sq_in_sum, pad)));
    return MVS_CHECK_LAUNCH("mvs_warp_var_fwd");
}

extern "C" int mvs_warp_var_bwd(const void* grad_var, const void* ref, const void* const* srcs, int nsrc, const float* rt,
                                const float* depth, int per_pixel, float* grad_ref, float* const* grad_srcs, int B, int C,
                                int D, int H, int W, int dtype_in, int dtype_out, int align_corners, int ref_sq_in_sum,
                                int pad, void* stream) {
    int rc = check_warp_var(ref, srcs, nsrc, rt, depth, B, C, D, H, W);
    if (rc) return rc;
    MVS_REQUIRE(grad_var && grad_srcs, MVS_E_ARG, "mvs_warp_var_bwd: null pointer");
    SrcPtrs sp;
    GradPtrs gp;
    for (int s = 0; s < MVS_MAX_SRC; ++s) { sp.p[s] = s < nsrc ? srcs[s] : nullptr; gp.p[s] = s < nsrc ? grad_srcs[s] : nullptr; }
    const int CB = C / 8, HW = H * W;
    const int dper = depth_chunk(D, HW, B, CB);
    const dim3 grid(mvs_cdiv(HW, 128), (unsigned)(B * CB * ((D + dper - 1) / dper)));
#ifndef MVS_CPU_EMU
    if (dtype_in == dtype_out && dtype_in != MVS_F32 && pad) {
        // the training path: 16-bit storage, zero-bordered maps -> register-merged vector reductions
        MVS_REQUIRE((int64_t)(H + 3) * (W + 2) * 16 < (1ll << 31), MVS_E_SHAPE, "mvs_warp_var_bwd: maps too large");
#define MVS_WB_ARGS(T) (const T*)grad_var, (const T*)ref, sp, nsrc, rt, depth, per_pixel, grad_ref, gp, B, CB, D, H, W, dper, align_corners
        if (mvs_knob(MVS_KNOB_WARP_BWD_SPLIT, 1)) {
            // one thread per (pixel, channel block, source): G lanes per pixel
            const int G = nsrc <= 2 ? 2 : (nsrc <= 4 ? 4 : 8);
            const dim3 grids(mvs_cdiv(HW, 128 / G), grid.y);
#define MVS_WS_LAUNCH(T, GG) do { if (ref_sq_in_sum) warp_var_bwd16s_kernel<T, GG, true><<<grids, 128, 0, (cudaStream_t)stream>>>(MVS_WB_ARGS(T)); \
                                  else warp_var_bwd16s_kernel<T, GG, false><<<grids, 128, 0, (cudaStream_t)stream>>>(MVS_WB_ARGS(T)); } while (0)
#define MVS_WS_BY_G(T) do { if (G == 2) MVS_WS_LAUNCH(T, 2); else if (G == 4) MVS_WS_LAUNCH(T, 4); else MVS_WS_LAUNCH(T, 8); } while (0)
            if (dtype_in == MVS_F16) MVS_WS_BY_G(__half); else MVS_WS_BY_G(__nv_bfloat16);
#undef MVS_WS_BY_G
#undef MVS_WS_LAUNCH
            return MVS_CHECK_LAUNCH("mvs_warp_var_bwd");
        }
#define MVS_WB_LAUNCH(T, NS) do { if (ref_sq_in_sum) warp_var_bwd16_kernel<T, NS, true><<<grid, 128, 0, (cudaStream_t)stream>>>(MVS_WB_ARGS(T)); \
                                  else warp_var_bwd16_kernel<T, NS, false><<<grid, 128, 0, (cudaStream_t)stream>>>(MVS_WB_ARGS(T)); } while (0)
#define MVS_WB_BY_NS(T) do { if (nsrc <= 2) MVS_WB_LAUNCH(T, 2); else if (nsrc <= 4) MVS_WB_LAUNCH(T, 4); else if (nsrc <= 6) MVS_WB_LAUNCH(T, 6); \
                             else MVS_WB_LAUNCH(T, 8); } while (0)
        if (dtype_in == MVS_F16) MVS_WB_BY_NS(__half); else MVS_WB_BY_NS(__nv_bfloat16);
#undef MVS_WB_BY_NS
#undef MVS_WB_LAUNCH
#undef MVS_WB_ARGS
        return MVS_CHECK_LAUNCH("mvs_warp_var_bwd");
    }
#endif
    MVS_DISPATCH_DTYPE(dtype_in, TI, MVS_DISPATCH_DTYPE(dtype_out, TO,
        MVS_LAUNCH((warp_var_bwd_kernel<TI, TO>), grid, dim3(128), stream, (const TO*)grad_var, (const TI*)ref, sp, nsrc, rt,
                   depth, per_pixel, grad_ref, gp, B, CB, D, H, W, dper, align_corners, ref_sq_in_sum, pad)));
    return MVS_CHECK_LAUNCH("mvs_warp_var_bwd");
}

// ------------------------------------------------------------------------------------------------ stand-alone warp
// thread = (b, d, h, w); loops over channels of an NCHW fp32 map.  API-parity path, not the fused one.
__global__ void __launch_bounds__(256)
homo_warp_fwd_kernel(const float* __restrict__ src, const float* __restrict__ rt, const float* __restrict__ depth,
                     int per_pixel, float* __restrict__ out, int B, int C, int D, int H, int W, int align_corners) {
    const int HW = H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * D * HW) return;
    const int p = (int)(i % HW);
    const int d = (int)((i / HW) % D);
    const int b = (int)(i / ((int64_t)HW * D));
    float ray[3], ix, iy;
    pixel_ray(rt + (int64_t)b * 12, (float)(p % W), (float)(p / W), ray);
    const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
    source_coord(ray, rt + (int64_t)b * 12, dv, H, W, align_corners, ix, iy);
    Taps t;
    bilinear_taps(ix, iy, H, W, t);
    for (int c = 0; c < C; ++c) {
        const float* m = src + ((int64_t)b * C + c) * HW;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (t.off[k] >= 0) acc += __ldg(m + t.off[k]) * t.w[k];
        out[(((int64_t)b * C + c) * D + d) * HW + p] = acc;
    }
}

__global__ void __launch_bounds__(256)
homo_warp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ rt, const float* __restrict__ depth,
                     int per_pixel, float* __restrict__ gsrc, int B, int C, int D, int H, int W, int align_corners) {
    const int HW = H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * D * HW) return;
    const int p = (int)(i % HW);
    const int d = (int)((i / HW) % D);
    const int b = (int)(i / ((int64_t)HW * D));
    float ray[3], ix, iy;
    pixel_ray(rt + (int64_t)b * 12, (float)(p % W), (float)(p / W), ray);
    const float dv = per_pixel ? __ldg(depth + ((int64_t)b * D + d) * HW + p) : __ldg(depth + (int64_t)b * D + d);
    source_coord(ray, rt + (int64_t)b * 12, dv, H, W, align_corners, ix, iy);
    Taps t;
    bilinear_taps(ix, iy, H, W, t);
    for (int c = 0; c < C; ++c) {
        const float g = __ldg(gout + (((int64_t)b * C + c) * D + d) * HW + p);
        float* m = gsrc + ((int64_t)b * C + c) * HW;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (t.off[k] >= 0) atomicAdd(m + t.off[k], g * t.w[k]);
    }
}

extern "C" int mvs_homo_warp_fwd(const float* src, const float* rt, const float* depth, int per_pixel, float* out,
                                 int B, int C, int D, int H, int W, int align_corners, void* stream) {
    MVS_REQUIRE(src && rt && depth && out, MVS_E_ARG, "mvs_homo_warp_fwd: null pointer");
    MVS_REQUIRE(B > 0 && C > 0 && D > 0 && H > 1 && W > 1, MVS_E_SHAPE, "mvs_homo_warp_fwd: bad dims");
    const int64_t total = (int64_t)B * D * H * W;
    MVS_LAUNCH(homo_warp_fwd_kernel, dim3(mvs_cdiv(total, 256)), dim3(256), stream, src, rt, depth, per_pixel, out, B, C, D, H, W, align_corners);
    return MVS_CHECK_LAUNCH("mvs_homo_warp_fwd");
}

extern "C" int mvs_homo_warp_bwd(const float* grad_out, const float* rt, const float* depth, int per_pixel, float* grad_src,
                                 int B, int C, int D, int H, int W, int align_corners, void* stream) {
    MVS_REQUIRE(grad_out && rt && depth && grad_src, MVS_E_ARG, "mvs_homo_warp_bwd: null pointer");
    MVS_REQUIRE(B > 0 && C > 0 && D > 0 && H > 1 && W > 1, MVS_E_SHAPE, "mvs_homo_warp_bwd: bad dims");
    const int64_t total = (int64_t)B * D * H * W;
    MVS_LAUNCH(homo_warp_bwd_kernel, dim3(mvs_cdiv(total, 256)), dim3(256), stream, grad_out, rt, depth, per_pixel, grad_src, B, C, D, H, W, align_corners);
    return MVS_CHECK_LAUNCH("mvs_homo_warp_bwd");
}
