// core.cu — library identity, error reporting, C8 layout packing, projection algebra.
#include "mvs_rt.h"
#include "linalg.h"
#include <string.h>

#ifdef MVS_CPU_EMU
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#endif

static thread_local char g_err[512] = "";

int mvs_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int g_mvs_knobs[MVS_KNOB_COUNT] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
static const char* const kKnobNames[MVS_KNOB_COUNT] = {"warp_tma", "warp_dc", "warp_tma_minb", "warp_cpt", "warp_minb", "warp_dz",
                                                      "tc_kdfold", "tc_planes", "tc_nm", "tc_stages", "tc_nseg", "warp_bwd_split", "tc_kwfold_max"};
extern "C" int mvs_set_knob(const char* name, int value) {
    MVS_REQUIRE(name, MVS_E_ARG, "mvs_set_knob: null name");
    for (int i = 0; i < MVS_KNOB_COUNT; ++i)
        if (strcmp(name, kKnobNames[i]) == 0) { g_mvs_knobs[i] = value < 0 ? -1 : value; return MVS_OK; }
    return mvs_set_error(MVS_E_ARG, "mvs_set_knob: unknown knob '%s'", name);
}

extern "C" int mvs_version(void) { return MVS_B200_VERSION; }
extern "C" const char* mvs_last_error(void) { return g_err; }
extern "C" int mvs_is_emulation(void) {
#ifdef MVS_CPU_EMU
    return 1;
#else
    return 0;
#endif
}

// ---------------------------------------------------------------------------------------------- layout
// One thread moves one (b, channel-block, s) vector.  Reads are strided by S across the 8 channels but
// coalesced across threads (consecutive s); writes are one 16/32-byte vector per thread.
template <typename T>
__global__ void pack_c8_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int64_t S, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t s = i % S;
    const int64_t bc = i / S;  // b * (C/8) + cb
    const int cb = (int)(bc % (C / 8));
    const int64_t b = bc / (C / 8);
    const float* p = src + (b * C + (int64_t)cb * 8) * S + s;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(p + (int64_t)k * S);
    V8<T>::store(dst + i * 8, v);
}

template <typename T>
__global__ void unpack_c8_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int64_t S, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t s = i % S;
    const int64_t bc = i / S;
    const int cb = (int)(bc % (C / 8));
    const int64_t b = bc / (C / 8);
    float v[8];
    V8<T>::load(src + i * 8, v);
    float* p = dst + (b * C + (int64_t)cb * 8) * S + s;
#pragma unroll
    for (int k = 0; k < 8; ++k) p[(int64_t)k * S] = v[k];
}

// channels-last [B][S][C] (what the library feature extractor emits) -> C8 [B][C/8][S][8], same storage type: each
// thread moves one 16/32-byte channel block, consecutive threads walk the channel blocks of one pixel (coalesced read).
template <typename T>
__global__ void nhwc_to_c8_kernel(const T* __restrict__ src, T* __restrict__ dst, int CB, int64_t S, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // i = (b * S + s) * CB + cb
    if (i >= total) return;
    const int cb = (int)(i % CB);
    const int64_t bs = i / CB;
    const int64_t s = bs % S, b = bs / S;
    float v[8];
    V8<T>::load(src + i * 8, v);
    V8<T>::store(dst + ((b * CB + cb) * S + s) * 8, v);
}

extern "C" int mvs_nhwc_to_c8(const void* src, void* dst, int B, int C, int64_t S, int dtype, void* stream) {
    MVS_REQUIRE(src && dst, MVS_E_ARG, "mvs_nhwc_to_c8: null pointer");
    MVS_REQUIRE(B > 0 && C > 0 && S > 0 && C % 8 == 0, MVS_E_SHAPE, "mvs_nhwc_to_c8: C=%d must be a positive multiple of 8", C);
    const int64_t total = (int64_t)B * S * (C / 8);
    MVS_DISPATCH_DTYPE(dtype, T, MVS_LAUNCH(nhwc_to_c8_kernel<T>, dim3(mvs_cdiv(total, 256)), dim3(256), stream, (const T*)src, (T*)dst, C / 8, S, total));
    return MVS_CHECK_LAUNCH("mvs_nhwc_to_c8");
}

extern "C" int mvs_pack_c8(const float* src, void* dst, int B, int C, int64_t S, int dtype, void* stream) {
    MVS_REQUIRE(src && dst, MVS_E_ARG, "mvs_pack_c8: null pointer");
    MVS_REQUIRE(B > 0 && C > 0 && S > 0 && C % 8 == 0, MVS_E_SHAPE, "mvs_pack_c8: C=%d must be a positive multiple of 8", C);
    const int64_t total = (int64_t)B * (C / 8) * S;
    MVS_DISPATCH_DTYPE(dtype, T, MVS_LAUNCH(pack_c8_kernel<T>, dim3(mvs_cdiv(total, 256)), dim3(256), stream, src, (T*)dst, C, S, total));
    return MVS_CHECK_LAUNCH("mvs_pack_c8");
}

extern "C" int mvs_unpack_c8(const void* src, float* dst, int B, int C, int64_t S, int dtype, void* stream) {
    MVS_REQUIRE(src && dst, MVS_E_ARG, "mvs_unpack_c8: null pointer");
    MVS_REQUIRE(B > 0 && C > 0 && S > 0 && C % 8 == 0, MVS_E_SHAPE, "mvs_unpack_c8: C=%d must be a positive multiple of 8", C);
    const int64_t total = (int64_t)B * (C / 8) * S;
    MVS_DISPATCH_DTYPE(dtype, T, MVS_LAUNCH(unpack_c8_kernel<T>, dim3(mvs_cdiv(total, 256)), dim3(256), stream, (const T*)src, dst, C, S, total));
    return MVS_CHECK_LAUNCH("mvs_unpack_c8");
}

// Zero-bordered C8 maps ("C8P") for the plane-sweep gather: dst [M][C/8][H + 3][W + 2][8], pixel (y, x) at row y + 1,
// column x + 1, zeros elsewhere.  One thread writes one (map, channel block, padded pixel) vector; src_layout selects the
// reader: 0 = fp32 [M][C][H][W], 1 = channels-last `T` [M][H][W][C], 2 = C8 `T` [M][C/8][H][W][8].
template <typename T>
__global__ void pack_c8_padded_kernel(const void* __restrict__ src, T* __restrict__ dst, int CB, int H, int W, int layout, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // ((m * CB + cb) * (H + 3) + yp) * (W + 2) + xp
    if (i >= total) return;
    const int Wp = W + 2, Hp = H + 3;
    const int xp = (int)(i % Wp);
    int64_t r = i / Wp;
    const int yp = (int)(r % Hp); r /= Hp;
    const int cb = (int)(r % CB);
    const int64_t m = r / CB;
    const int x = xp - 1, y = yp - 1;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.f;
    if (x >= 0 && x < W && y >= 0 && y < H) {
        const int64_t S = (int64_t)H * W, s = (int64_t)y * W + x;
        if (layout == 0) {
            const float* p = reinterpret_cast<const float*>(src) + (m * CB * 8 + (int64_t)cb * 8) * S + s;
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldg(p + (int64_t)k * S);
        } else if (layout == 1) {
            V8<T>::load(reinterpret_cast<const T*>(src) + ((m * S + s) * CB + cb) * 8, v);
        } else {
            V8<T>::load(reinterpret_cast<const T*>(src) + ((m * CB + cb) * S + s) * 8, v);
        }
    }
    V8<T>::store(dst + i * 8, v);
}

extern "C" int mvs_pack_c8_padded(const void* src, void* dst, int M, int C, int H, int W, int src_layout, int dtype, void* stream) {
    MVS_REQUIRE(src && dst, MVS_E_ARG, "mvs_pack_c8_padded: null pointer");
    MVS_REQUIRE(M > 0 && C > 0 && H > 0 && W > 0 && C % 8 == 0, MVS_E_SHAPE, "mvs_pack_c8_padded: C=%d must be a positive multiple of 8", C);
    MVS_REQUIRE(src_layout >= 0 && src_layout <= 2, MVS_E_ARG, "mvs_pack_c8_padded: src_layout must be 0 (fp32 NCHW), 1 (NHWC) or 2 (C8)");
    const int64_t total = (int64_t)M * (C / 8) * (H + 3) * (W + 2);
    MVS_DISPATCH_DTYPE(dtype, T, MVS_LAUNCH(pack_c8_padded_kernel<T>, dim3(mvs_cdiv(total, 256)), dim3(256), stream, src, (T*)dst, C / 8, H, W, src_layout, total));
    return MVS_CHECK_LAUNCH("mvs_pack_c8_padded");
}

// images [B][N][3][H][W] (fp32, or already 16-bit as a host pipeline uploads them) -> C8 image stack [M = N*B][H][W][8]
// (m = v * B + b), channels 3..7 zero: the input of the tcgen05 feature extractor.  One thread per pixel: three coalesced
// plane reads, one 16-byte store.
template <typename TS> __device__ __forceinline__ float img_load(const TS* p) { return (float)__ldg(p); }
#ifndef MVS_CPU_EMU
template <> __device__ __forceinline__ float img_load<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float img_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }
#endif
template <typename TS, typename T>
__global__ void pack_images_c8_kernel(const TS* __restrict__ imgs, T* __restrict__ dst, int B, int N, int64_t S, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // m * S + s
    if (i >= total) return;
    const int64_t s = i % S;
    const int m = (int)(i / S);
    const int v = m / B, b = m % B;
    const TS* p = imgs + (((int64_t)b * N + v) * 3) * S + s;
    float o[8];
    o[0] = img_load<TS>(p); o[1] = img_load<TS>(p + S); o[2] = img_load<TS>(p + 2 * S);
#pragma unroll
    for (int k = 3; k < 8; ++k) o[k] = 0.f;
    V8<T>::store(dst + i * 8, o);
}

extern "C" int mvs_pack_images_c8(const void* imgs, int src_dtype, void* dst, int B, int N, int H, int W, int dtype, void* stream) {
    MVS_REQUIRE(imgs && dst, MVS_E_ARG, "mvs_pack_images_c8: null pointer");
    MVS_REQUIRE(B > 0 && N > 0 && H > 0 && W > 0, MVS_E_SHAPE, "mvs_pack_images_c8: bad dims");
    const int64_t S = (int64_t)H * W, total = (int64_t)B * N * S;
    MVS_DISPATCH_DTYPE(src_dtype, TS, MVS_DISPATCH_DTYPE(dtype, T,
        MVS_LAUNCH((pack_images_c8_kernel<TS, T>), dim3(mvs_cdiv(total, 256)), dim3(256), stream, (const TS*)imgs, (T*)dst, B, N, S, total)));
    return MVS_CHECK_LAUNCH("mvs_pack_images_c8");
}

// ---------------------------------------------------------------------------------------------- projections
__device__ static void rel_rt(const double* src, const double* ref, float* rt) {
    double inv[16];
    inv4(ref, inv);
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += src[r * 4 + k] * inv[k * 4 + c];
            if (c < 3) rt[r * 3 + c] = (float)s; else rt[9 + r] = (float)s;
        }
    }
}

__global__ void compose_proj_kernel(const float* __restrict__ proj, float* __restrict__ rt, int B, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // i = s * B + b
    if (i >= (N - 1) * B) return;
    const int s = i / B, b = i % B;
    double ref[16], src[16];
    for (int k = 0; k < 16; ++k) {
        ref[k] = (double)proj[((int64_t)b * N) * 16 + k];
        src[k] = (double)proj[((int64_t)b * N + s + 1) * 16 + k];
    }
    rel_rt(src, ref, rt + (int64_t)i * 12);
}

__device__ static void ke_to_proj(const float* K, const float* E, float down, double* P) {
    // the reference divides K[:2] by the level ratio and multiplies K @ E[:3] in fp32; the products are
    // formed in fp64 here from the same fp32 inputs.
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) {
                const double kv = (r < 2) ? (double)(K[r * 3 + k] / down) : (double)K[r * 3 + k];
                s += kv * (double)E[k * 4 + c];
            }
            P[r * 4 + c] = s;
        }
    P[12] = 0.0; P[13] = 0.0; P[14] = 0.0; P[15] = 1.0;
}

__global__ void compose_proj_ke_kernel(const float* __restrict__ ref_in, const float* __restrict__ src_in,
                                       const float* __restrict__ ref_ex, const float* __restrict__ src_ex, float down,
                                       float* __restrict__ rt, int B, int nsrc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // i = s * B + b
    if (i >= nsrc * B) return;
    const int s = i / B, b = i % B;
    double ref[16], src[16];
    ke_to_proj(ref_in + (int64_t)b * 9, ref_ex + (int64_t)b * 16, down, ref);
    ke_to_proj(src_in + ((int64_t)b * nsrc + s) * 9, src_ex + ((int64_t)b * nsrc + s) * 16, down, src);
    rel_rt(src, ref, rt + (int64_t)i * 12);
}

extern "C" int mvs_compose_proj(const float* proj, float* rt, int B, int N, void* stream) {
    MVS_REQUIRE(proj && rt, MVS_E_ARG, "mvs_compose_proj: null pointer");
    MVS_REQUIRE(B > 0 && N >= 2 && N - 1 <= MVS_MAX_SRC, MVS_E_SHAPE, "mvs_compose_proj: need 2 <= N <= %d views, got %d", MVS_MAX_SRC + 1, N);
    MVS_LAUNCH(compose_proj_kernel, dim3(mvs_cdiv((N - 1) * B, 64)), dim3(64), stream, proj, rt, B, N);
    return MVS_CHECK_LAUNCH("mvs_compose_proj");
}

extern "C" int mvs_compose_proj_ke(const float* ref_in, const float* src_in, const float* ref_ex, const float* src_ex,
                                   float down, float* rt, int B, int nsrc, void* stream) {
    MVS_REQUIRE(ref_in && src_in && ref_ex && src_ex && rt, MVS_E_ARG, "mvs_compose_proj_ke: null pointer");
    MVS_REQUIRE(B > 0 && nsrc >= 1 && nsrc <= MVS_MAX_SRC, MVS_E_SHAPE, "mvs_compose_proj_ke: need 1 <= nsrc <= %d, got %d", MVS_MAX_SRC, nsrc);
    MVS_REQUIRE(down > 0.f, MVS_E_ARG, "mvs_compose_proj_ke: down must be positive");
    MVS_LAUNCH(compose_proj_ke_kernel, dim3(mvs_cdiv(nsrc * B, 64)), dim3(64), stream, ref_in, src_in, ref_ex, src_ex, down, rt, B, nsrc);
    return MVS_CHECK_LAUNCH("mvs_compose_proj_ke");
}
