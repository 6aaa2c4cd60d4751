// featnet_front.cu — the full-resolution front of FeatureNet (jdacs/models/mvsnet.py:20-23, 39-40) as ONE kernel:
//
//     image [3, H, W]  ->  conv0 3x3 (3 -> 8) + BN + ReLU  ->  conv1 3x3 (8 -> 8) + BN + ReLU  ->  conv2 5x5 stride 2 (8 -> 16) + BN + ReLU
//
// Why: these three layers work on 8-channel FULL-RESOLUTION maps (512 x 640 x 5 views per item) and do almost no arithmetic per
// byte: run layer by layer (pack + three launches of the tcgen05 kernel) they move the 8-channel maps through HBM three times
// (~1.3 GB per 8-item step) and take 0.6 ms of a 4.9 ms step; fused, the image is read once, the 8-channel maps live in shared
// memory only, and the half-resolution 16-channel map is written once (the layout the next tcgen05 layer consumes).
//
// A CTA (8 warps) owns an 8 x 16 tile of conv2's output and recomputes the halos it needs:
//     image tile 23 x 39 (zero outside the image = conv0's padding)           -> smem rows of 8 channels (3 used), 16 bytes each
//     conv0 on 21 x 37 positions (zero outside the image = conv1's padding)   -> smem [777][8]
//     conv1 on 19 x 35 positions (zero outside the image = conv2's padding)   -> smem [665][8]
//     conv2 on  8 x 16 positions                                              -> global, C8 stack [2][M][H/2][W/2][8]
// Every stage is an implicit GEMM on warp-level tensor-core MMAs (mma.sync m16n8k16, fp32 accumulate): M = 16 positions, K = two
// filter taps x 8 channels, N = 8 output channels (the 3x3 stages fold kw into N: see stage3x3).  The 16-byte channel rows of the staged maps are exactly ldmatrix rows, so the A
// fragment of a (tap pair, 16 positions) is one ldmatrix.x4 whose 32 row addresses carry the tap shifts (and, for conv2, the
// stride); B fragments (weights in the storage type, packed on the host in fragment order) sit in shared memory.  Folded
// BatchNorm + ReLU run on the accumulator registers.  Arithmetic is the layered path's: 16-bit operands, fp32 accumulation,
// fp32 affine, one rounding to the storage type per layer.
#include "mvs_rt.h"

#ifndef MVS_CPU_EMU
namespace {

#ifndef FF_MINB
#define FF_MINB 3
#endif
constexpr int kTOH = 8, kTOW = 16;                         // conv2 output tile
constexpr int kR1H = 2 * kTOH + 3, kR1W = 2 * kTOW + 3;    // 19 x 35 conv1 outputs
constexpr int kR0H = kR1H + 2, kR0W = kR1W + 2;            // 21 x 37 conv0 outputs
constexpr int kRIH = kR0H + 2, kRIW = kR0W + 2;            // 23 x 39 image pixels
constexpr int kU3 = 6, kKS5 = 13;                          // B-fragment units of a 3x3 stage (2 kh pairs x 3 kw); k-steps (tap pairs) of the 5x5 filter
constexpr int kFragWords = (kU3 + kU3 + 2 * kKS5) * 32 * 2;     // B fragments: per unit 32 lanes x 2 words
constexpr int kR1Even = (kR1W + 1) / 2;                    // conv1's output is staged de-interleaved by column parity: even columns first

// byte offset, inside the parity-staged conv1 output, of tap `tap` of the 5x5 stride-2 filter relative to (row 2 oy, pair column ox)
constexpr uint32_t ff_c2_off(int tap) {
    return (uint32_t)(((tap > 24 ? 24 : tap) / 5) * kR1W + (((tap > 24 ? 24 : tap) % 5) & 1) * kR1Even + (((tap > 24 ? 24 : tap) % 5) >> 1)) * 16u;
}

__device__ __forceinline__ uint32_t ff_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ff_ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T> __device__ __forceinline__ void ff_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void ff_mma<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void ff_mma<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <typename T> __device__ __forceinline__ uint32_t ff_pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t ff_pack2<__half>(float a, float b) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
template <> __device__ __forceinline__ uint32_t ff_pack2<__nv_bfloat16>(float a, float b) { const __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
template <typename T> __device__ __forceinline__ T ff_from_float(float v);
template <> __device__ __forceinline__ __half ff_from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 ff_from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <typename T> __device__ __forceinline__ float ff_to_float(T v);
template <> __device__ __forceinline__ float ff_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float ff_to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float ff_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// One 3x3 stage: out region OH x OW from the input region (OH + 2) x IW (IW = OW + 2), both staged as 16-byte channel rows.
// The kw taps are folded into N, as in the tcgen05 kernel: an m-tile is 16 CONSECUTIVE INPUT positions p (pitch-linear over the
// input region), one ldmatrix.x4 per kh pair brings rows p + kh IW, and three MMAs (one per kw) give
//     D_kw[p][co] = sum_{kh, ci} in[p + kh IW][ci] W[kh][kw][ci][co],       out[o] = D_0[o] + D_1[o + 1] + D_2[o + 2]
// -- 2 fragment loads per m-tile instead of 5 (the shared-memory pipe bounds this kernel), at the price of 8 shuffles that
// fetch rows o + 1, o + 2 from the neighbouring lanes, 14 finished outputs per 16-row tile, and junk at the two columns where a
// row of the region wraps.  Positions outside the image are written as zeros (they are the NEXT layer's padding).
// PAR: store the output de-interleaved by column parity (the stride-2 layer that reads it then gets contiguous ldmatrix rows).
template <typename T, int OH, int OW, bool PAR>
__device__ __forceinline__ void stage3x3(uint32_t in_base, uint8_t* out, const uint32_t* __restrict__ frag, const float* __restrict__ aff,
                                         int y0, int x0, int H, int W, int warp, int lane) {
    // tiles whose whole output region lies inside the image (all but the border tiles) skip the per-position bounds test
    const bool interior = y0 >= 0 && x0 >= 0 && y0 + OH <= H && x0 + OW <= W;
    constexpr int IW = OW + 2, NPOS = OH * IW, MT = (NPOS + 13) / 14;
    const int lj = lane >> 3, li = lane & 7, g = lane >> 2, q = lane & 3;
    const float sc0 = aff[2 * q], sc1 = aff[2 * q + 1], sh0 = aff[16 + 2 * q], sh1 = aff[16 + 2 * q + 1];
    uint2 bfr[2][3];                                                     // the stage's weights stay in registers
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) bfr[ks][kw] = *reinterpret_cast<const uint2*>(frag + ((ks * 3 + kw) * 32 + lane) * 2);
    // this lane's ldmatrix row: position li (+ 8 for matrices 1, 3), tap row kh = 2 ks + (lj >> 1) (kh = 3 has zero weights: reads row 2)
    const uint32_t lane_row = in_base + (uint32_t)(li + (lj & 1) * 8) * 16u;
    const uint32_t koff0 = (uint32_t)((lj >> 1) * IW) * 16u, koff1 = (uint32_t)(2 * IW) * 16u;
    const int src4 = (lane + 4) & 31, src8 = (lane + 8) & 31;
    for (int mt = warp; mt < MT; mt += 8) {
        const int p0 = mt * 14;
        float acc[3][4];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) { acc[kw][0] = acc[kw][1] = acc[kw][2] = acc[kw][3] = 0.f; }
        uint32_t a[4];
        ff_ldmatrix_x4(lane_row + (uint32_t)p0 * 16u + koff0, a);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) ff_mma<T>(acc[kw], a, bfr[0][kw].x, bfr[0][kw].y);
        ff_ldmatrix_x4(lane_row + (uint32_t)p0 * 16u + koff1, a);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) ff_mma<T>(acc[kw], a, bfr[1][kw].x, bfr[1][kw].y);
        // rows o + 1 / o + 2 of D_1 / D_2: lane g + 1 / g + 2, or (rows 8, 9) the second half of lanes g = 0, 1
        float o1[2], o2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float a1 = __shfl_sync(0xffffffffu, acc[1][e], src4), b1 = __shfl_sync(0xffffffffu, acc[1][2 + e], src4);
            const float a2 = __shfl_sync(0xffffffffu, acc[2][e], src8), b2 = __shfl_sync(0xffffffffu, acc[2][2 + e], src8);
            o1[e] = acc[0][e] + (g < 7 ? a1 : b1) + (g < 6 ? a2 : b2);     // row g
            o2[e] = acc[0][2 + e] + b1 + b2;                               // row g + 8 (complete for g <= 5)
        }
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int o = p0 + g + hrow * 8;
            const int py = o / IW, px = o - py * IW;
            if ((hrow == 0 || g <= 5) && o < NPOS && px < OW) {
                const bool inside = interior || ((unsigned)(y0 + py) < (unsigned)H && (unsigned)(x0 + px) < (unsigned)W);
                const float v0 = fmaxf((hrow ? o2[0] : o1[0]) * sc0 + sh0, 0.f), v1 = fmaxf((hrow ? o2[1] : o1[1]) * sc1 + sh1, 0.f);
                const int idx = PAR ? py * OW + (px & 1) * ((OW + 1) / 2) + (px >> 1) : py * OW + px;
                *reinterpret_cast<uint32_t*>(out + (size_t)idx * 16 + q * 4) = inside ? ff_pack2<T>(v0, v1) : 0u;
            }
        }
    }
}

template <typename T, typename TS>
__global__ void __launch_bounds__(256, FF_MINB)
featnet_front_kernel(const TS* __restrict__ imgs, const uint32_t* __restrict__ wfrag, const float* __restrict__ affine, T* __restrict__ out,
                     int B, int N, int H, int W, int tiles_x, int tiles_y) {
    // (+16 rows: the last m-tile of a 3x3 stage reads up to 15 + 2 IW rows past its first position; those rows only feed
    // accumulator rows that are never stored)
    __shared__ __align__(16) uint8_t s_img[(kRIH * kRIW + 16) * 16];
    __shared__ __align__(16) uint8_t s_r0[(kR0H * kR0W + 16) * 16];
    __shared__ __align__(16) uint8_t s_r1[kR1H * kR1W * 16];
    __shared__ __align__(16) uint32_t s_frag[kFragWords];
    __shared__ float s_aff[3][32];                                         // per layer: scale[16] | shift[16]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kFragWords; i += blockDim.x) s_frag[i] = __ldg(wfrag + i);
    if (threadIdx.x < 96) s_aff[threadIdx.x >> 5][threadIdx.x & 31] = __ldg(affine + threadIdx.x);
    const int Ho = H / 2, Wo = W / 2, M = B * N;
    const int64_t plane = (int64_t)H * W;
    const int tiles = tiles_x * tiles_y;
    bool img_zeroed = false;

    for (int t = blockIdx.x; t < tiles * M; t += gridDim.x) {
        const int m = t / tiles, tt = t - m * tiles;                       // image m = v * B + b of the output stack
        const int ty = tt / tiles_x, tx = tt - ty * tiles_x;
        const int v = m / B, b = m - v * B;
        const TS* im = imgs + ((int64_t)b * N + v) * 3 * plane;
        const int oy0 = ty * kTOH, ox0 = tx * kTOW;
        const int iy0 = 2 * oy0 - 4, ix0 = 2 * ox0 - 4;                    // image-space origin of the staged image tile
        __syncthreads();                                                   // previous tile fully consumed (and s_frag / s_aff visible)
        // ---- image tile -> 16-byte rows (3 channels + zeros), zero outside the image
        if ((W & 3) == 0) {
            // one work item = 4 consecutive pixels of one row of one channel plane: ONE vector load (the tile starts at a multiple
            // of 4 pixels and W % 4 == 0, so a group is wholly inside or wholly outside the image), four 2-byte stores into the
            // channel slot of the four rows.  Slots 3..7 of every row are zeroed once per CTA and never written again.
            constexpr int kGroups = (kRIW + 3) / 4;
            if (!img_zeroed) {
                for (int i = threadIdx.x; i < kRIH * kRIW; i += blockDim.x) *reinterpret_cast<uint4*>(s_img + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
                img_zeroed = true;
                __syncthreads();
            }
            for (int i = threadIdx.x; i < kRIH * 3 * kGroups; i += blockDim.x) {
                const int gx = i % kGroups, c = (i / kGroups) % 3, ry = i / (3 * kGroups);
                const int y = iy0 + ry, x = ix0 + 4 * gx;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) {
                    const TS* src = im + (int64_t)c * plane + (int64_t)y * W + x;
                    if (sizeof(TS) == 4) {
                        const float4 f = __ldg(reinterpret_cast<const float4*>(src));
                        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
                    } else {
                        const uint2 r = __ldg(reinterpret_cast<const uint2*>(src));
                        const TS* h = reinterpret_cast<const TS*>(&r);
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = ff_to_float<TS>(h[e]);
                    }
                }
                T* dst = reinterpret_cast<T*>(s_img) + (size_t)(ry * kRIW + 4 * gx) * 8 + c;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (4 * gx + e < kRIW) dst[e * 8] = ff_from_float<T>(v[e]);
            }
        } else {
            for (int i = threadIdx.x; i < kRIH * kRIW; i += blockDim.x) {
                const int ry = i / kRIW, rx = i - ry * kRIW;
                const int y = iy0 + ry, x = ix0 + rx;
                uint4 row = make_uint4(0u, 0u, 0u, 0u);
                if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) {
                    const int64_t o = (int64_t)y * W + x;
                    const float c0 = ff_to_float<TS>(__ldg(im + o)), c1 = ff_to_float<TS>(__ldg(im + plane + o)), c2 = ff_to_float<TS>(__ldg(im + 2 * plane + o));
                    row.x = ff_pack2<T>(c0, c1); row.y = ff_pack2<T>(c2, 0.f);
                }
                *reinterpret_cast<uint4*>(s_img + (size_t)i * 16) = row;
            }
        }
        __syncthreads();
        stage3x3<T, kR0H, kR0W, false>(ff_smem_u32(s_img), s_r0, s_frag, s_aff[0], iy0 + 1, ix0 + 1, H, W, warp, lane);
        __syncthreads();
        stage3x3<T, kR1H, kR1W, true>(ff_smem_u32(s_r0), s_r1, s_frag + kU3 * 64, s_aff[1], iy0 + 2, ix0 + 2, H, W, warp, lane);
        __syncthreads();
        // ---- conv2: 5x5, stride 2, 8 -> 16: one m-tile (16 of the 128 output positions = one tile row) per warp, two n-tiles
        {
            const int lj = lane >> 3, li = lane & 7, g = lane >> 2, q = lane & 3;
            const int ox_l = li + (lj & 1) * 8;                            // m-tile = output row `warp`, position = column ox_l
            // conv1's output is de-interleaved by column parity: column 2 ox + kw sits at (kw & 1) * kR1Even + ox + (kw >> 1)
            const uint32_t row0 = ff_smem_u32(s_r1) + (uint32_t)((2 * warp) * kR1W + ox_l) * 16u;
            const uint32_t* f2 = s_frag + 2 * kU3 * 64;
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < kKS5; ++ks) {
                // matrices 0, 1 read tap 2 ks, matrices 2, 3 tap 2 ks + 1: two compile-time offsets and one select per k-step
                uint32_t a[4];
                ff_ldmatrix_x4(row0 + ((lj >> 1) ? ff_c2_off(2 * ks + 1) : ff_c2_off(2 * ks)), a);
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    const uint2 bb = *reinterpret_cast<const uint2*>(f2 + ((ks * 2 + n) * 32 + lane) * 2);
                    ff_mma<T>(acc[n], a, bb.x, bb.y);
                }
            }
            const int oy = oy0 + warp;
            if (oy < Ho) {
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    const float sc0 = s_aff[2][n * 8 + 2 * q], sc1 = s_aff[2][n * 8 + 2 * q + 1];
                    const float sh0 = s_aff[2][16 + n * 8 + 2 * q], sh1 = s_aff[2][16 + n * 8 + 2 * q + 1];
#pragma unroll
                    for (int hrow = 0; hrow < 2; ++hrow) {
                        const int ox = ox0 + g + hrow * 8;
                        if (ox < Wo) {
                            const float v0 = fmaxf(acc[n][2 * hrow] * sc0 + sh0, 0.f), v1 = fmaxf(acc[n][2 * hrow + 1] * sc1 + sh1, 0.f);
                            T* o = out + ((((int64_t)n * M + m) * Ho + oy) * Wo + ox) * 8 + 2 * q;
                            *reinterpret_cast<uint32_t*>(o) = ff_pack2<T>(v0, v1);
                        }
                    }
                }
            }
        }
    }
}

// Weights -> B fragments of mma.m16n8k16 (col-major B): lane (g = lane / 4, q = lane % 4) holds for k-step ks and n-tile nt the
// words { B[2q][g], B[2q+1][g] } and { B[2q+8][g], B[2q+9][g] } with B[k][n] = w[nt * 8 + n][c = k % 8][tap = 2 ks + k / 8].
template <typename T>
__global__ void featnet_front_pack_kernel(const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ w2,
                                          uint32_t* __restrict__ frag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;                   // one 32-bit word each
    if (i >= kFragWords) return;
    const int word = i & 1, lane = (i >> 1) & 31, unit = i >> 6;          // unit over the three layers: 6 + 6 + 26
    const int g = lane >> 2, q = lane & 3;
    float v[2];
    for (int e = 0; e < 2; ++e) {
        const int k = 2 * q + e + word * 8, c = k % 8;
        float w = 0.f;
        if (unit < 2 * kU3) {
            // 3x3 stages, kw folded into N: unit = ks * 3 + kw, K = kh pair (2 ks, 2 ks + 1) x 8 channels, N = 8 output channels
            const int layer = unit / kU3, u = unit % kU3, ks = u / 3, kw = u % 3, kh = 2 * ks + k / 8, n = g;
            if (kh < 3) {
                if (layer == 0) { if (c < 3) w = w0[(n * 3 + c) * 9 + kh * 3 + kw]; }
                else w = w1[(n * 8 + c) * 9 + kh * 3 + kw];
            }
        } else {
            // 5x5 stride 2: unit = ks * 2 + n-tile, K = tap pair (2 ks, 2 ks + 1) x 8 channels
            const int u = unit - 2 * kU3, ks = u >> 1, nt = u & 1, tap = 2 * ks + k / 8, n = nt * 8 + g;
            if (tap < 25) w = w2[(n * 8 + c) * 25 + tap];
        }
        v[e] = w;
    }
    frag[i] = ff_pack2<T>(v[0], v[1]);
}

}  // namespace

extern "C" int64_t mvs_featnet_front_workspace_bytes(void) { return (int64_t)kFragWords * 4; }

extern "C" int mvs_featnet_front_pack(const float* w0, const float* w1, const float* w2, void* wfrag, int dtype, void* stream) {
    MVS_REQUIRE(w0 && w1 && w2 && wfrag, MVS_E_ARG, "mvs_featnet_front_pack: null pointer");
    MVS_REQUIRE(dtype == MVS_F16 || dtype == MVS_BF16, MVS_E_UNSUPPORTED, "mvs_featnet_front_pack: 16-bit storage only");
    const dim3 grid(mvs_cdiv(kFragWords, 256));
    if (dtype == MVS_F16) featnet_front_pack_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(w0, w1, w2, (uint32_t*)wfrag);
    else featnet_front_pack_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(w0, w1, w2, (uint32_t*)wfrag);
    return MVS_CHECK_LAUNCH("mvs_featnet_front_pack");
}

extern "C" int mvs_featnet_front(const void* imgs, int img_dtype, const void* wfrag, const float* affine, void* out, int B, int N,
                                 int H, int W, int dtype, void* stream) {
    MVS_REQUIRE(imgs && wfrag && affine && out, MVS_E_ARG, "mvs_featnet_front: null pointer");
    MVS_REQUIRE(dtype == MVS_F16 || dtype == MVS_BF16, MVS_E_UNSUPPORTED, "mvs_featnet_front: 16-bit storage only");
    MVS_REQUIRE(B > 0 && N > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, MVS_E_SHAPE, "mvs_featnet_front: bad dims (even H, W)");
    const int tiles_x = (W / 2 + kTOW - 1) / kTOW, tiles_y = (H / 2 + kTOH - 1) / kTOH;
    const int64_t total = (int64_t)tiles_x * tiles_y * B * N;
    MVS_REQUIRE(total < (1ll << 31), MVS_E_SHAPE, "mvs_featnet_front: too many tiles");
    const unsigned grid = (unsigned)(total < 148 * FF_MINB ? total : 148 * FF_MINB);
    cudaStream_t st = (cudaStream_t)stream;
#define MVS_FF_LAUNCH(T, TS) featnet_front_kernel<T, TS><<<grid, 256, 0, st>>>((const TS*)imgs, (const uint32_t*)wfrag, affine, (T*)out, B, N, H, W, tiles_x, tiles_y)
    if (dtype == MVS_F16) {
        if (img_dtype == MVS_F32) MVS_FF_LAUNCH(__half, float); else if (img_dtype == MVS_F16) MVS_FF_LAUNCH(__half, __half);
        else return mvs_set_error(MVS_E_UNSUPPORTED, "mvs_featnet_front: images must be fp32 or stored in the volume dtype");
    } else {
        if (img_dtype == MVS_F32) MVS_FF_LAUNCH(__nv_bfloat16, float); else if (img_dtype == MVS_BF16) MVS_FF_LAUNCH(__nv_bfloat16, __nv_bfloat16);
        else return mvs_set_error(MVS_E_UNSUPPORTED, "mvs_featnet_front: images must be fp32 or stored in the volume dtype");
    }
#undef MVS_FF_LAUNCH
    return MVS_CHECK_LAUNCH("mvs_featnet_front");
}
#else
extern "C" int64_t mvs_featnet_front_workspace_bytes(void) { return 0; }
extern "C" int mvs_featnet_front_pack(const float*, const float*, const float*, void*, int, void*) {
    return mvs_set_error(MVS_E_UNSUPPORTED, "mvs_featnet_front_pack: tensor-core kernels do not exist in the emulation build");
}
extern "C" int mvs_featnet_front(const void*, int, const void*, const float*, void*, int, int, int, int, int, void*) {
    return mvs_set_error(MVS_E_UNSUPPORTED, "mvs_featnet_front: tensor-core kernels do not exist in the emulation build");
}
#endif
