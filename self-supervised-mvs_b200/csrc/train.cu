// train.cu — the training half of the regularisation U-Net with 16-bit activations (fp32 master weights, fp32 accumulation):
//
//   forward   z = conv(x)            tcgen05 kernel of conv3d_tc.cu, no affine (scale/shift NULL)
//             sums = stats(z)        mvs_bn_stats_t      per-channel sum / sum of squares, fp64 accumulation across blocks
//             a, b, mean, invstd     mvs_bn_finalize     batch statistics -> per-channel affine, running-stat update
//             y = relu(z a + b) + s  mvs_bn_act_fwd_t
//   backward  red = reduce(z, gy)    mvs_bn_act_bwd_reduce_t   sum g, sum g xhat  (g = gy [z a + b > 0])
//             gz, ggamma, gbeta      mvs_bn_act_bwd_apply_t
//             gx = conv^T(gz)        tcgen05 kernel again with the adjoint tap program (the same torch weight packed under the
//                                    opposite `transposed` flag), no entry point of its own
//             gw = x (*) gz          mvs_conv3d_wgrad_mma      this file: warp-level tensor-core MMAs (mma.sync m16n8k16)
//
// Why the weight gradient is NOT a tcgen05 kernel: gw[tap][ci][co] = sum_voxels x[v + tap][ci] gz[v][co] is a GEMM whose M x N is
// at most 64 x 64 (per tap) and whose K is the voxel count.  One tcgen05.mma covers K = 16 and costs >= ~32 cycles of operand
// fetch whatever M x N is (tools/umma_bench.cu), so at M x N = 32 x 8 (conv0, 68 % of the flops) it would run at ~120 MAC/clk per
// SM, ten times below the forward kernel; four independently issuing warp schedulers doing m16n8k16 MMAs on shifted views of one
// shared-memory tile are the right tool for tiny-MN / huge-K products.
// Reference: the autograd of ConvBnReLU3D / ConvTranspose3d + BatchNorm3d + ReLU, jdacs/models/module.py:35-42, mvsnet.py:37-74,
// jdacs-ms/models/network.py:44-74 (PyTorch derives these gradients; nothing is written out in the reference).
#include "mvs_rt.h"

// ------------------------------------------------------------------------------------------------ batch-norm passes, any storage type
// C8 volumes [B][C/8][S][8]; blockIdx.y = b * CB + cb; threads stride over s.

// Per-thread partial sums of 2 x 8 channels -> fp64 totals: a thread sums a few dozen values and a warp 32 threads in fp32;
// everything above that is accumulated in fp64 (E[x^2] - E[x]^2 cancels when |mean| >> std).  The block's 8 warps
// meet in shared memory, so a block issues 16 fp64 atomics (it was 128: on the small feature-extractor layers the same-address
// atomics, not the 5-10 MB read, set the kernel time).
__device__ __forceinline__ void bn_block_sums(float (&p0)[8], float (&p1)[8], double* dst0, double* dst1) {
#ifndef MVS_CPU_EMU
    __shared__ float part[8][16];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { p0[k] += __shfl_xor_sync(0xffffffffu, p0[k], o); p1[k] += __shfl_xor_sync(0xffffffffu, p1[k], o); }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { part[warp][k] = p0[k]; part[warp][8 + k] = p1[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)part[w][threadIdx.x];
        atomicAdd(threadIdx.x < 8 ? dst0 + threadIdx.x : dst1 + (threadIdx.x - 8), t);
    }
#else
    for (int k = 0; k < 8; ++k) { atomicAdd(dst0 + k, (double)p0[k]); atomicAdd(dst1 + k, (double)p1[k]); }
#endif
}
template <typename T>
__global__ void __launch_bounds__(256)
bn_stats_t_kernel(const T* __restrict__ z, double* __restrict__ sums, int C, int64_t S) {
    const int CB = C / 8;
    const int cb = blockIdx.y % CB;
    const int64_t base = (int64_t)blockIdx.y * S;
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        float v[8];
        V8<T>::load(z + (base + s) * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s1[k] += v[k]; s2[k] += v[k] * v[k]; }
    }
    bn_block_sums(s1, s2, sums + cb * 8, sums + C + cb * 8);
}

// sums -> a = gamma * invstd, b = beta - mean * a, mean, invstd (fp32), and the running statistics exactly as nn.BatchNorm3d
// updates them in training mode (momentum, unbiased variance).  One thread per channel.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, double count, float* __restrict__ a, float* __restrict__ b,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sums[c] / count;
    double var = sums[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float av = gamma[c] * is;
    a[c] = av; b[c] = beta[c] - (float)m * av; mean[c] = (float)m; invstd[c] = is;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * (count / (count > 1.0 ? count - 1.0 : 1.0)));
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_fwd_t_kernel(const T* __restrict__ z, const float* __restrict__ a, const float* __restrict__ b, const T* __restrict__ skip,
                    T* __restrict__ y, int C, int64_t S, int64_t total, int relu) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cb = (int)((i / S) % (C / 8));
    float v[8];
    V8<T>::load(z + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float o = v[k] * __ldg(a + cb * 8 + k) + __ldg(b + cb * 8 + k);
        if (relu) o = fmaxf(o, 0.f);
        v[k] = o;
    }
    if (skip) {
        float sv[8];
        V8<T>::load(skip + i * 8, sv);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += sv[k];
    }
    V8<T>::store(y + i * 8, v);
}

// g = gy [z a + b > 0];  red[0][c] += g (= grad beta), red[1][c] += g * xhat (= grad gamma), xhat = (z - mean) invstd
template <typename T>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_t_kernel(const T* __restrict__ z, const T* __restrict__ gy, const float* __restrict__ a, const float* __restrict__ b,
                           const float* __restrict__ mean, const float* __restrict__ invstd, double* __restrict__ red, int C,
                           int64_t S, int relu) {
    const int CB = C / 8;
    const int cb = blockIdx.y % CB;
    const int64_t base = (int64_t)blockIdx.y * S;
    float av[8], bv[8], m[8], is[8], r0[8], r1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = cb * 8 + k;
        av[k] = __ldg(a + c); bv[k] = __ldg(b + c); m[k] = __ldg(mean + c); is[k] = __ldg(invstd + c);
        r0[k] = 0.f; r1[k] = 0.f;
    }
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        float v[8], g[8];
        V8<T>::load(z + (base + s) * 8, v);
        V8<T>::load(gy + (base + s) * 8, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float gg = (relu && !(v[k] * av[k] + bv[k] > 0.f)) ? 0.f : g[k];
            r0[k] += gg; r1[k] += gg * ((v[k] - m[k]) * is[k]);
        }
    }
    bn_block_sums(r0, r1, red + cb * 8, red + C + cb * 8);
}

// batch statistics: gz = a (g - red0/M - xhat red1/M);  frozen statistics (eval-mode fine-tuning): gz = a g.
// Block 0 also hands out grad_gamma = red1, grad_beta = red0.
template <typename T>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_t_kernel(const T* __restrict__ z, const T* __restrict__ gy, const float* __restrict__ a, const float* __restrict__ b,
                          const float* __restrict__ mean, const float* __restrict__ invstd, const double* __restrict__ red,
                          T* __restrict__ gz, float* __restrict__ ggamma, float* __restrict__ gbeta, int C, int64_t S, int64_t total,
                          double inv_m, int relu, int frozen) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && (int)threadIdx.x < C) {
        if (ggamma) ggamma[threadIdx.x] = (float)red[C + threadIdx.x];
        if (gbeta) gbeta[threadIdx.x] = (float)red[threadIdx.x];
    }
    if (i >= total) return;
    const int cb = (int)((i / S) % (C / 8));
    float v[8], g[8];
    V8<T>::load(z + i * 8, v);
    V8<T>::load(gy + i * 8, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = cb * 8 + k;
        const float av = __ldg(a + c);
        const float gg = (relu && !(v[k] * av + __ldg(b + c) > 0.f)) ? 0.f : g[k];
        if (frozen) { v[k] = av * gg; continue; }
        const float xh = (v[k] - __ldg(mean + c)) * __ldg(invstd + c);
        v[k] = av * (gg - (float)(red[c] * inv_m) - xh * (float)(red[C + c] * inv_m));
    }
    V8<T>::store(gz + i * 8, v);
}

// plain fp32 [n] -> C8 block [n][8] with the value in channel 0 (the gradient of the single-channel `prob` output)
template <typename T>
__global__ void lift_c1_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v[8] = {__ldg(src + i), 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    V8<T>::store(dst + i * 8, v);
}

// blocks along s of a reduction pass: >= 8 voxels per thread, and about 8 resident blocks per SM over the whole grid
static unsigned bn_reduce_blocks(int64_t S, int rows) {
    int64_t bx = (S + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = (148 * 8 + rows - 1) / rows;
    if (bx > cap) bx = cap;
    return (unsigned)(bx < 1 ? 1 : bx);
}

static int check_bn_t(const char* who, int B, int C, int64_t S) {
    MVS_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && C <= 256 && S > 0, MVS_E_SHAPE, "%s: bad dims (C must be a multiple of 8, <= 256)", who);
    MVS_REQUIRE((int64_t)B * (C / 8) <= 65535, MVS_E_SHAPE, "%s: B*C/8 too large for the launch grid", who);
    return MVS_OK;
}

extern "C" int mvs_bn_stats_t(const void* z, int dtype, double* sums, int B, int C, int64_t S, void* stream) {
    MVS_REQUIRE(z && sums, MVS_E_ARG, "mvs_bn_stats_t: null pointer");
    int rc = check_bn_t("mvs_bn_stats_t", B, C, S);
    if (rc) return rc;
    const unsigned bx = bn_reduce_blocks(S, B * (C / 8));
    MVS_DISPATCH_DTYPE(dtype, T, MVS_LAUNCH(bn_stats_t_kernel<T>, dim3(bx, (unsigned)(B * (C / 8))), dim3(256), stream, (const T*)z, sums, C, S));
    return MVS_CHECK_LAUNCH("mvs_bn_stats_t");
}

extern "C" int mvs_bn_finalize(const double* sums, const float* gamma, const float* beta, float eps, float momentum, double count,
                               float* a, float* b, float* mean, float* invstd, float* running_mean, float* running_var, int C,
                               void* stream) {
    MVS_REQUIRE(sums && gamma && beta && a && b && mean && invstd, MVS_E_ARG, "mvs_bn_finalize: null pointer");
    MVS_REQUIRE(C > 0 && count >= 1.0, MVS_E_SHAPE, "mvs_bn_finalize: bad dims");
    MVS_LAUNCH(bn_finalize_kernel, dim3(mvs_cdiv(C, 64)), dim3(64), stream, sums, gamma, beta, eps, momentum, count, a, b, mean, invstd,
               running_mean, running_var, C);
    return MVS_CHECK_LAUNCH("mvs_bn_finalize");
}

extern "C" int mvs_bn_act_fwd_t(const void* z, const float* a, const float* b, const void* skip, void* y, int dtype, int B, int C,
                                int64_t S, int relu, void* stream) {
    MVS_REQUIRE(z && a && b && y, MVS_E_ARG, "mvs_bn_act_fwd_t: null pointer");
    int rc = check_bn_t("mvs_bn_act_fwd_t", B, C, S);
    if (rc) return rc;
    const int64_t total = (int64_t)B * (C / 8) * S;
    MVS_DISPATCH_DTYPE(dtype, T, MVS_LAUNCH(bn_act_fwd_t_kernel<T>, dim3(mvs_cdiv(total, 256)), dim3(256), stream, (const T*)z, a, b,
                                            (const T*)skip, (T*)y, C, S, total, relu));
    return MVS_CHECK_LAUNCH("mvs_bn_act_fwd_t");
}

extern "C" int mvs_bn_act_bwd_reduce_t(const void* z, const void* grad_y, const float* a, const float* b, const float* mean,
                                       const float* invstd, double* red, int dtype, int B, int C, int64_t S, int relu, void* stream) {
    MVS_REQUIRE(z && grad_y && a && b && mean && invstd && red, MVS_E_ARG, "mvs_bn_act_bwd_reduce_t: null pointer");
    int rc = check_bn_t("mvs_bn_act_bwd_reduce_t", B, C, S);
    if (rc) return rc;
    const unsigned bx = bn_reduce_blocks(S, B * (C / 8));
    MVS_DISPATCH_DTYPE(dtype, T, MVS_LAUNCH(bn_act_bwd_reduce_t_kernel<T>, dim3(bx, (unsigned)(B * (C / 8))), dim3(256), stream, (const T*)z,
                                            (const T*)grad_y, a, b, mean, invstd, red, C, S, relu));
    return MVS_CHECK_LAUNCH("mvs_bn_act_bwd_reduce_t");
}

extern "C" int mvs_bn_act_bwd_apply_t(const void* z, const void* grad_y, const float* a, const float* b, const float* mean,
                                      const float* invstd, const double* red, void* grad_z, float* grad_gamma, float* grad_beta,
                                      int dtype, int B, int C, int64_t S, int relu, int frozen, void* stream) {
    MVS_REQUIRE(z && grad_y && a && b && mean && invstd && red && grad_z, MVS_E_ARG, "mvs_bn_act_bwd_apply_t: null pointer");
    int rc = check_bn_t("mvs_bn_act_bwd_apply_t", B, C, S);
    if (rc) return rc;
    const int64_t total = (int64_t)B * (C / 8) * S;
    const double inv_m = 1.0 / ((double)B * (double)S);
    MVS_DISPATCH_DTYPE(dtype, T, MVS_LAUNCH(bn_act_bwd_apply_t_kernel<T>, dim3(mvs_cdiv(total, 256)), dim3(256), stream, (const T*)z,
                                            (const T*)grad_y, a, b, mean, invstd, red, (T*)grad_z, grad_gamma, grad_beta, C, S, total,
                                            inv_m, relu, frozen));
    return MVS_CHECK_LAUNCH("mvs_bn_act_bwd_apply_t");
}

extern "C" int mvs_lift_c1(const float* src, void* dst, int dtype, int64_t n, void* stream) {
    MVS_REQUIRE(src && dst, MVS_E_ARG, "mvs_lift_c1: null pointer");
    MVS_REQUIRE(n > 0, MVS_E_SHAPE, "mvs_lift_c1: empty tensor");
    MVS_DISPATCH_DTYPE(dtype, T, MVS_LAUNCH(lift_c1_kernel<T>, dim3(mvs_cdiv(n, 256)), dim3(256), stream, src, (T*)dst, n));
    return MVS_CHECK_LAUNCH("mvs_lift_c1");
}

#ifndef MVS_CPU_EMU
#include <cuda.h>
#include <string.h>
// ------------------------------------------------------------------------------------------------ weight gradient on tensor cores
// grad[pc][qc][tap] = sum_b sum_p P[b][p][pc] Q[b][p * stride - 1 + k][qc]            (the formulation of conv3d_wgrad_kernel)
//   Conv3d          : P = gz (output grid, Cout), Q = x  (input grid, Cin)   -> grad_w [Cout][Cin][27]
//   ConvTranspose3d : P = x  (input grid, Cin),   Q = gz (output grid, Cout) -> grad_w [Cin][Cout][27]
// As MMAs:  D[m][n] += A[m][k] B[k][n] with k = 16 consecutive positions of a P tile, m = 16 channels of one tensor, n = 8 channels
// of the other, one accumulator tile per filter tap.  Both tensors are staged per tile by TMA (zero fill outside the volume = the
// convolution's padding and the ragged tile edges) as [channel block][position][8] -- 16-byte rows, the C8 layout itself -- so
// either can be the M side: ldmatrix.trans turns 8 position rows x 8 channels into the fragment of the channel-major operand, and
// the row ADDRESSES carry the tap shift and the stride (Q row = (hh s + kh) TWq + (ww s + kw)).
// A CTA owns the 9 (kh, kw) taps of one kd (or, when 16 x MT x NB accumulators would not fit the registers, the 3 kw taps of one
// (kd, kh)) and walks P tiles of TH x 32 positions.  WARP w OWNS TAP w (times a slice of the m-tiles): every warp has the same
// work, the shifted operand is fetched once per k-step and warp, and nothing but ldmatrix + mma is left in the inner loop
// (first version: 58 instructions per MMA from index arithmetic and a per-unit dispatch; ncu: tensor pipe 22 %).
constexpr int kWgTW = 32;

struct WgParams {
    float* grad;
    int PCB, QCB, PCreal, QCreal;          // channel blocks of P / Q; real channel counts (padding blocks are not written)
    int Dp, Dq, stride;
    int a_is_p;                            // M side: 1 = P, 0 = Q
    int MT;                                // m-tiles (16 channels) of the M side
    int tpc, msplit;                       // taps per CTA (9 / 3); warps = tpc * msplit, each owning MT / msplit m-tiles of one tap
    int TH, nht, nwt, ntiles, B;           // P tile rows; tiles per plane; tiles in total (B * Dp * nht * nwt)
    int THq, TWq, q_rows;                  // staged Q tile (with halo)
    int kd0;                               // first kd plane of taps this launch covers (0; 1 for planar = 2-D layers: kd = 1 only)
    uint32_t p_bytes, q_bytes, buf_stride; // bytes one TMA brings; distance between the two tile buffers (128-byte aligned)
};
struct WgMaps { CUtensorMap p, q; };

__device__ __forceinline__ uint32_t wg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T> __device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]);
template <> __device__ __forceinline__ void mma_16816<__half>(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <> __device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// MTW = m-tiles per warp, NB = n-blocks: MTW * NB * 4 accumulator registers per thread.
template <typename T, int MTW, int NB>
__global__ void __launch_bounds__(288, (MTW * NB >= 16) ? 1 : 2)
conv3d_wgrad_mma_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgParams p) {
    extern __shared__ __align__(128) uint8_t wg_smem[];
    __shared__ uint64_t bar[2];
    uint8_t* sP = wg_smem;                                              // per buffer: [PCB][TH * TW][8] then [QCB][THq * TWq][8]
    uint8_t* sQ = wg_smem + p.p_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.y + (p.tpc == 9 ? p.kd0 : 3 * p.kd0);    // tap group: kd (tpc = 9) or kd * 3 + kh (tpc = 3)
    const int kd = p.tpc == 9 ? group : group / 3;
    const int tapl = warp % p.tpc, mslice = warp / p.tpc;               // this warp's tap inside the group, and its slice of the m-tiles
    const int kh = p.tpc == 9 ? tapl / 3 : group % 3, kw = tapl % 3;
    const int mt0 = mslice * MTW;

    float acc[MTW][NB][4];
#pragma unroll
    for (int j = 0; j < MTW; ++j)
#pragma unroll
        for (int n = 0; n < NB; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][n][e] = 0.f;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(wg_smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(wg_smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.p) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.q) : "memory");
    }
    __syncthreads();

    const uint32_t sP32 = wg_smem_u32(sP), sQ32 = wg_smem_u32(sQ);
    const uint32_t p_blk = (uint32_t)(p.TH * kWgTW) * 16u, q_blk = (uint32_t)p.q_rows * 16u;      // bytes per channel block
    const int lj = lane >> 3, li = lane & 7;          // ldmatrix: this lane supplies row li of matrix lj
    const bool a_is_p = p.a_is_p != 0;
    const int MCB = a_is_p ? p.PCB : p.QCB, NCB = a_is_p ? p.QCB : p.PCB;              // channel blocks of the M / N side
    // Lane-constant parts of the row addresses.  P rows: position q = hh * TW + ww.  Q rows: (hh s + kh) TWq + ww s + kw.
    //   A matrices (x4): j = 0: (block 2 mt, k 0-7), 1: (2 mt + 1, k 0-7), 2: (2 mt, k 8-15), 3: (2 mt + 1, k 8-15)
    //   B matrices (x4): j = 0: (block n, k 0-7), 1: (n, k 8-15), 2: (n + 1, k 0-7), 3: (n + 1, k 8-15)
    const uint32_t tap_shift = (uint32_t)(kh * p.TWq + kw) * 16u;
    uint32_t a_lane[MTW], b_lane[(NB + 1) / 2];
    bool half_m[MTW];
#pragma unroll
    for (int j = 0; j < MTW; ++j) {
        const int cbk = min(2 * (mt0 + j) + (lj & 1), MCB - 1), kk = (lj >> 1) * 8 + li;
        a_lane[j] = a_is_p ? sP32 + (uint32_t)cbk * p_blk + (uint32_t)kk * 16u
                           : sQ32 + (uint32_t)cbk * q_blk + tap_shift + (uint32_t)(kk * p.stride) * 16u;
        half_m[j] = MCB * 8 < (mt0 + j + 1) * 16;     // an 8-channel M side: rows 8..15 of the tile do not exist
    }
#pragma unroll
    for (int n = 0; n < NB; n += 2) {
        const int nbc = min(n + (lj >> 1), NCB - 1), kk = (lj & 1) * 8 + li;
        b_lane[n / 2] = a_is_p ? sQ32 + (uint32_t)nbc * q_blk + tap_shift + (uint32_t)(kk * p.stride) * 16u
                               : sP32 + (uint32_t)nbc * p_blk + (uint32_t)kk * 16u;
    }
    const uint32_t p_row_h = kWgTW * 16u, q_row_h = (uint32_t)(p.stride * p.TWq) * 16u;          // bytes per tile row hh
    const uint32_t p_k16 = 16u * 16u, q_k16 = (uint32_t)(16 * p.stride) * 16u;                   // bytes per 16 positions
    const uint32_t a_row_h = a_is_p ? p_row_h : q_row_h, a_k16 = a_is_p ? p_k16 : q_k16;
    const uint32_t b_row_h = a_is_p ? q_row_h : p_row_h, b_k16 = a_is_p ? q_k16 : p_k16;

    // Two tile buffers: the TMA of the NEXT tile of this CTA is in flight while the warps run the MMAs of the current one.
    const uint32_t buf_bytes = p.p_bytes + p.q_bytes, buf_stride = p.buf_stride;
    auto tile_coords = [&](int t, int& b, int& dp, int& h0, int& w0, int& qd) {
        int tt = t;
        const int wt = tt % p.nwt; tt /= p.nwt;
        const int ht = tt % p.nht; tt /= p.nht;
        dp = tt % p.Dp; b = tt / p.Dp;
        h0 = ht * p.TH; w0 = wt * kWgTW;
        qd = dp * p.stride - 1 + kd;
        return qd >= 0 && qd < p.Dq;                                    // false: this tap plane lies in the zero padding (block-uniform)
    };
    auto next_valid = [&](int t) {                                      // first tile >= t of this CTA whose tap plane exists
        int b, dp, h0, w0, qd;
        while (t < p.ntiles && !tile_coords(t, b, dp, h0, w0, qd)) t += gridDim.x;
        return t;
    };
    auto issue = [&](int t, int buf) {                                  // thread 0: both TMA loads of tile t into buffer buf
        int b, dp, h0, w0, qd;
        tile_coords(t, b, dp, h0, w0, qd);
        const uint32_t mb = wg_smem_u32(&bar[buf]), dst = sP32 + (uint32_t)buf * buf_stride;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(buf_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(dst), "l"(&maps.p), "r"(mb), "r"(0), "r"(w0), "r"(h0), "r"(dp), "r"(b * p.PCB) : "memory");
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(dst + p.p_bytes), "l"(&maps.q), "r"(mb), "r"(0), "r"(w0 * p.stride - 1), "r"(h0 * p.stride - 1), "r"(qd), "r"(b * p.QCB) : "memory");
    };
    int t = next_valid(blockIdx.x);
    if (threadIdx.x == 0 && t < p.ntiles) issue(t, 0);
    uint32_t phase[2] = {0u, 0u};
    for (int it = 0; t < p.ntiles; ++it) {
        const int buf = it & 1;
        const int tn = next_valid(t + gridDim.x);
        // buffer buf ^ 1 was last read in iteration it - 1; the barrier at the end of that iteration makes it free
        if (threadIdx.x == 0 && tn < p.ntiles) issue(tn, buf ^ 1);
        asm volatile("{\n\t.reg .pred p;\n\tWG_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WG_DONE_%=;\n\tbra WG_WAIT_%=;\n\tWG_DONE_%=:\n\t}"
                     ::"r"(wg_smem_u32(&bar[buf])), "r"(phase[buf]) : "memory");
        phase[buf] ^= 1u;
        const uint32_t bufo = (uint32_t)buf * buf_stride;
        // ---- TH rows x 2 k-steps of 16 positions
#pragma unroll 2
        for (int hh = 0; hh < p.TH; ++hh) {
#pragma unroll
            for (int kk = 0; kk < kWgTW / 16; ++kk) {
                const uint32_t ao = bufo + (uint32_t)hh * a_row_h + (uint32_t)kk * a_k16, bo = bufo + (uint32_t)hh * b_row_h + (uint32_t)kk * b_k16;
                uint32_t bfr[NB][2];
#pragma unroll
                for (int n = 0; n < NB; n += 2) {
                    uint32_t f[4];
                    ldmatrix_x4_trans(b_lane[n / 2] + bo, f);
                    bfr[n][0] = f[0]; bfr[n][1] = f[1];
                    if (n + 1 < NB) { bfr[n + 1 < NB ? n + 1 : n][0] = f[2]; bfr[n + 1 < NB ? n + 1 : n][1] = f[3]; }
                }
#pragma unroll
                for (int j = 0; j < MTW; ++j) {
                    uint32_t afrag[4];
                    ldmatrix_x4_trans(a_lane[j] + ao, afrag);
                    if (half_m[j]) { afrag[1] = 0u; afrag[3] = 0u; }
#pragma unroll
                    for (int n = 0; n < NB; ++n) mma_16816<T>(acc[j][n], afrag, bfr[n]);
                }
            }
        }
        __syncthreads();                                                // every warp is done with buffer buf: it may be refilled next iteration
        t = tn;
    }
    // ---- write-out: C fragment (m = lane / 4 (+ 8), n = 2 (lane % 4) + {0, 1}) -> grad[pc][qc][tap], torch layout
    const int QC = p.QCreal, PC = p.PCreal;
    const int tap = (kd * 3 + kh) * 3 + kw;
#pragma unroll
    for (int j = 0; j < MTW; ++j)
#pragma unroll
        for (int n = 0; n < NB; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = (mt0 + j) * 16 + (lane >> 2) + ((e >> 1) ? 8 : 0), nn = n * 8 + 2 * (lane & 3) + (e & 1);
                const int pc = a_is_p ? m : nn, qc = a_is_p ? nn : m;
                if (pc < PC && qc < QC && acc[j][n][e] != 0.f) atomicAdd(p.grad + ((int64_t)pc * QC + qc) * 27 + tap, acc[j][n][e]);
            }
}

typedef CUresult (*WgEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WgEncodeTiledFn wg_encode_tiled() {
    static WgEncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<WgEncodeTiledFn>(p);
    }
    return fn;
}

// C8 volume [B][CB][D][H][W][8] as a 5-D tensor {8, W, H, D, B * CB}; box {8, bw, bh, 1, CB}: all channel blocks of one tile
static int wg_make_map(CUtensorMap* m, const void* base, int is_bf16, int B, int CB, int D, int H, int W, int bw, int bh) {
    WgEncodeTiledFn enc = wg_encode_tiled();
    MVS_REQUIRE(enc, MVS_E_LAUNCH, "mvs_conv3d_wgrad_mma: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B * CB};
    const cuuint64_t gstr[4] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
    const cuuint32_t box[5] = {8, (cuuint32_t)bw, (cuuint32_t)bh, 1, (cuuint32_t)CB}, estr[5] = {1, 1, 1, 1, 1};
    const CUresult cr = enc(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), gdim, gstr, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVS_REQUIRE(cr == CUDA_SUCCESS, MVS_E_LAUNCH, "mvs_conv3d_wgrad_mma: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return MVS_OK;
}

static int wgrad_mma_impl(const mvs_conv3d_desc* d, const void* x, const void* grad_z, float* grad_w, int cout_real, int planar,
                          void* stream) {
    MVS_REQUIRE(d && x && grad_z && grad_w, MVS_E_ARG, "mvs_conv3d_wgrad_mma: null pointer");
    MVS_REQUIRE(d->dtype_in == MVS_F16 || d->dtype_in == MVS_BF16, MVS_E_UNSUPPORTED, "mvs_conv3d_wgrad_mma: 16-bit storage only (fp32 takes mvs_conv3d_bwd_weight)");
    MVS_REQUIRE(d->Cin % 8 == 0 && d->Cout % 8 == 0 && d->Cin <= 64 && d->Cout <= 64, MVS_E_SHAPE,
                "mvs_conv3d_wgrad_mma: Cin=%d / Cout=%d must be multiples of 8, <= 64 (lift a single-channel gradient with mvs_lift_c1)", d->Cin, d->Cout);
    MVS_REQUIRE(d->stride == 1 || d->stride == 2, MVS_E_SHAPE, "mvs_conv3d_wgrad_mma: stride must be 1 or 2");
    MVS_REQUIRE(cout_real >= 1 && cout_real <= d->Cout, MVS_E_ARG, "mvs_conv3d_wgrad_mma: bad cout_real");
    WgParams p;
    memset(&p, 0, sizeof(p));
    const bool tr = d->transposed != 0;
    const void* P = tr ? x : grad_z;
    const void* Q = tr ? grad_z : x;
    p.grad = grad_w; p.B = d->B;
    const int PC = tr ? d->Cin : d->Cout, QC = tr ? d->Cout : d->Cin;
    p.PCB = PC / 8; p.QCB = QC / 8;
    p.PCreal = tr ? d->Cin : cout_real; p.QCreal = tr ? cout_real : d->Cin;
    const int Hp = tr ? d->Hin : d->Hout, Wp = tr ? d->Win : d->Wout, Hq = tr ? d->Hout : d->Hin, Wq = tr ? d->Wout : d->Win;
    p.Dp = tr ? d->Din : d->Dout; p.Dq = tr ? d->Dout : d->Din;
    p.stride = d->stride;
    // M side: a tensor whose channel count is a multiple of 16 (the wider one if both are); else P with its upper half empty
    if (PC % 16 == 0 && (QC % 16 != 0 || PC >= QC)) p.a_is_p = 1;
    else if (QC % 16 == 0) p.a_is_p = 0;
    else p.a_is_p = 1;
    const int MC = p.a_is_p ? PC : QC, NC = p.a_is_p ? QC : PC;
    p.MT = (MC + 15) / 16;
    const int NB = NC / 8;
    MVS_REQUIRE(p.MT == 1 || p.MT == 2 || p.MT == 4, MVS_E_UNSUPPORTED, "mvs_conv3d_wgrad_mma: %d channels on the M side (16, 32 or 64 supported)", MC);
    MVS_REQUIRE(NB == 1 || NB == 2 || NB == 4 || NB == 8, MVS_E_UNSUPPORTED, "mvs_conv3d_wgrad_mma: %d channels on the N side (8, 16, 32 or 64 supported)", NC);
    // one warp per tap; when a warp's MT x NB accumulator tiles exceed 64 registers the CTA takes 3 taps and two warps share a tap
    if (p.MT * NB <= 16) { p.tpc = 9; p.msplit = 1; } else { p.tpc = 3; p.msplit = 2; }
    const int mtw = p.MT / p.msplit;
    const int nwarps = p.tpc * p.msplit;
    // tile rows: 16 when both tiles stay small (thin layers: amortises the per-tile TMA round trip), else 8
    p.TH = 16;
    for (;;) {
        p.THq = p.TH * p.stride + 2; p.TWq = kWgTW * p.stride + 2; p.q_rows = p.THq * p.TWq;
        p.p_bytes = (uint32_t)p.PCB * p.TH * kWgTW * 16u; p.q_bytes = (uint32_t)p.QCB * p.q_rows * 16u;
        if (p.TH == 8 || p.p_bytes + p.q_bytes <= 48 * 1024) break;
        p.TH = 8;
    }
    p.buf_stride = (p.p_bytes + p.q_bytes + 127u) & ~127u;              // TMA destinations are 128-byte aligned
    const size_t smem = 2 * (size_t)p.buf_stride;                       // two tile buffers
    MVS_REQUIRE(smem <= 220 * 1024, MVS_E_UNSUPPORTED, "mvs_conv3d_wgrad_mma: tiles of %d + %d channels at stride %d need %zu bytes of shared memory", PC, QC, p.stride, smem);
    MVS_REQUIRE(p.TWq <= 256 && p.THq <= 256, MVS_E_UNSUPPORTED, "mvs_conv3d_wgrad_mma: tile exceeds the TMA box limit");
    p.nht = (Hp + p.TH - 1) / p.TH; p.nwt = (Wp + kWgTW - 1) / kWgTW;
    p.ntiles = p.B * p.Dp * p.nht * p.nwt;
    WgMaps maps;
    memset(&maps, 0, sizeof(maps));
    const int is_bf16 = d->dtype_in == MVS_BF16;
    int rc = wg_make_map(&maps.p, P, is_bf16, p.B, p.PCB, p.Dp, Hp, Wp, kWgTW, p.TH);
    if (rc) return rc;
    rc = wg_make_map(&maps.q, Q, is_bf16, p.B, p.QCB, p.Dq, Hq, Wq, p.TWq, p.THq);
    if (rc) return rc;
    // planar: the depth axis indexes independent images (2-D layers run as zero-kd 3-D layers): only the kd = 1 taps exist
    p.kd0 = planar ? 1 : 0;
    const int ngroups = (planar ? 9 : 27) / p.tpc;
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;                      // resident CTAs (shared memory; registers allow 2 at most)
    int gx = (148 * per_sm + ngroups - 1) / ngroups;
    if (gx > p.ntiles) gx = p.ntiles;
    const dim3 grid((unsigned)gx, (unsigned)ngroups);
    cudaStream_t st = (cudaStream_t)stream;
#define MVS_WG_LAUNCH(T, M, N) do { cudaFuncSetAttribute(conv3d_wgrad_mma_kernel<T, M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                                    conv3d_wgrad_mma_kernel<T, M, N><<<grid, nwarps * 32, smem, st>>>(maps, p); } while (0)
#define MVS_WG_BY_NB(T, M) do { if (NB == 1) MVS_WG_LAUNCH(T, M, 1); else if (NB == 2) MVS_WG_LAUNCH(T, M, 2); \
                                else if (NB == 4) { if (M * 4 <= 16) MVS_WG_LAUNCH(T, M, 4); } else { if (M * 8 <= 16) MVS_WG_LAUNCH(T, (M * 8 <= 16 ? M : 1), 8); } } while (0)
#define MVS_WG_BY_SHAPE(T) do { if (mtw == 1) MVS_WG_BY_NB(T, 1); else if (mtw == 2) MVS_WG_BY_NB(T, 2); else MVS_WG_BY_NB(T, 4); } while (0)
    if (is_bf16) MVS_WG_BY_SHAPE(__nv_bfloat16); else MVS_WG_BY_SHAPE(__half);
#undef MVS_WG_BY_SHAPE
#undef MVS_WG_BY_NB
#undef MVS_WG_LAUNCH
    return MVS_CHECK_LAUNCH("mvs_conv3d_wgrad_mma");
}

extern "C" int mvs_conv3d_wgrad_mma(const mvs_conv3d_desc* d, const void* x, const void* grad_z, float* grad_w, int cout_real,
                                    void* stream) {
    return wgrad_mma_impl(d, x, grad_z, grad_w, cout_real, 0, stream);
}

extern "C" int mvs_conv2d_wgrad_mma(const mvs_conv3d_desc* d, const void* x, const void* grad_z, float* grad_w, int cout_real,
                                    void* stream) {
    return wgrad_mma_impl(d, x, grad_z, grad_w, cout_real, 1, stream);
}
#else
extern "C" int mvs_conv2d_wgrad_mma(const mvs_conv3d_desc*, const void*, const void*, float*, int, void*) {
    return mvs_set_error(MVS_E_UNSUPPORTED, "mvs_conv2d_wgrad_mma: tensor-core kernels do not exist in the emulation build");
}
extern "C" int mvs_conv3d_wgrad_mma(const mvs_conv3d_desc*, const void*, const void*, float*, int, void*) {
    return mvs_set_error(MVS_E_UNSUPPORTED, "mvs_conv3d_wgrad_mma: tensor-core kernels do not exist in the emulation build");
}
#endif
