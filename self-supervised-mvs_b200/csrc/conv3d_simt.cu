// conv3d_simt.cu — fp32-accumulate direct 3x3x3 convolution on CUDA cores over the C8 layout.
//
// This is the PRECISE path of the regularisation network (algo = 1): every product is formed in fp32
// exactly once, so it is the kernel the parity tests hold against the oracle to ~1e-6, the reference
// the tcgen05 implicit-GEMM path (conv3d_tc.cu, algo = 2) is validated against on the GPU, and the
// kernel behind training-mode gradients.  It also carries the batch-norm helpers of training mode.
// Reference: ConvBnReLU3D jdacs/models/module.py:35-42; CostRegNet jdacs/models/mvsnet.py:37-74 and
// jdacs-ms/models/network.py:44-74.
#include "mvs_rt.h"

// ---------------------------------------------------------------------------------------------- weights
// torch Conv3d weight [Cout][Cin][27] / ConvTranspose3d weight [Cin][Cout][27]  ->  G[27][Cin][CoutPad]
__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ g, int Cin, int Cout, int CoutPad,
                                   int transposed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 27 * Cin * CoutPad
    if (i >= 27 * Cin * CoutPad) return;
    const int co = i % CoutPad;
    const int ci = (i / CoutPad) % Cin;
    const int tap = i / (CoutPad * Cin);
    float v = 0.f;
    if (co < Cout) v = transposed ? w[((int64_t)ci * Cout + co) * 27 + tap] : w[((int64_t)co * Cin + ci) * 27 + tap];
    g[i] = v;
}

extern "C" int mvs_pack_conv3d_weight(const float* w, float* g, int Cin, int Cout, int transposed, void* stream) {
    MVS_REQUIRE(w && g, MVS_E_ARG, "mvs_pack_conv3d_weight: null pointer");
    MVS_REQUIRE(Cin > 0 && Cout > 0, MVS_E_SHAPE, "mvs_pack_conv3d_weight: bad channel counts");
    const int CoutPad = (Cout + 7) / 8 * 8;
    MVS_LAUNCH(pack_weight_kernel, dim3(mvs_cdiv(27 * Cin * CoutPad, 256)), dim3(256), stream, w, g, Cin, Cout, CoutPad, transposed);
    return MVS_CHECK_LAUNCH("mvs_pack_conv3d_weight");
}

// ---------------------------------------------------------------------------------------------- forward
// Gather form: out[o] = sum_tap x[in(o, tap)] . G[tap]
//   Conv3d          : in = o * stride - 1 + k
//   ConvTranspose3d : in = (o + 1 - k) / stride   when divisible (output_padding only enlarges the output grid)
// thread = (output voxel, block of 8 output channels); lanes along w.
template <bool TRANSPOSED>
__device__ __forceinline__ bool in_coord(int o, int k, int stride, int n_in, int& i) {
    if (TRANSPOSED) {
        const int t = o + 1 - k;
        if (t < 0 || (t % stride) != 0) return false;
        i = t / stride;
    } else {
        i = o * stride - 1 + k;
    }
    return i >= 0 && i < n_in;
}

template <typename TI, typename TO, bool TRANSPOSED>
__global__ void __launch_bounds__(128)
conv3d_simt_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ g, const float* __restrict__ scale,
                       const float* __restrict__ shift, const TO* __restrict__ skip, TO* __restrict__ y,
                       mvs_conv3d_desc d) {
    const int64_t Vout = (int64_t)d.Dout * d.Hout * d.Wout;
    const int64_t Vin = (int64_t)d.Din * d.Hin * d.Win;
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= Vout) return;
    const int CoutPad = (d.Cout + 7) / 8 * 8, CoB = CoutPad / 8, CiB = d.Cin / 8;
    const int cob = blockIdx.y % CoB;
    const int b = blockIdx.y / CoB;
    const int ow = (int)(v % d.Wout), oh = (int)((v / d.Wout) % d.Hout), od = (int)(v / ((int64_t)d.Wout * d.Hout));

    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;

    for (int kd = 0; kd < 3; ++kd) {
        int id;
        if (!in_coord<TRANSPOSED>(od, kd, d.stride, d.Din, id)) continue;
        for (int kh = 0; kh < 3; ++kh) {
            int ih;
            if (!in_coord<TRANSPOSED>(oh, kh, d.stride, d.Hin, ih)) continue;
            for (int kw = 0; kw < 3; ++kw) {
                int iw;
                if (!in_coord<TRANSPOSED>(ow, kw, d.stride, d.Win, iw)) continue;
                const int tap = (kd * 3 + kh) * 3 + kw;
                const int64_t vin = ((int64_t)id * d.Hin + ih) * d.Win + iw;
                for (int cib = 0; cib < CiB; ++cib) {
                    float xv[8];
                    V8<TI>::load(x + (((int64_t)b * CiB + cib) * Vin + vin) * 8, xv);
                    const float* gw = g + ((int64_t)tap * d.Cin + cib * 8) * CoutPad + cob * 8;
#pragma unroll
                    for (int ci = 0; ci < 8; ++ci) {
                        float wv[8];
                        V8<float>::load(gw + (int64_t)ci * CoutPad, wv);
#pragma unroll
                        for (int co = 0; co < 8; ++co) acc[co] += xv[ci] * wv[co];
                    }
                }
            }
        }
    }
    // epilogue: affine (folded BN or bias), ReLU, skip add
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int co = cob * 8 + k;
        if (co < d.Cout) {
            if (scale) acc[k] *= __ldg(scale + co);
            if (shift) acc[k] += __ldg(shift + co);
            if (d.relu) acc[k] = fmaxf(acc[k], 0.f);
        }
    }
    if (d.Cout == 1) {  // plain fp32 [B][D][H][W]
        float* yp = reinterpret_cast<float*>(y);
        float o = acc[0];
        if (skip) o += reinterpret_cast<const float*>(skip)[(int64_t)b * Vout + v];
        yp[(int64_t)b * Vout + v] = o;
    } else {
        const int64_t off = (((int64_t)b * CoB + cob) * Vout + v) * 8;
        if (skip) {
            float sv[8];
            V8<TO>::load(skip + off, sv);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += sv[k];
        }
        V8<TO>::store(y + off, acc);
    }
}

static int check_conv_desc(const mvs_conv3d_desc* d, const char* who) {
    MVS_REQUIRE(d, MVS_E_ARG, "%s: null descriptor", who);
    MVS_REQUIRE(d->B > 0 && d->Cin > 0 && d->Cout > 0, MVS_E_SHAPE, "%s: bad B/Cin/Cout", who);
    MVS_REQUIRE(d->Cin % 8 == 0, MVS_E_SHAPE, "%s: Cin=%d must be a multiple of 8 (C8 layout)", who, d->Cin);
    MVS_REQUIRE(d->Cout == 1 || d->Cout % 8 == 0, MVS_E_SHAPE, "%s: Cout=%d must be 1 or a multiple of 8", who, d->Cout);
    MVS_REQUIRE(d->stride == 1 || d->stride == 2, MVS_E_SHAPE, "%s: stride must be 1 or 2", who);
    MVS_REQUIRE(d->Din > 0 && d->Hin > 0 && d->Win > 0 && d->Dout > 0 && d->Hout > 0 && d->Wout > 0, MVS_E_SHAPE, "%s: empty volume", who);
    const int in[3] = {d->Din, d->Hin, d->Win}, out[3] = {d->Dout, d->Hout, d->Wout};
    for (int a = 0; a < 3; ++a) {
        if (d->stride == 1) {
            MVS_REQUIRE(out[a] == in[a], MVS_E_SHAPE, "%s: stride-1 output extent %d != input extent %d", who, out[a], in[a]);
        } else if (!d->transposed) {
            MVS_REQUIRE(in[a] % 2 == 0 && out[a] == in[a] / 2, MVS_E_SHAPE,
                        "%s: stride-2 Conv3d needs even input extents (got %d -> %d); the reference U-Net has the same constraint", who, in[a], out[a]);
        } else {
            MVS_REQUIRE(out[a] == in[a] * 2, MVS_E_SHAPE, "%s: stride-2 ConvTranspose3d output extent %d != 2 * %d", who, out[a], in[a]);
        }
    }
    return MVS_OK;
}

#ifndef MVS_CPU_EMU
int mvs_conv3d_fwd_tc(const mvs_conv3d_desc* d, const void* x, const float* g, const float* scale, const float* shift,
                      const void* skip, void* y, void* ws, void* stream);  // conv3d_tc.cu
int mvs_conv3d_tc_supported(const mvs_conv3d_desc* d);
int64_t mvs_conv3d_tc_workspace_bytes(const mvs_conv3d_desc* d);
int64_t mvs_conv2d_tc_workspace_bytes(const mvs_conv2d_desc* d);
int mvs_conv2d_fwd_tc(const mvs_conv2d_desc* d, const void* x, const float* g, const float* scale, const float* shift, void* y,
                      void* ws, void* stream);
#endif

// 2-D feature-extractor layers: tcgen05 only (eval mode, 16-bit storage); training keeps the library 2-D convolutions.
extern "C" int64_t mvs_conv2d_workspace_bytes(const mvs_conv2d_desc* d) {
#ifndef MVS_CPU_EMU
    if (d) return mvs_conv2d_tc_workspace_bytes(d);
#endif
    (void)d;
    return 0;
}

extern "C" int mvs_conv2d_fwd(const mvs_conv2d_desc* d, const void* x, const float* g, const float* scale, const float* shift,
                              void* y, void* ws, void* stream) {
    MVS_REQUIRE(d && x && g && y, MVS_E_ARG, "mvs_conv2d_fwd: null pointer");
    MVS_REQUIRE(d->M > 0 && d->Hin > 0 && d->Win > 0, MVS_E_SHAPE, "mvs_conv2d_fwd: bad dims");
    MVS_REQUIRE(d->stride == 1 ? (d->Hout == d->Hin && d->Wout == d->Win) : (d->Hout * 2 == d->Hin && d->Wout * 2 == d->Win), MVS_E_SHAPE,
                "mvs_conv2d_fwd: output extent %dx%d does not match input %dx%d at stride %d", d->Hout, d->Wout, d->Hin, d->Win, d->stride);
#ifndef MVS_CPU_EMU
    return mvs_conv2d_fwd_tc(d, x, g, scale, shift, y, ws, stream);
#else
    (void)scale; (void)shift; (void)ws; (void)stream;
    return mvs_set_error(MVS_E_UNSUPPORTED, "mvs_conv2d_fwd: the tcgen05 path does not exist in the emulation build");
#endif
}

extern "C" int64_t mvs_conv3d_workspace_bytes(const mvs_conv3d_desc* d) {
#ifndef MVS_CPU_EMU
    if (d && d->algo != 1) return mvs_conv3d_tc_workspace_bytes(d);
#endif
    (void)d;
    return 0;
}

extern "C" int mvs_conv3d_fwd(const mvs_conv3d_desc* d, const void* x, const float* g, const float* scale,
                              const float* shift, const void* skip, void* y, void* ws, void* stream) {
    int rc = check_conv_desc(d, "mvs_conv3d_fwd");
    if (rc) return rc;
    MVS_REQUIRE(x && g && y, MVS_E_ARG, "mvs_conv3d_fwd: null pointer");
    MVS_REQUIRE(d->algo >= 0 && d->algo <= 3, MVS_E_ARG, "mvs_conv3d_fwd: unknown algo %d", d->algo);
    MVS_REQUIRE(d->Cout != 1 || d->dtype_out == MVS_F32, MVS_E_ARG, "mvs_conv3d_fwd: single-channel output is plain fp32");
#ifndef MVS_CPU_EMU
    if (d->algo >= 2 || (d->algo == 0 && mvs_conv3d_tc_supported(d))) return mvs_conv3d_fwd_tc(d, x, g, scale, shift, skip, y, ws, stream);
#else
    MVS_REQUIRE(d->algo < 2, MVS_E_UNSUPPORTED, "mvs_conv3d_fwd: the tcgen05 path does not exist in the emulation build");
#endif
    const int64_t Vout = (int64_t)d->Dout * d->Hout * d->Wout;
    const int CoB = ((d->Cout + 7) / 8);
    MVS_REQUIRE((int64_t)d->B * CoB <= 65535, MVS_E_SHAPE, "mvs_conv3d_fwd: B*Cout/8 too large for the launch grid");
    const dim3 grid(mvs_cdiv(Vout, 128), (unsigned)(d->B * CoB));
    if (d->transposed) {
        MVS_DISPATCH_DTYPE(d->dtype_in, TI, MVS_DISPATCH_DTYPE(d->dtype_out, TO,
            MVS_LAUNCH((conv3d_simt_fwd_kernel<TI, TO, true>), grid, dim3(128), stream, (const TI*)x, g, scale, shift, (const TO*)skip, (TO*)y, *d)));
    } else {
        MVS_DISPATCH_DTYPE(d->dtype_in, TI, MVS_DISPATCH_DTYPE(d->dtype_out, TO,
            MVS_LAUNCH((conv3d_simt_fwd_kernel<TI, TO, false>), grid, dim3(128), stream, (const TI*)x, g, scale, shift, (const TO*)skip, (TO*)y, *d)));
    }
    return MVS_CHECK_LAUNCH("mvs_conv3d_fwd");
}

// ---------------------------------------------------------------------------------------------- weight gradient
// grad[pc][qc][k] = sum_b sum_p P[b][p][pc] * Q[b][p*stride - 1 + k][qc]
//   Conv3d          : P = grad_y (out grid, Cout), Q = x (in grid, Cin)   -> grad_w [Cout][Cin][27]
//   ConvTranspose3d : P = x (in grid, Cin),        Q = grad_y (out grid, Cout) -> grad_w [Cin][Cout][27]
// thread = (w lane of the P grid) x (tap, pc block, qc block) x (slab of (b,d,h) rows); 64 accumulators,
// warp-reduced, then one atomic per weight per warp.
template <typename TP, typename TQ>
__global__ void __launch_bounds__(128)
conv3d_wgrad_kernel(const TP* __restrict__ P, const TQ* __restrict__ Q, float* __restrict__ grad, int B, int PC, int QC,
                    int PCreal, int Dp, int Hp, int Wp, int Dq, int Hq, int Wq, int stride, int rows_per_slab) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int PCB = PC / 8, QCB = QC / 8;
    int yb = blockIdx.y;
    const int qcb = yb % QCB; yb /= QCB;
    const int pcb = yb % PCB; yb /= PCB;
    const int tap = yb;
    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    const int64_t Vp = (int64_t)Dp * Hp * Wp, Vq = (int64_t)Dq * Hq * Wq;
    const int64_t rows = (int64_t)B * Dp * Hp;
    const int64_t r0 = (int64_t)blockIdx.z * rows_per_slab;
    const int64_t r1 = (r0 + rows_per_slab < rows) ? r0 + rows_per_slab : rows;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int qw = w * stride - 1 + kw;
    const bool wok = (w < Wp) && qw >= 0 && qw < Wq;
    if (wok) {
        for (int64_t r = r0; r < r1; ++r) {
            const int h = (int)(r % Hp), dd = (int)((r / Hp) % Dp), b = (int)(r / ((int64_t)Hp * Dp));
            const int qd = dd * stride - 1 + kd, qh = h * stride - 1 + kh;
            if (qd < 0 || qd >= Dq || qh < 0 || qh >= Hq) continue;
            float pv[8], qv[8];
            V8<TP>::load(P + (((int64_t)b * PCB + pcb) * Vp + ((int64_t)dd * Hp + h) * Wp + w) * 8, pv);
            V8<TQ>::load(Q + (((int64_t)b * QCB + qcb) * Vq + ((int64_t)qd * Hq + qh) * Wq + qw) * 8, qv);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] += pv[i] * qv[j];
        }
    }
#ifndef MVS_CPU_EMU
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = acc[i][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[i][j] = v;
        }
    if ((threadIdx.x & 31) != 0) return;
#else
    if (!wok) return;
#endif
    // torch layout: grad[pc][qc][27]; PCreal trims the padding block of a single-channel P
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int pc = pcb * 8 + i;
        if (pc >= PCreal) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(grad + ((int64_t)pc * QC + qcb * 8 + j) * 27 + tap, acc[i][j]);
    }
}

extern "C" int mvs_conv3d_bwd_weight(const mvs_conv3d_desc* d, const void* x, const float* grad_y, float* grad_w,
                                     void* stream) {
    int rc = check_conv_desc(d, "mvs_conv3d_bwd_weight");
    if (rc) return rc;
    MVS_REQUIRE(x && grad_y && grad_w, MVS_E_ARG, "mvs_conv3d_bwd_weight: null pointer");
    MVS_REQUIRE(d->Cout % 8 == 0, MVS_E_SHAPE, "mvs_conv3d_bwd_weight: Cout=%d must be a multiple of 8 (pad a single-channel gradient to a C8 block)", d->Cout);
    // P grid is the coarser (or equal) grid: the conv output, or the transposed conv input
    const int Dp = d->transposed ? d->Din : d->Dout, Hp = d->transposed ? d->Hin : d->Hout, Wp = d->transposed ? d->Win : d->Wout;
    const int Dq = d->transposed ? d->Dout : d->Din, Hq = d->transposed ? d->Hout : d->Hin, Wq = d->transposed ? d->Wout : d->Win;
    const int PC = d->transposed ? d->Cin : d->Cout, QC = d->transposed ? d->Cout : d->Cin;
    const int64_t rows = (int64_t)d->B * Dp * Hp;
    int slabs = (int)((rows + 63) / 64);
    if (slabs > 1024) slabs = 1024;
    const int rows_per_slab = (int)((rows + slabs - 1) / slabs);
    const dim3 grid(mvs_cdiv(Wp, 128), (unsigned)(27 * (PC / 8) * (QC / 8)), (unsigned)slabs);
    MVS_REQUIRE(grid.y <= 65535, MVS_E_SHAPE, "mvs_conv3d_bwd_weight: channel product too large for the launch grid");
    if (d->transposed) {
        MVS_DISPATCH_DTYPE(d->dtype_in, TI,
            MVS_LAUNCH((conv3d_wgrad_kernel<TI, float>), grid, dim3(128), stream, (const TI*)x, grad_y, grad_w, d->B, PC, QC, PC,
                       Dp, Hp, Wp, Dq, Hq, Wq, d->stride, rows_per_slab));
    } else {
        MVS_DISPATCH_DTYPE(d->dtype_in, TI,
            MVS_LAUNCH((conv3d_wgrad_kernel<float, TI>), grid, dim3(128), stream, grad_y, (const TI*)x, grad_w, d->B, PC, QC, PC,
                       Dp, Hp, Wp, Dq, Hq, Wq, d->stride, rows_per_slab));
    }
    return MVS_CHECK_LAUNCH("mvs_conv3d_bwd_weight");
}

// ---------------------------------------------------------------------------------------------- batch norm (training)
// C8 fp32 volumes [B][C/8][S][8]; thread strides over s for one (b, channel block).
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ x, float* __restrict__ sums, int C, int64_t S) {
    const int CB = C / 8;
    const int cb = blockIdx.y % CB;
    const int64_t base = (int64_t)blockIdx.y * S;  // (b * CB + cb) * S
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        float v[8];
        V8<float>::load(x + (base + s) * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s1[k] += v[k]; s2[k] += v[k] * v[k]; }
    }
#ifndef MVS_CPU_EMU
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o); s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o); }
    if ((threadIdx.x & 31) != 0) return;
#endif
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(sums + cb * 8 + k, s1[k]); atomicAdd(sums + C + cb * 8 + k, s2[k]); }
}

__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                  const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ skip,
                  float* __restrict__ y, int C, int64_t S, int64_t total, int relu) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*CB*S vectors
    if (i >= total) return;
    const int cb = (int)((i / S) % (C / 8));
    float v[8];
    V8<float>::load(x + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = cb * 8 + k;
        float o = (v[k] - __ldg(mean + c)) * __ldg(invstd + c) * __ldg(gamma + c) + __ldg(beta + c);
        if (relu) o = fmaxf(o, 0.f);
        v[k] = o;
    }
    if (skip) {
        float sv[8];
        V8<float>::load(skip + i * 8, sv);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += sv[k];
    }
    V8<float>::store(y + i * 8, v);
}

// g = grad_y * [pre-activation > 0];  red[0][c] += g, red[1][c] += g * xhat
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ red, int C, int64_t S, int relu) {
    const int CB = C / 8;
    const int cb = blockIdx.y % CB;
    const int64_t base = (int64_t)blockIdx.y * S;
    float m[8], is[8], ga[8], be[8], r0[8], r1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = cb * 8 + k;
        m[k] = __ldg(mean + c); is[k] = __ldg(invstd + c); ga[k] = __ldg(gamma + c); be[k] = __ldg(beta + c);
        r0[k] = 0.f; r1[k] = 0.f;
    }
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        float v[8], g[8];
        V8<float>::load(x + (base + s) * 8, v);
        V8<float>::load(gy + (base + s) * 8, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float xh = (v[k] - m[k]) * is[k];
            const float gg = (relu && !(xh * ga[k] + be[k] > 0.f)) ? 0.f : g[k];
            r0[k] += gg; r1[k] += gg * xh;
        }
    }
#ifndef MVS_CPU_EMU
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { r0[k] += __shfl_xor_sync(0xffffffffu, r0[k], o); r1[k] += __shfl_xor_sync(0xffffffffu, r1[k], o); }
    if ((threadIdx.x & 31) != 0) return;
#endif
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(red + cb * 8 + k, r0[k]); atomicAdd(red + C + cb * 8 + k, r1[k]); }
}

__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                        const float* __restrict__ red, float* __restrict__ gx, int C, int64_t S, int64_t total, float inv_m,
                        int relu) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cb = (int)((i / S) % (C / 8));
    float v[8], g[8];
    V8<float>::load(x + i * 8, v);
    V8<float>::load(gy + i * 8, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = cb * 8 + k;
        const float is = __ldg(invstd + c), ga = __ldg(gamma + c);
        const float xh = (v[k] - __ldg(mean + c)) * is;
        const float gg = (relu && !(xh * ga + __ldg(beta + c) > 0.f)) ? 0.f : g[k];
        v[k] = ga * is * (gg - __ldg(red + c) * inv_m - xh * __ldg(red + C + c) * inv_m);
    }
    V8<float>::store(gx + i * 8, v);
}

static int check_bn(const char* who, int B, int C, int64_t S) {
    MVS_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && S > 0, MVS_E_SHAPE, "%s: bad dims (C must be a multiple of 8)", who);
    MVS_REQUIRE((int64_t)B * (C / 8) <= 65535, MVS_E_SHAPE, "%s: B*C/8 too large for the launch grid", who);
    return MVS_OK;
}

extern "C" int mvs_bn_stats(const float* x, float* sums, int B, int C, int64_t S, void* stream) {
    MVS_REQUIRE(x && sums, MVS_E_ARG, "mvs_bn_stats: null pointer");
    int rc = check_bn("mvs_bn_stats", B, C, S);
    if (rc) return rc;
    unsigned bx = mvs_cdiv(S, 256 * 16);
    if (bx > 1024) bx = 1024;
    MVS_LAUNCH(bn_stats_kernel, dim3(bx, (unsigned)(B * (C / 8))), dim3(256), stream, x, sums, C, S);
    return MVS_CHECK_LAUNCH("mvs_bn_stats");
}

extern "C" int mvs_bn_act_fwd(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                              const float* skip, float* y, int B, int C, int64_t S, int relu, void* stream) {
    MVS_REQUIRE(x && mean && invstd && gamma && beta && y, MVS_E_ARG, "mvs_bn_act_fwd: null pointer");
    int rc = check_bn("mvs_bn_act_fwd", B, C, S);
    if (rc) return rc;
    const int64_t total = (int64_t)B * (C / 8) * S;
    MVS_LAUNCH(bn_act_fwd_kernel, dim3(mvs_cdiv(total, 256)), dim3(256), stream, x, mean, invstd, gamma, beta, skip, y, C, S, total, relu);
    return MVS_CHECK_LAUNCH("mvs_bn_act_fwd");
}

extern "C" int mvs_bn_act_bwd_reduce(const float* x, const float* grad_y, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, float* red, int B, int C, int64_t S, int relu,
                                     void* stream) {
    MVS_REQUIRE(x && grad_y && mean && invstd && gamma && beta && red, MVS_E_ARG, "mvs_bn_act_bwd_reduce: null pointer");
    int rc = check_bn("mvs_bn_act_bwd_reduce", B, C, S);
    if (rc) return rc;
    unsigned bx = mvs_cdiv(S, 256 * 16);
    if (bx > 1024) bx = 1024;
    MVS_LAUNCH(bn_act_bwd_reduce_kernel, dim3(bx, (unsigned)(B * (C / 8))), dim3(256), stream, x, grad_y, mean, invstd, gamma, beta, red, C, S, relu);
    return MVS_CHECK_LAUNCH("mvs_bn_act_bwd_reduce");
}

extern "C" int mvs_bn_act_bwd_apply(const float* x, const float* grad_y, const float* mean, const float* invstd,
                                    const float* gamma, const float* beta, const float* red, float* grad_x, int B, int C,
                                    int64_t S, int relu, void* stream) {
    MVS_REQUIRE(x && grad_y && mean && invstd && gamma && beta && red && grad_x, MVS_E_ARG, "mvs_bn_act_bwd_apply: null pointer");
    int rc = check_bn("mvs_bn_act_bwd_apply", B, C, S);
    if (rc) return rc;
    const int64_t total = (int64_t)B * (C / 8) * S;
    const float inv_m = 1.f / (float)((double)B * (double)S);
    MVS_LAUNCH(bn_act_bwd_apply_kernel, dim3(mvs_cdiv(total, 256)), dim3(256), stream, x, grad_y, mean, invstd, gamma, beta, red, grad_x, C, S, total, inv_m, relu);
    return MVS_CHECK_LAUNCH("mvs_bn_act_bwd_apply");
}
