// conv3d_tc.cu — 3x3x3 stride-1 convolution as an implicit GEMM on the 5th-generation tensor cores (algo = 2).
//
//   D[128 voxels x N couts] (fp32, TMEM)  +=  A[128 voxels x 16 cin] (smem)  .  B[N couts x 16 cin]^T (smem)
//
// No im2col: a CTA owns a (16 or 8) x 30 output tile in (h, w) and marches along depth.  Every input plane of the tile
// (with its 1-voxel halo, zero-filled by TMA outside the volume) is staged ONCE in shared memory as
//     [cin block][row = hh * 32 + ww][8 cin]                      (hh < PH = TH + 2, ww < 32 = TW + 2)
// which is exactly the canonical no-swizzle K-major UMMA operand with SBO = 128 B: rows are 16 bytes apart, so the A
// operand of tap (kd, kh, kw) is the SAME buffer addressed at (plane + kd, row + kh * 32 + kw) — 27 descriptors
// over one tile instead of 27 gathered copies.  Rows whose ww >= 30 (or hh >= TH) are junk outputs that are never stored.
// The C8 activation layout makes the TMA box {32 w x 8 cin, PH, 1, 1} a run of 512-byte rows.
//
// Warp roles (224 threads): 0 = TMA producer of input planes (4-stage ring over depth), 1 = MMA issuer (one elected
// lane) + TMEM owner, 2 = weight-tile loader (resident when the 27 tap tiles fit, else a ring streamed per plane),
// 3..6 = epilogue (tcgen05.ld -> folded-BN affine, ReLU, skip add -> C8 store), double-buffered against the MMAs.
#include "mvs_rt.h"
#include <cuda.h>

namespace {

constexpr int kThreads = 224;
constexpr int kPW = 32;          // padded tile width (30 outputs + 2 halo columns)
constexpr int kTW = 30;
constexpr int kStages = 4;       // input-plane ring
constexpr int kBStages = 4;      // weight-tile ring (streaming mode)
constexpr int kSmemLimit = 227 * 1024;

struct TcParams {
    const void* w;        // [27][CiB][N][8] storage dtype
    const float* scale;   // [Cout] or null
    const float* shift;   // [Cout] or null
    const void* skip;     // y's layout or null
    void* y;
    int B, CiB, Cout, N, D, H, W;
    int TH, PH, nM, rows_alloc, nwt, nht, LD, nseg;
    int b_resident, relu, flip, is_bf16;
    uint32_t chunk_bytes, plane_bytes, btile_bytes, tmem_cols;
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// no-swizzle K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row core matrices of 16-byte rows,
// SBO = distance between 8-row groups, LBO = distance between the two 8-element K halves of one K=16 step.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);  // version = 1 (sm_100), layout = SWIZZLE_NONE
}

// ------------------------------------------------------------------------------------------------ kernel
template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* planes = smem;
    uint8_t* bsm = planes + (size_t)kStages * p.plane_bytes;
    const uint32_t b_bytes = p.b_resident ? 27u * p.btile_bytes : (uint32_t)kBStages * p.btile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + b_bytes);
    uint64_t* plane_full = bars;                  // [kStages]
    uint64_t* plane_empty = bars + kStages;       // [kStages]
    uint64_t* b_full = bars + 2 * kStages;        // [kBStages]
    uint64_t* b_empty = b_full + kBStages;        // [kBStages]
    uint64_t* acc_full = b_empty + kBStages;      // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x;
    const int wt = t % p.nwt; t /= p.nwt;
    const int ht = t % p.nht; t /= p.nht;
    const int seg = t % p.nseg;
    const int b = t / p.nseg;
    const int w0 = wt * kTW, h0 = ht * p.TH, d0 = seg * p.LD;
    const int nout = min(p.LD, p.D - d0);          // output planes of this CTA
    const int nplanes = nout + 2;                  // input planes d0-1 .. d0+nout

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(plane_full + i, 1); mbar_init(plane_empty + i, 1); }
        for (int i = 0; i < kBStages; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {  // TMEM allocation is warp-collective; the same warp frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== input planes: one TMA box per cin block per plane =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            const uint32_t box_bytes = (uint32_t)p.PH * kPW * 16u;
            for (int j = 0; j < nplanes; ++j) {
                const int s = j % kStages;
                mbar_wait(plane_empty + s, ((j / kStages) & 1) ^ 1);
                mbar_expect_tx(plane_full + s, box_bytes * (uint32_t)p.CiB);
                for (int cb = 0; cb < p.CiB; ++cb)
                    tma_load_4d(&tmap, plane_full + s, planes + (size_t)s * p.plane_bytes + (size_t)cb * p.chunk_bytes,
                                (w0 - 1) * 8, h0 - 1, d0 - 1 + j, b * p.CiB + cb);
            }
        }
    } else if (warp == 2) {
        // ===================== weight tap tiles =====================
        if (lane == 0) {
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w);
            if (p.b_resident) {
                mbar_expect_tx(b_full, 27u * p.btile_bytes);
                for (int tap = 0; tap < 27; ++tap) bulk_load(bsm + (size_t)tap * p.btile_bytes, wsrc + (size_t)tap * p.btile_bytes, p.btile_bytes, b_full);
            } else {
                for (int i = 0, u = 0; i < nout; ++i)
                    for (int tap = 0; tap < 27; ++tap, ++u) {
                        const int s = u % kBStages;
                        mbar_wait(b_empty + s, ((u / kBStages) & 1) ^ 1);
                        mbar_expect_tx(b_full + s, p.btile_bytes);
                        bulk_load(bsm + (size_t)s * p.btile_bytes, wsrc + (size_t)tap * p.btile_bytes, p.btile_bytes, b_full + s);
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D = f32, A/B = f16|bf16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t fmt = p.is_bf16 ? 1u : 0u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t planes_addr = smem_u32(planes), b_addr = smem_u32(bsm);
            const int ksteps = p.CiB / 2;
            if (p.b_resident) mbar_wait(b_full, 0);
            for (int i = 0, u = 0; i < nout; ++i) {
                const int buf = i & 1;
                mbar_wait(acc_empty + buf, ((i >> 1) & 1) ^ 1);
                for (int j = (i == 0 ? 0 : i + 2); j <= i + 2; ++j) mbar_wait(plane_full + (j % kStages), (j / kStages) & 1);
                tc_fence_after();
                for (int tap = 0; tap < 27; ++tap, ++u) {
                    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                    // gather form: Conv3d reads x[o - 1 + k]; stride-1 ConvTranspose3d reads x[o + 1 - k]
                    const int od = p.flip ? 2 - kd : kd, oh = p.flip ? 2 - kh : kh, ow = p.flip ? 2 - kw : kw;
                    uint32_t btile;
                    if (p.b_resident) {
                        btile = b_addr + (uint32_t)tap * p.btile_bytes;
                    } else {
                        const int s = u % kBStages;
                        mbar_wait(b_full + s, (u / kBStages) & 1);
                        tc_fence_after();
                        btile = b_addr + (uint32_t)s * p.btile_bytes;
                    }
                    const uint32_t a_plane = planes_addr + (uint32_t)((i + od) % kStages) * p.plane_bytes + (uint32_t)(oh * kPW + ow) * 16u;
                    for (int m = 0; m < p.nM; ++m) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)((buf * p.nM + m) * p.N);
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t ad = smem_desc(a_plane + (uint32_t)(2 * ks) * p.chunk_bytes + (uint32_t)m * 128u * 16u, p.chunk_bytes, 128u);
                            const uint64_t bd = smem_desc(btile + (uint32_t)(2 * ks) * (uint32_t)p.N * 16u, (uint32_t)p.N * 16u, 128u);
                            umma_f16(d_tmem, ad, bd, idesc, (tap | ks) != 0 ? 1u : 0u);
                        }
                    }
                    if (!p.b_resident) umma_commit(b_empty + (u % kBStages));
                }
                umma_commit(plane_empty + (i % kStages));   // plane i is the oldest of the three: free once these MMAs retire
                umma_commit(acc_full + buf);
            }
        }
    } else {
        // ===================== epilogue: TMEM -> affine / ReLU / skip -> C8 store =====================
        const int quad = warp & 3;                 // TMEM lanes [32 quad, 32 quad + 32) belong to this warp
        const int CoB = (p.Cout + 7) / 8;
        const int64_t HW = (int64_t)p.H * p.W;
        for (int i = 0; i < nout; ++i) {
            const int buf = i & 1;
            mbar_wait(acc_full + buf, (i >> 1) & 1);
            tc_fence_after();
            const int q = d0 + i;
            for (int m = 0; m < p.nM; ++m) {
                const int r = m * 128 + quad * 32 + lane;
                const int hh = r >> 5, ww = r & 31;
                const int h = h0 + hh, w = w0 + ww;
                const bool valid = (hh < p.TH) && (ww < kTW) && (h < p.H) && (w < p.W);
                for (int c0 = 0; c0 < p.Cout; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((buf * p.nM + m) * p.N + c0), v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (!valid) continue;
                    if (p.Cout == 1) {
                        float o = __uint_as_float(v[0]);
                        if (p.scale) o *= __ldg(p.scale);
                        if (p.shift) o += __ldg(p.shift);
                        if (p.relu) o = fmaxf(o, 0.f);
                        const int64_t off = ((int64_t)b * p.D + q) * HW + (int64_t)h * p.W + w;
                        if (p.skip) o += reinterpret_cast<const float*>(p.skip)[off];
                        reinterpret_cast<float*>(p.y)[off] = o;
                        continue;
                    }
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int cb = c0 / 8 + half;
                        if (cb >= CoB) break;
                        float o[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int co = cb * 8 + k;
                            float x = __uint_as_float(v[half * 8 + k]);
                            if (p.scale) x *= __ldg(p.scale + co);
                            if (p.shift) x += __ldg(p.shift + co);
                            if (p.relu) x = fmaxf(x, 0.f);
                            o[k] = x;
                        }
                        const int64_t off = ((((int64_t)b * CoB + cb) * p.D + q) * HW + (int64_t)h * p.W + w) * 8;
                        if (p.skip) {
                            float sv[8];
                            V8<T>::load(reinterpret_cast<const T*>(p.skip) + off, sv);
#pragma unroll
                            for (int k = 0; k < 8; ++k) o[k] += sv[k];
                        }
                        V8<T>::store(reinterpret_cast<T*>(p.y) + off, o);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + buf);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ weights: G fp32 -> tap tiles
// wtc[tap][cib][n][8] = G[tap][cib*8 + k][n]  (zero for n >= CoutPad)
template <typename T>
__global__ void pack_weight_tc_kernel(const float* __restrict__ g, T* __restrict__ w, int Cin, int CoutPad, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 27 * CiB * N
    const int CiB = Cin / 8;
    if (i >= 27 * CiB * N) return;
    const int n = i % N, cib = (i / N) % CiB, tap = i / (N * CiB);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = n < CoutPad ? g[((int64_t)tap * Cin + cib * 8 + k) * CoutPad + n] : 0.f;
    V8<T>::store(w + (int64_t)i * 8, v);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int n_pad(int cout) { return cout <= 16 ? 16 : (cout + 15) / 16 * 16; }

}  // namespace

int mvs_conv3d_tc_supported(const mvs_conv3d_desc* d) {
    if (d->stride != 1) return 0;
    if (d->dtype_in != MVS_F16 && d->dtype_in != MVS_BF16) return 0;
    if (d->Cout != 1 && d->dtype_out != d->dtype_in) return 0;
    if (d->Cin % 16 != 0 || d->Cin > 64) return 0;
    if (d->Cout != 1 && (d->Cout % 8 != 0 || d->Cout > 64)) return 0;
    return 1;
}

int64_t mvs_conv3d_tc_workspace_bytes(const mvs_conv3d_desc* d) {
    if (!mvs_conv3d_tc_supported(d)) return 0;
    return (int64_t)27 * (d->Cin / 8) * n_pad(d->Cout) * 16;  // tap tiles [27][Cin/8][N][8] in the storage dtype
}

int mvs_conv3d_fwd_tc(const mvs_conv3d_desc* d, const void* x, const float* g, const float* scale, const float* shift,
                      const void* skip, void* y, void* ws, void* stream) {
    MVS_REQUIRE(mvs_conv3d_tc_supported(d), MVS_E_UNSUPPORTED,
                "mvs_conv3d_fwd: tcgen05 path needs stride 1, fp16/bf16 storage, Cin in {16,32,48,64}, Cout in {1,8..64}");
    EncodeTiledFn enc = encode_tiled();
    MVS_REQUIRE(enc, MVS_E_LAUNCH, "mvs_conv3d_fwd: cuTensorMapEncodeTiled is not available from this driver");
    const int N = n_pad(d->Cout), CiB = d->Cin / 8, CoutPad = (d->Cout + 7) / 8 * 8;
    cudaStream_t st = (cudaStream_t)stream;

    // ---- tap tiles, re-packed from the gather form into the caller's workspace on every call
    MVS_REQUIRE(ws, MVS_E_ARG, "mvs_conv3d_fwd: the tcgen05 path needs a workspace of mvs_conv3d_workspace_bytes() bytes");
    void* wt = ws;
    // (<= 110 K elements: negligible next to the convolution, and always coherent with in-place weight updates)
    if (d->dtype_in == MVS_F16) pack_weight_tc_kernel<__half><<<mvs_cdiv(27 * CiB * N, 256), 256, 0, st>>>(g, (__half*)wt, d->Cin, CoutPad, N);
    else pack_weight_tc_kernel<__nv_bfloat16><<<mvs_cdiv(27 * CiB * N, 256), 256, 0, st>>>(g, (__nv_bfloat16*)wt, d->Cin, CoutPad, N);

    // ---- tiling plan
    TcParams p;
    p.w = wt; p.scale = scale; p.shift = shift; p.skip = skip; p.y = y;
    p.B = d->B; p.CiB = CiB; p.Cout = d->Cout; p.N = N; p.D = d->Din; p.H = d->Hin; p.W = d->Win;
    p.relu = d->relu; p.flip = d->transposed; p.is_bf16 = d->dtype_in == MVS_BF16;
    p.btile_bytes = (uint32_t)CiB * N * 16;
    p.b_resident = 27u * p.btile_bytes <= 56u * 1024u;
    const uint32_t b_bytes = p.b_resident ? 27u * p.btile_bytes : (uint32_t)kBStages * p.btile_bytes;
    int nM = 4;
    for (;; --nM) {  // largest M-tile count whose 4-stage plane ring fits beside the weights and 2 accumulator sets fit TMEM
        p.nM = nM; p.TH = 4 * nM; p.PH = p.TH + 2;
        p.rows_alloc = (nM * 128 + 2 * kPW + 2 + 7) / 8 * 8;
        p.chunk_bytes = (uint32_t)p.rows_alloc * 16u;
        p.plane_bytes = p.chunk_bytes * (uint32_t)CiB;
        const size_t need = (size_t)kStages * p.plane_bytes + b_bytes + 256;
        if ((need <= (size_t)kSmemLimit && 2 * nM * N <= 512) || nM == 1) break;
    }
    const size_t smem = (size_t)kStages * p.plane_bytes + b_bytes + 256;
    MVS_REQUIRE(smem <= (size_t)kSmemLimit, MVS_E_UNSUPPORTED, "mvs_conv3d_fwd: tcgen05 tile does not fit shared memory (%zu bytes)", smem);
    uint32_t cols = 32;
    while ((int)cols < 2 * p.nM * N) cols <<= 1;
    p.tmem_cols = cols;
    p.nwt = (p.W + kTW - 1) / kTW;
    p.nht = (p.H + p.TH - 1) / p.TH;
    // depth segments: enough CTAs for >= ~3 waves of 148 SMs, but >= 8 planes each (2 halo planes per segment)
    const int64_t tiles = (int64_t)p.B * p.nwt * p.nht;
    int LD = p.D;
    while (LD > 8 && tiles * ((p.D + LD - 1) / LD) < 148 * 3) LD = (LD + 1) / 2;
    p.LD = LD; p.nseg = (p.D + LD - 1) / LD;
    const int64_t nblocks = tiles * p.nseg;
    MVS_REQUIRE(nblocks < (1ll << 31), MVS_E_SHAPE, "mvs_conv3d_fwd: too many tiles");

    // ---- tensor map over x: dims (innermost first) {W*8, H, D, B*CiB}, box {256, PH, 1, 1}, zero fill outside
    CUtensorMap tmap;
    const cuuint64_t gdim[4] = {(cuuint64_t)p.W * 8, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)p.B * CiB};
    const cuuint64_t gstr[3] = {(cuuint64_t)p.W * 16, (cuuint64_t)p.H * p.W * 16, (cuuint64_t)p.D * p.H * p.W * 16};
    const cuuint32_t box[4] = {(cuuint32_t)kPW * 8, (cuuint32_t)p.PH, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult cr = enc(&tmap, p.is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), gdim,
                            gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVS_REQUIRE(cr == CUDA_SUCCESS, MVS_E_LAUNCH, "mvs_conv3d_fwd: cuTensorMapEncodeTiled failed (%d)", (int)cr);

    if (p.is_bf16) {
        cudaFuncSetAttribute(conv3d_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv3d_tc_kernel<__nv_bfloat16><<<(unsigned)nblocks, kThreads, smem, st>>>(tmap, p);
    } else {
        cudaFuncSetAttribute(conv3d_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv3d_tc_kernel<__half><<<(unsigned)nblocks, kThreads, smem, st>>>(tmap, p);
    }
    return MVS_CHECK_LAUNCH("mvs_conv3d_fwd (tcgen05)");
}
