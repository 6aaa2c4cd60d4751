// conv3d_tc.cu — 3x3x3 (transposed) convolutions of the regularisation U-Nets as implicit GEMMs on the 5th-generation
// tensor cores (algo = 2):
//
//   D[128 positions x N] (fp32, TMEM)  +=  A[128 positions x 16 cin] (smem)  .  B[N x 16 cin]^T (smem)
//
// No im2col.  A CTA owns a (4 nM) x 30 tile of positions in (h, w) and marches along depth.  Every input plane of the tile
// (with its halo, zero-filled by TMA outside the volume) is staged ONCE in shared memory as
//     [cin block][row = hh * 32 + ww][8 cin]
// which is the canonical no-swizzle K-major UMMA operand with SBO = 128 B: rows are 16 bytes apart, so the A operand of
// a filter tap is the SAME buffer addressed at (plane slot, row + shift) — a list of shifted descriptors over one tile
// instead of gathered copies.  The C8 activation layout makes the TMA box {32 w x 8 cin, PH, 1, 1} a run of 512-byte rows.
// Rows with ww >= 30 or hh >= TH are junk accumulator rows that are never stored.
//
// A tcgen05.mma costs ~100 cycles of A-operand fetch per 128 x 16 tile whatever N is (measured: N = 16 and N = 32 take the
// same time, and the kernel time is #MMAs x ~97 cycles), so filter taps that read the SAME A rows are folded into N:
// one kernel, three tap programs built on the host (struct Entry), each MMA writing `nblk` column blocks of Cout:
//   S1  stride-1 Conv3d / ConvTranspose3d : slots = planes d-1, d, d+1; 9 entries (kd, kh) with row shift kh*32; the three
//                                           kw taps are the 3 column blocks and the epilogue adds lanes l, l+1, l+2 (shuffles).
//   S2  stride-2 Conv3d                   : the input is staged de-interleaved by (h, w) parity (4 strided tensor maps),
//                                           so tap k reads parity (k != 1) at shift {0,1,1}[k]; 27 entries, one block.
//   T2  stride-2 ConvTranspose3d          : positions are INPUT voxels; 8 entries = input offsets (od, oh, ow) in {0,1}^3 and
//                                           8 column blocks = output parities, so each lane stores a 2x2x2 block of voxels.
//   Cin = 8 layers pair two taps into one K = 16 step: the descriptor's LBO is the row distance between the two taps.
//
// Warp roles (448 threads): 0 = TMA producer of input planes (ring over depth); 1..4 = MMA issuers, one elected lane each,
// issuer k owns the accumulators of M-tile k (these MMAs are only N = 16..64 wide, so ONE issuing thread is the
// bottleneck: measured 225 cycles per MMA against ~36 cycles of operand reads) and warp 1 also owns TMEM; 5 = weight-tile
// loader (resident when all tiles fit, else a ring streamed per step); 6..13 = epilogue, two warps per TMEM lane quadrant
// (tcgen05.ld -> folded-BN affine, ReLU, skip add -> C8 store), double-buffered against the MMAs through TMEM.
#include "mvs_rt.h"
#include <cuda.h>
#include <stdlib.h>

namespace {

#ifndef MVS_TC_EPIWARPS
#define MVS_TC_EPIWARPS 8
#endif
// 8: two warps per TMEM lane quadrant taking alternate M-tiles; 16: one warp per (quadrant, M-tile).  Measured equal (2.48 vs 2.51 ms
// per 4-item step): a step is bound by the MMA issue pattern, not by the epilogue (tools/umma_pattern.cu, and the trace with the
// epilogue body compiled out keeps the same step period), so the smaller block with 96 registers and no spills is kept.
constexpr int kEpiWarps = MVS_TC_EPIWARPS;
constexpr int kMStride = kEpiWarps / 4;      // M-tiles are dealt round-robin to the warps of a quadrant
constexpr int kThreads = (6 + kEpiWarps) * 32;   // producer | 4 MMA issuers | weight loader | epilogue warps
constexpr int kPW = 32;          // staged tile width (30 positions + halo)
constexpr int kTW = 30;
constexpr int kBStages = 16;     // max depth of the weight-tile ring (streaming mode; p.bstages are used)
constexpr int kMaxStages = 12;   // max depth of the input-slot ring (p.stages are used)
constexpr int kMaxEntries = 27;
constexpr int kAccRing = 4;      // accumulator ring slots of the kd-folded program (output planes in flight)
constexpr int kPG = 32;          // kd-fold: TMEM columns per output plane (3 kw blocks of 8 + pad; N must be a multiple of 16)
constexpr int kSmemLimit = 227 * 1024;
constexpr int kTail = 1536;      // barriers (512 B) + TMEM slot (16 B) + affine table (512 B) + entry table (432 B) behind the rings
enum { MODE_S1 = 0, MODE_S2 = 1, MODE_T2 = 2 };

struct Entry {
    int16_t row_shift;     // rows added to the A start address
    uint16_t lbo_rows;     // 0: K halves are consecutive cin blocks (LBO = chunk stride); else paired taps, LBO = rows * 16 B
    uint8_t slot_off;      // input slot relative to the first live slot of the step
    uint8_t sub;           // parity sub-plane (S2)
    uint8_t group;         // accumulator group (always 0 now: parities are column blocks)
    uint8_t first;         // first entry of its group: overwrite instead of accumulate
};

struct TcParams {
    const void* w;         // [entry][kchunk][N][8] storage dtype
    const float* scale;    // [Cout] or null
    const float* shift;    // [Cout] or null
    const void* skip;      // y's layout or null
    void* y;
    int mode, B, CiB, Cout, CoP, nblk, N, ksteps, kchunks;   // CoP = columns per block, nblk blocks folded into N
    int Di, Hi, Wi, Do, Ho, Wo;      // input / output grids
    int Dt, Ht, Wt;                  // grid the tiles walk (S1/S2: output, T2: input)
    int TH, PH, nM, nwt, nht, LD, nseg;
    int nsub, stages, sps, live, groups, nentries, bstages, nwork;
    int b_resident, relu, is_bf16, kdfold;
    int d_mul, d_org;                // depth coordinate of slot j of an item: d_mul * d0 + d_org + j
    int kwfold;                      // stride-1 programs: the 3 kw taps are column blocks summed by the epilogue (else 9+ shifted entries)
    float slope;                     // activation as max(v, v * slope): 0 = ReLU, 1 = none, 0.1 = LeakyReLU(0.1)
    int64_t ys_b, ys_cb, ys_d, ys_h, y_org;   // output (and skip) addressing in voxels: b, channel block, d, h strides + origin
    uint32_t chunk_bytes, sub_bytes, slot_bytes, btile_bytes, tmem_cols;
    Entry prog[kMaxEntries];
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
#ifdef MVS_TC_TESTWAIT
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#endif
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// (Measured and rejected: one lane polling + __syncwarp() for the warp-uniform roles is ~40 % SLOWER than all 32 lanes executing
// the try_wait: conv0 391 -> 575 us, prob 199 -> 318 us at 4 items.)
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
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
// Warp-uniform callers: every lane executes the call, one elected lane issues.  Keeping the issue loop warp-uniform lets the
// compiler hold descriptors in uniform registers instead of moving them there (R2UR) for every tcgen05.mma.
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptors (cute::UMMA::SmemDescriptor), no-swizzle K-major: 8-row core matrices of 16-byte rows;
// lo word = start >> 4 | (LBO >> 4) << 16 with LBO = distance between the two 8-element K halves of one K = 16 step,
// hi word = SBO >> 4 (distance between 8-row groups = 128 B) | version 1 (sm_100) at bit 46; layout type 0 = SWIZZLE_NONE.
// They are assembled inline by the MMA issuers (one add per K step).

template <typename T> __device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]);
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& raw, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& raw, float (&v)[8]) {
    const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(u[i] << 16); v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
}

// 256-bit global accesses (sm_100: LDG / STG .256): the two w-parity voxels of a transposed stride-2 output row are adjacent
// 16-byte vectors, so one lane moves both with one instruction and a warp's access is 1 KB contiguous instead of 32 half-used
// 32-byte sectors per instruction.
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
template <typename T> __device__ __forceinline__ uint4 pack8(const float (&v)[8]);
template <> __device__ __forceinline__ uint4 pack8<__half>(const float (&v)[8]) {
    uint4 r; __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    return r;
}
template <> __device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float (&v)[8]) {
    uint4 r; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return r;
}

struct TensorMaps { CUtensorMap m[4]; };

#ifdef MVS_TC_TRACE
// Debug build only (tools/tc_trace.py): CTA 0 records clock64() at pipeline events; 8 lanes x 1024 slots.
__device__ long long g_trace[10][1024];
__device__ int g_tcount;
#define TRACE(lane_id, idx) do { if (blockIdx.x == 0 && (idx) < 1024) g_trace[lane_id][idx] = clock64(); } while (0)
#else
#define TRACE(lane_id, idx) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------ epilogue of one M-tile row
// One accumulator row (= one lane) -> NOV output voxels x CoB channel blocks.  Every global read (skip tensor) and every
// TMEM read of a channel block is issued before the first use, so their latencies overlap instead of adding up.
//   NOV = 8 : transposed stride 2, column block ov = output parity (pd, ph, pw) of the 2x2x2 voxel block of this input voxel
//   KWFOLD  : stride 1, column blocks 0,1,2 hold the taps reading input column (lane): output j = blk0[j] + blk1[j+1] + blk2[j+2]
// One lane's output row of NT M-tiles at once (NT = 2: the TMEM and skip loads of both tiles are in flight together, which
// halves the exposed latency of this otherwise serial, single-warp code).
struct RowAt { uint32_t trow; bool valid; int64_t base; };   // base = b * ys_b + oh * ys_h + ow + y_org: constant over the depth steps of a work item

template <typename T, int NOV, bool KWFOLD, bool C1, bool SKIP, int NT>
__device__ __forceinline__ void epilogue_rows_v(const TcParams& p, const float* __restrict__ aff, const RowAt (&ra)[NT], int b, int od0,
                                                int CoB, int64_t HWo) {
    constexpr int NLD = KWFOLD ? 3 : NOV;
    constexpr int64_t vs = C1 ? 1 : 8;
    const int64_t plane = p.ys_d, row = p.ys_h;
    const float slope = p.slope;                            // branch-free activation: max(v, v * slope)
#ifdef MVS_TC_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 192) { g_trace[8][g_tcount & 1023] = clock64(); }
#endif
    for (int cb = 0; cb < CoB; ++cb) {
        // element offset of the row's first output voxel; the other voxels of a 2x2x2 block are +pd*plane +ph*row +pw
        int64_t off0[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t)
            off0[t] = (ra[t].base + (int64_t)cb * p.ys_cb + (int64_t)od0 * plane) * vs;
        // folded-BN affine of this channel block: read once (the asm memory clobbers below would force re-reads per voxel)
        float sc[8], sh[8];
#pragma unroll
        for (int k = 0; k < (C1 ? 1 : 8); ++k) { sc[k] = aff[cb * 8 + k]; sh[k] = aff[64 + cb * 8 + k]; }
        // the 8 voxels of a transposed stride-2 row go in two batches of 4 (one output depth plane each): half the registers
        constexpr int NB = NOV == 8 ? 4 : NOV, NLB = KWFOLD ? 3 : NB;
#pragma unroll
        for (int o0 = 0; o0 < NOV; o0 += NB) {
            uint4 sk[NT][NB];
            constexpr bool PAIR = (NOV == 8) && !C1;       // transposed stride 2, C8 output: voxels q, q + 1 (w parity) are adjacent
            if (SKIP && PAIR) {
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int q = 0; q < NB; q += 2) {
                        const int ov = o0 + q;
                        sk[t][q] = sk[t][q + 1] = make_uint4(0u, 0u, 0u, 0u);
                        const int64_t off = off0[t] + ((ov >> 2) * plane + ((ov >> 1) & 1) * row) * vs;
                        if (ra[t].valid) ldg256(reinterpret_cast<const T*>(p.skip) + off, sk[t][q], sk[t][q + 1]);
                    }
            } else if (SKIP) {
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        const int ov = o0 + q;
                        sk[t][q] = make_uint4(0u, 0u, 0u, 0u);
                        const int64_t off = off0[t] + (NOV == 8 ? ((ov >> 2) * plane + ((ov >> 1) & 1) * row + (ov & 1)) * vs : 0);
                        if (ra[t].valid) {
                            if (C1) sk[t][q].x = __float_as_uint(__ldg(reinterpret_cast<const float*>(p.skip) + off));
                            else sk[t][q] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.skip) + off));
                        }
                    }
            }
            uint32_t v[NT][NLB][8];
#ifdef MVS_TC_TRACE
    #ifndef MVS_TC_TRACE2
        if (blockIdx.x == 0 && threadIdx.x == 192 && cb == 0 && o0 == 0) { g_trace[6][g_tcount & 1023] = clock64(); }
#endif
#endif
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int l = 0; l < NLB; ++l) tmem_ld8(ra[t].trow + (uint32_t)((KWFOLD ? l : o0 + l) * p.CoP + cb * 8), v[t][l]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#ifdef MVS_TC_TRACE
            if (blockIdx.x == 0 && threadIdx.x == 192 && cb == 0 && o0 == 0) { g_trace[7][g_tcount & 1023] = clock64(); g_tcount++; }
#endif
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const int ov = o0 + q;
                    float o[8];
#pragma unroll
                    for (int k = 0; k < (C1 ? 1 : 8); ++k) {
                        if (KWFOLD) o[k] = __uint_as_float(v[t][0][k]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v[t][1][k]), 1) +
                                           __shfl_down_sync(0xffffffffu, __uint_as_float(v[t][2][k]), 2);
                        else o[k] = __uint_as_float(v[t][q][k]);
                    }
                    const int64_t off = off0[t] + (NOV == 8 ? ((ov >> 2) * plane + ((ov >> 1) & 1) * row + (ov & 1)) * vs : 0);
                    if (C1) {
                        float x = o[0] * sc[0] + sh[0];
                        x = fmaxf(x, x * slope);
                        if (SKIP) x += __uint_as_float(sk[t][q].x);
                        if (ra[t].valid) reinterpret_cast<float*>(p.y)[off] = x;
                        continue;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) { const float a = o[k] * sc[k] + sh[k]; o[k] = fmaxf(a, a * slope); }
                    if (SKIP) {
                        float sv[8];
                        unpack8<T>(sk[t][q], sv);
#pragma unroll
                        for (int k = 0; k < 8; ++k) o[k] += sv[k];
                    }
                    if (PAIR) {
                        // the even voxel's vector waits in sk[t][q] (its skip value is consumed); the odd one stores both
                        if ((q & 1) == 0) sk[t][q] = pack8<T>(o);
                        else if (ra[t].valid) stg256(reinterpret_cast<T*>(p.y) + off - 8, sk[t][q - 1], pack8<T>(o));
                    } else if (ra[t].valid) V8<T>::store(reinterpret_cast<T*>(p.y) + off, o);
                }
        }
    }
#ifdef MVS_TC_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 192) { g_trace[9][(g_tcount - 1) & 1023] = clock64(); }
#endif
}

// Runtime (warp-uniform) selection of the specialised epilogue: single-channel output or C8, with or without a skip tensor.
template <typename T, int NOV, bool KWFOLD, int NT>
__device__ __forceinline__ void epilogue_rows(const TcParams& p, const float* __restrict__ aff, const RowAt (&ra)[NT], int b, int od0,
                                              int CoB, int64_t HWo) {
    if (p.Cout == 1) {
        if (p.skip) epilogue_rows_v<T, NOV, KWFOLD, true, true, NT>(p, aff, ra, b, od0, CoB, HWo);
        else epilogue_rows_v<T, NOV, KWFOLD, true, false, NT>(p, aff, ra, b, od0, CoB, HWo);
    } else {
        if (p.skip) epilogue_rows_v<T, NOV, KWFOLD, false, true, NT>(p, aff, ra, b, od0, CoB, HWo);
        else epilogue_rows_v<T, NOV, KWFOLD, false, false, NT>(p, aff, ra, b, od0, CoB, HWo);
    }
}

// L2 prefetch of the skip vectors one accumulator row will add (same addressing as epilogue_rows).  The skip tensor was written
// several layers earlier and is cold in DRAM: without this each batch of 8 loads costs a full DRAM round trip per M-tile
// (measured: 13 k cycles per step with the skip, 5 k without).
template <typename T, int NOV>
__device__ __forceinline__ void prefetch_skip_rows(const TcParams& p, bool valid, int b, int od0, int oh0, int ow0, int CoB, int64_t HWo) {
    if (!valid || !p.skip) return;
    const int64_t plane = p.ys_d, row = p.ys_h;
    for (int cb = 0; cb < CoB; ++cb) {
        const int64_t vs = p.Cout == 1 ? 1 : 8;
        const int64_t off0 = ((int64_t)b * p.ys_b + (int64_t)cb * p.ys_cb + (int64_t)od0 * plane + (int64_t)oh0 * row + ow0 + p.y_org) * vs;
#pragma unroll
        for (int ov = 0; ov < NOV; ov += (NOV == 8 ? 2 : 1)) {   // the two w-parity voxels share a 32-byte sector
            const int64_t off = off0 + (NOV == 8 ? ((ov >> 2) * plane + ((ov >> 1) & 1) * row) * vs : 0);
            const void* a = p.Cout == 1 ? (const void*)(reinterpret_cast<const float*>(p.skip) + off) : (const void*)(reinterpret_cast<const T*>(p.skip) + off);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
    }
}

// ------------------------------------------------------------------------------------------------ kernel
template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_tc_kernel(const __grid_constant__ TensorMaps maps, const __grid_constant__ TcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* slots = smem;
    uint8_t* bsm = slots + (size_t)p.stages * p.slot_bytes;
    const uint32_t b_bytes = p.b_resident ? (uint32_t)p.nentries * p.btile_bytes : (uint32_t)p.bstages * p.btile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + b_bytes);
    uint64_t* slot_full = bars;                    // [kMaxStages]
    uint64_t* slot_empty = bars + kMaxStages;      // [kMaxStages]
    uint64_t* b_full = bars + 2 * kMaxStages;      // [kBStages]
    uint64_t* b_empty = b_full + kBStages;         // [kBStages]
    uint64_t* acc_full = b_empty + kBStages;       // [kAccRing]  (2 are used unless kd-folded)
    uint64_t* acc_empty = acc_full + kAccRing;     // [kAccRing]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccRing);
    float* aff = reinterpret_cast<float*>(tmem_slot + 4);    // [2][64] folded-BN scale / shift (1 / 0 when absent)
    uint4* etab = reinterpret_cast<uint4*>(aff + 128);       // [kMaxEntries] decoded tap program for the MMA issuers

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Persistent CTA: work items (tile x depth segment) are dealt round-robin; every role walks the same list, and the
    // slot / accumulator / weight rings simply keep running across items (no pipeline drain or TMEM re-allocation between).
    struct Work { int b, w0, h0, d0, nsteps, nslots; };
    auto decode = [&](int t) {
        Work k;
        const int wt = t % p.nwt; t /= p.nwt;
        const int ht = t % p.nht; t /= p.nht;
        const int seg = t % p.nseg;
        k.b = t / p.nseg; k.w0 = wt * kTW; k.h0 = ht * p.TH; k.d0 = seg * p.LD;
        k.nsteps = min(p.LD, p.Dt - k.d0);                     // depth steps of this item
        k.nslots = (k.nsteps - 1) * p.sps + p.live;            // input slots it consumes
        return k;
    };

    if (threadIdx.x == 0) {
        // every issuer commits its own MMAs, so the barriers the MMAs release count one arrival per issuer
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(slot_full + i, 1); mbar_init(slot_empty + i, (uint32_t)p.nM); }
        for (int i = 0; i < kBStages; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, (uint32_t)p.nM); }
        for (int i = 0; i < kAccRing; ++i) { mbar_init(acc_full + i, (uint32_t)p.nM); mbar_init(acc_empty + i, kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int c = threadIdx.x; c < 128; c += blockDim.x) {
        const int co = c & 63;
        aff[c] = c < 64 ? ((p.scale && co < p.Cout) ? __ldg(p.scale + co) : 1.f) : ((p.shift && co < p.Cout) ? __ldg(p.shift + co) : 0.f);
    }
    if ((int)threadIdx.x < p.nentries) {
        const Entry en = p.prog[threadIdx.x];
        etab[threadIdx.x] = make_uint4((uint32_t)en.sub * p.sub_bytes + (uint32_t)((int)en.row_shift * 16),
                                       (en.lbo_rows ? (uint32_t)en.lbo_rows : (p.chunk_bytes >> 4)) << 16, (uint32_t)en.slot_off, (uint32_t)en.first | ((uint32_t)en.group << 8));
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
        // ===================== input slots: one TMA box per (sub-plane, cin block) =====================
        if (lane == 0) {
            for (int s = 0; s < p.nsub; ++s) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[s]) : "memory");
            const uint32_t box_bytes = (uint32_t)p.PH * kPW * 16u;
            int J = 0;                                             // slots loaded so far by this CTA
            for (int t = blockIdx.x; t < p.nwork; t += gridDim.x) {
                const Work k = decode(t);
                for (int j = 0; j < k.nslots; ++j, ++J) {
                    const int st = J % p.stages;
                    mbar_wait(slot_empty + st, ((J / p.stages) & 1) ^ 1);
                    mbar_expect_tx(slot_full + st, box_bytes * (uint32_t)(p.CiB * p.nsub));
                    uint8_t* dst = slots + (size_t)st * p.slot_bytes;
                    for (int s = 0; s < p.nsub; ++s)
                        for (int cb = 0; cb < p.CiB; ++cb, dst += p.chunk_bytes) {
                            const int dcoord = p.d_mul * k.d0 + p.d_org + j;
                            if (p.mode == MODE_S1) tma_load_4d(&maps.m[0], slot_full + st, dst, (k.w0 - 1) * 8, k.h0 - 1, dcoord, k.b * p.CiB + cb);
                            else if (p.mode == MODE_T2) tma_load_4d(&maps.m[0], slot_full + st, dst, k.w0 * 8, k.h0, dcoord, k.b * p.CiB + cb);
                            else tma_load_5d(&maps.m[s], slot_full + st, dst, 0, k.w0 - 1, k.h0 - 1, dcoord, k.b * p.CiB + cb);
                        }
                    TRACE(0, J);
                }
            }
        }
    } else if (warp == 5) {
        // ===================== weight tiles (one per program entry) =====================
        if (lane == 0) {
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w);
            if (p.b_resident) {
                mbar_expect_tx(b_full, (uint32_t)p.nentries * p.btile_bytes);
                for (int e = 0; e < p.nentries; ++e) bulk_load(bsm + (size_t)e * p.btile_bytes, wsrc + (size_t)e * p.btile_bytes, p.btile_bytes, b_full);
            } else {
                int u = 0;
                for (int t = blockIdx.x; t < p.nwork; t += gridDim.x) {
                    const Work k = decode(t);
                    for (int i = 0; i < k.nsteps; ++i)
                        for (int e = 0; e < p.nentries; ++e, ++u) {
                            const int st = u % p.bstages;
                            mbar_wait(b_empty + st, ((u / p.bstages) & 1) ^ 1);
                            mbar_expect_tx(b_full + st, p.btile_bytes);
                            bulk_load(bsm + (size_t)st * p.btile_bytes, wsrc + (size_t)e * p.btile_bytes, p.btile_bytes, b_full + st);
                        }
                }
            }
        }
    } else if (warp >= 1 && warp <= 4) {
        // ===================== MMA issuers: issuer m accumulates M-tile m =====================
        // The whole warp runs this (warp-uniform) code and one elected lane issues each tcgen05.mma / commit.  Measured on the
        // way here: a step's issue loop took ~4500 cycles per issuer when one lane ran it with the tap decode, local-memory
        // segment arrays and register->uniform moves per operand, against ~1800 cycles of operand fetch; so per-entry data
        // comes decoded from shared memory (`etab`), ring positions advance by compare-and-subtract, the (at most 3) column
        // segments of a kd-folded step live in scalars, and a K step is two adds.
        const int m = __shfl_sync(0xffffffffu, warp - 1, 0);
        if (m < p.nM) {
            // instruction descriptor: D = f32, A/B = f16|bf16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t fmt = p.is_bf16 ? 1u : 0u;
            const uint32_t idesc0 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((128u >> 4) << 24);
            const uint32_t idesc = idesc0 | ((uint32_t)(p.N >> 3) << 17);
            const uint32_t slots_addr = smem_u32(slots) + (uint32_t)m * 128u * 16u, b_addr = smem_u32(bsm);
            // descriptor words: lo = start >> 4 | (LBO >> 4) << 16 ; hi = SBO >> 4 (128 B) | version 1 at bit 46
            const uint64_t desc_hi = (uint64_t)(8u | (1u << 14)) << 32;
            const uint32_t a_kstep = (2u * p.chunk_bytes) >> 4, b_kstep = 2u * (uint32_t)p.N, b_lbo = (uint32_t)p.N << 16;
            const uint32_t btile16 = p.btile_bytes >> 4;
            const int nent = p.nentries, ksteps = p.ksteps, stages = p.stages;
            if (p.b_resident) mbar_wait(b_full, 0);
            if (p.kdfold) {
                // ---- kd-folded stride-1 program (Cout <= 8): a step is an INPUT plane p; its kh entries multiply all 9 (kd, kw)
                // taps at once: N = 3 plane groups of kPG columns, for the output planes p-1, p, p+1, which live in a ring of
                // kAccRing accumulator slots per M-tile.  Plane p+1 is first touched here (overwrite), so in the step's first MMA
                // group 2 stands alone, and groups split wherever their slots wrap around the ring.
                const uint32_t kb_lbo = (uint32_t)(3 * kPG) << 16, kb_kstep = 2u * 3u * kPG;
                const uint32_t d_m = tmem_base + (uint32_t)(m * kAccRing * kPG);
                const uint32_t id1 = idesc0 | ((uint32_t)(kPG >> 3) << 17), id2 = idesc0 | ((uint32_t)((2 * kPG) >> 3) << 17),
                               id3 = idesc0 | ((uint32_t)((3 * kPG) >> 3) << 17);
                int st = 0, sph = 0, Qb = 0;   // ring stage / phase of the next input slot, output planes started (running over items)
                for (int t = blockIdx.x; t < p.nwork; t += gridDim.x) {
                    const Work k = decode(t);
                    for (int j = 0; j < k.nslots; ++j) {
                        mbar_wait(slot_full + st, (uint32_t)sph);
                        if (m == 0 && lane == 0) TRACE(1, Qb + 2 * (t / gridDim.x) + j);
                        if (j < k.nsteps) { const int q = Qb + j; mbar_wait(acc_empty + (q & (kAccRing - 1)), ((q / kAccRing) & 1) ^ 1); }
                        if (m == 0 && lane == 0) TRACE(2, Qb + 2 * (t / gridDim.x) + j);
                        tc_fence_after();
                        // plane groups g = 0,1,2 <-> item-local output planes j-2, j-1, j; ring slots r0, r0+1, r0+2 (mod kAccRing)
                        const bool v0 = j >= 2, v1 = j >= 1 && j - 1 < k.nsteps, v2 = j < k.nsteps;
                        const int r0 = (Qb + j - 2) & (kAccRing - 1), r1 = (r0 + 1) & (kAccRing - 1), r2 = (r0 + 2) & (kAccRing - 1);
                        const bool c01 = v0 && v1 && r1 == r0 + 1, c12 = v1 && v2 && r2 == r1 + 1;   // adjacent in TMEM (no ring wrap)
                        const uint32_t dg0 = d_m + (uint32_t)(r0 * kPG), dg1 = d_m + (uint32_t)(r1 * kPG), dg2 = d_m + (uint32_t)(r2 * kPG);
                        const uint32_t a_slot = slots_addr + (uint32_t)st * p.slot_bytes;
                        uint32_t b_ent = (b_addr >> 4) | kb_lbo;
                        bool first = true;
                        for (int e = 0; e < nent; ++e, b_ent += btile16) {
                            const uint4 en = etab[e];                      // x = A byte offset, y = LBO field
                            uint32_t a_lo = ((a_slot + en.x) >> 4) | en.y, b_lo = b_ent;
                            for (int ks = 0; ks < ksteps; ++ks, a_lo += a_kstep, b_lo += kb_kstep, first = false) {
                                const uint64_t ad = desc_hi | a_lo;
                                const bool join12 = c12 && !first;          // group 2 may ride along once it has been overwritten
                                // segment starting at group 0
                                if (v0) {
                                    if (c01 && join12) umma_f16_elect(dg0, ad, desc_hi | b_lo, id3, 1u);
                                    else if (c01) umma_f16_elect(dg0, ad, desc_hi | b_lo, id2, 1u);
                                    else umma_f16_elect(dg0, ad, desc_hi | b_lo, id1, 1u);
                                }
                                // segment starting at group 1 (when it did not ride with group 0)
                                if (v1 && !c01) {
                                    if (join12) umma_f16_elect(dg1, ad, desc_hi | (b_lo + kPG), id2, 1u);
                                    else umma_f16_elect(dg1, ad, desc_hi | (b_lo + kPG), id1, 1u);
                                }
                                // group 2 alone: always in the step's first MMA (overwrite), later only across a ring wrap
                                if (v2 && !join12) umma_f16_elect(dg2, ad, desc_hi | (b_lo + 2 * kPG), id1, first ? 0u : 1u);
                            }
                        }
#ifdef MVS_TC_TRACE2
                        if (m == 0 && lane == 0) TRACE(6, Qb + 2 * (t / gridDim.x) + j);
#endif
                        umma_commit_elect(slot_empty + st);
                        if (j >= 2) umma_commit_elect(acc_full + ((Qb + j - 2) & (kAccRing - 1)));      // output plane j-2 has all three contributions
#ifdef MVS_DIAG_EXTRA_COMMITS   // diagnostic: what does a tcgen05.commit cost?  (dummy barrier nobody waits on)
                        umma_commit_elect(b_empty + kBStages - 1); umma_commit_elect(b_empty + kBStages - 1);
#endif
                        if (m == 0 && lane == 0) TRACE(3, Qb + 2 * (t / gridDim.x) + j);
                        if (++st == stages) { st = 0; sph ^= 1; }
                    }
                    Qb += k.nsteps;
                }
            } else {
                int st0 = 0, ph0 = 0;          // ring stage / phase of the first live slot of the current step
                int I = 0, bst = 0, bph = 0;   // steps done; weight ring position (streaming)
                for (int t = blockIdx.x; t < p.nwork; t += gridDim.x) {
                    const Work k = decode(t);
                    for (int i = 0; i < k.nsteps; ++i, ++I) {
                        const int buf = I & 1;
                        if (m == 0 && lane == 0) TRACE(1, I);
                        mbar_wait(acc_empty + buf, ((I >> 1) & 1) ^ 1);
                        // live slots: stages st0, st0+1, .. (mod stages); only the newest `sps` (all of them at i = 0) can be unready
                        uint32_t sa0 = 0, sa1 = 0, sa2 = 0;
                        {
                            int s2 = st0, p2 = ph0;
                            for (int l = 0; l < p.live; ++l) {
                                if (i == 0 || l >= p.live - p.sps) mbar_wait(slot_full + s2, (uint32_t)p2);
                                const uint32_t sa = slots_addr + (uint32_t)s2 * p.slot_bytes;
                                if (l == 0) sa0 = sa; else if (l == 1) sa1 = sa; else sa2 = sa;
                                if (++s2 == stages) { s2 = 0; p2 ^= 1; }
                            }
                        }
                        tc_fence_after();
                        if (m == 0 && lane == 0) TRACE(2, I);
                        const uint32_t d_tmem0 = tmem_base + (uint32_t)((buf * p.groups * p.nM + m) * p.N), d_group = (uint32_t)(p.nM * p.N);
                        uint32_t b_ent = (b_addr >> 4) | b_lbo;
                        for (int e = 0; e < nent; ++e, b_ent += btile16) {
                            const uint4 en = etab[e];                      // x = A byte offset, y = LBO field, z = slot, w = first | group << 8
                            const uint32_t d_tmem = d_tmem0 + (en.w >> 8) * d_group;
                            uint32_t b_lo = b_ent;
                            if (!p.b_resident) {
                                mbar_wait(b_full + bst, (uint32_t)bph);
                                tc_fence_after();
                                b_lo = ((b_addr + (uint32_t)bst * p.btile_bytes) >> 4) | b_lbo;
                            }
                            const uint32_t sa = en.z == 0 ? sa0 : (en.z == 1 ? sa1 : sa2);
                            uint32_t a_lo = ((sa + en.x) >> 4) | en.y;
                            uint32_t acc = (en.w & 1u) ? 0u : 1u;
                            for (int ks = 0; ks < ksteps; ++ks, a_lo += a_kstep, b_lo += b_kstep, acc = 1u)
                                umma_f16_elect(d_tmem, desc_hi | a_lo, desc_hi | b_lo, idesc, acc);
                            if (!p.b_resident) { umma_commit_elect(b_empty + bst); if (++bst == p.bstages) { bst = 0; bph ^= 1; } }
                        }
                        // slots no later step of this item reads retire with these MMAs (all remaining ones after the last step)
                        const int nrel = (i == k.nsteps - 1) ? p.live : p.sps;
                        for (int l = 0; l < nrel; ++l) { umma_commit_elect(slot_empty + st0); if (++st0 == stages) { st0 = 0; ph0 ^= 1; } }
                        umma_commit_elect(acc_full + buf);
                        if (m == 0 && lane == 0) TRACE(3, I);
                    }
                }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> affine / ReLU / skip -> C8 store =====================
        const int quad = warp & 3;                 // TMEM lanes [32 quad, 32 quad + 32) belong to this warp
        const int mpar = (warp - 6) >> 2;          // the warps of a quadrant take M-tiles mpar, mpar + kMStride, ...
        const int CoB = (p.Cout + 7) / 8;
        const int64_t HWo = (int64_t)p.Ho * p.Wo;
        const bool kwfold = p.kwfold != 0;
        int I = 0;
        for (int t = blockIdx.x; t < p.nwork; t += gridDim.x) {
            const Work k = decode(t);
            const int b = k.b, w0 = k.w0, h0 = k.h0, d0 = k.d0;
            // accumulator row (= lane) of this warp's M-tiles: one h-row of the tile per warp, lane = w position.  Validity and the
            // output address base are per work item, not per depth step (the 64-bit address math used to cost ~200 cycles a step).
            bool valid_u[2];
            int64_t base_u[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int m = mpar + kMStride * u;
                const int r = m * 128 + quad * 32 + lane;
                const int hh = r >> 5, ww = r & 31;
                const int64_t mul = p.mode == MODE_T2 ? 2 : 1;       // transposed stride 2: the row's 2x2x2 output block starts at twice the input position
                base_u[u] = (int64_t)b * p.ys_b + mul * (h0 + hh) * p.ys_h + mul * (w0 + ww) + p.y_org;
                valid_u[u] = (m < p.nM) && (hh < p.TH) && (ww < kTW) && (h0 + hh < p.Ht) && (w0 + ww < p.Wt);
            }
            auto row_at = [&](int u, uint32_t trow) {
                RowAt a;
                a.trow = trow; a.base = base_u[u]; a.valid = valid_u[u];
                return a;
            };
            for (int i = 0; i < k.nsteps; ++i, ++I) {
                if (p.kdfold) {    // output plane I sits in ring slot I % kAccRing of every M-tile (kPG columns, kw blocks at +0, +8, +16)
                    const int rsl = I % kAccRing;
                    mbar_wait(acc_full + rsl, (I / kAccRing) & 1);
                    if (warp == 6 && lane == 0) TRACE(4, I);
                    tc_fence_after();
                    {
                        const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16);
                        RowAt ra[2];
                        for (int u = 0; u < 2; ++u) {
                            const int m = mpar + kMStride * u;
                            ra[u] = row_at(u, tq + (uint32_t)((m * kAccRing + rsl) * kPG));
                        }
#ifndef MVS_DIAG_NO_EPI   // diagnostic: with the epilogue body removed, what bounds a step?
                        if (kMStride == 2 && mpar + 2 < p.nM) epilogue_rows<T, 1, true, 2>(p, aff, ra, b, d0 + i, CoB, HWo);
                        else if (mpar < p.nM) { const RowAt r1[1] = {ra[0]}; epilogue_rows<T, 1, true, 1>(p, aff, r1, b, d0 + i, CoB, HWo); }
#endif
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + rsl);
                    if (warp == 6 && lane == 0) TRACE(5, I);
                    continue;
                }
                const int buf = I & 1;
                if (p.skip) {   // next step's skip rows (this step's at i = 0 as well) go to L2 while the MMAs run
                    for (int ii = (i == 0 ? 0 : i + 1); ii <= i + 1 && ii < k.nsteps; ++ii)
                        for (int m = mpar; m < p.nM; m += kMStride) {
                            const int r = m * 128 + quad * 32 + lane;
                            const int hh = r >> 5, ww = r & 31;
                            const bool valid = (hh < p.TH) && (ww < kTW) && (h0 + hh < p.Ht) && (w0 + ww < p.Wt);
                            if (p.mode == MODE_T2) prefetch_skip_rows<T, 8>(p, valid, b, 2 * (d0 + ii), 2 * (h0 + hh), 2 * (w0 + ww), CoB, HWo);
                            else prefetch_skip_rows<T, 1>(p, valid, b, d0 + ii, h0 + hh, w0 + ww, CoB, HWo);
                        }
                }
                mbar_wait(acc_full + buf, (I >> 1) & 1);
                if (warp == 6 && lane == 0) TRACE(4, I);
                tc_fence_after();
                for (int g = 0; g < p.groups; ++g) {
                    // accumulator group g = image plane g of a two-plane 2-D step (output plane 2 (d0 + i) + g); one group otherwise
                    const int od = p.groups == 1 ? d0 + i : p.groups * (d0 + i) + g;
                    if (p.groups > 1 && od >= p.Do) continue;            // odd image count: the last step's second plane is padding
                    const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16);
                    RowAt ra[2];
                    for (int u = 0; u < 2; ++u) {
                        const int m = mpar + kMStride * u;
                        ra[u] = row_at(u, tq + (uint32_t)(((buf * p.groups + g) * p.nM + m) * p.N));
                    }
                    const bool two = kMStride == 2 && mpar + 2 < p.nM;
                    if (p.mode == MODE_T2) {
                        for (int u = 0; u < (kMStride == 2 ? 2 : 1) && mpar + kMStride * u < p.nM; ++u) {
                            const RowAt r1[1] = {ra[u]};
                            epilogue_rows<T, 8, false, 1>(p, aff, r1, b, 2 * (d0 + i), CoB, HWo);
                        }
                    }
#ifndef MVS_DIAG_NO_EPI
                    else if (mpar < p.nM) {
                        const RowAt r1[1] = {ra[0]};
                        if (kwfold) { if (two) epilogue_rows<T, 1, true, 2>(p, aff, ra, b, od, CoB, HWo); else epilogue_rows<T, 1, true, 1>(p, aff, r1, b, od, CoB, HWo); }
                        else { if (two) epilogue_rows<T, 1, false, 2>(p, aff, ra, b, od, CoB, HWo); else epilogue_rows<T, 1, false, 1>(p, aff, r1, b, od, CoB, HWo); }
                    }
#endif
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + buf);
                if (warp == 6 && lane == 0) TRACE(5, I);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ weight tiles
// One tile per program entry: wt[entry][kchunk][n][8], n = block * CoP + co.  `src` lists, per (entry, kchunk, column
// block), which tap of the gather form G[27][Cin][CoutPad] (and which 8 input channels) it holds; -1 = zeros.
struct TileSrc { int8_t tap[kMaxEntries][8][12]; int8_t cib[kMaxEntries][8]; };

template <typename T>
__global__ void pack_tiles_kernel(const float* __restrict__ g, T* __restrict__ w, const __grid_constant__ TileSrc src, int nentries,
                                  int kchunks, int N, int Cin, int Cout, int CoutPad, int CoP, int nblk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over nentries * kchunks * N
    if (i >= nentries * kchunks * N) return;
    const int n = i % N, kc = (i / N) % kchunks, e = i / (N * kchunks);
    const int blk = n / CoP, co = n % CoP;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.f;
    if (blk < nblk && co < Cout) {
        const int tap = src.tap[e][kc][blk];
        if (tap >= 0) {
            const int ci0 = src.cib[e][kc] * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = g[((int64_t)tap * Cin + ci0 + k) * CoutPad + co];
        }
    }
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

// ------------------------------------------------------------------------------------------------ tap programs
struct Plan { TcParams p; TileSrc src; size_t smem; int stages_chosen; };

// What a launch convolves: a 3-D layer (mvs_conv3d_desc as is), or a stack of M images treated as a volume whose "depth" axis
// is the image index (two_d: Din = Dout = M, one live plane per step, no taps across images) with a ksize x ksize filter:
// 3x3 stride 1 (kw folded into N like the 3-D stride-1 program) or 5x5 stride 2 (parity-staged like the 3-D stride-2 one).
// out_pad: write the zero-bordered image-major C8P layout the plane-sweep gather reads instead of the stack layout.
struct Spec : mvs_conv3d_desc { int two_d, ksize, out_pad; float slope; };

int mode_of(const Spec* d) { return d->stride == 1 ? MODE_S1 : (d->transposed ? MODE_T2 : MODE_S2); }

int cop_of(const Spec* d) { return d->Cout == 1 ? 8 : d->Cout; }
// kd-fold (stride 1, <= 8 output channels): the 9 (kd, kw) taps of one kh share an MMA; see the issuer.  MVS_TC_KDFOLD=0 disables.
bool kdfold_of(const Spec* d) {
    return !d->two_d && mode_of(d) == MODE_S1 && cop_of(d) == 8 && mvs_knob(MVS_KNOB_TC_KDFOLD, 1) != 0;
}
// Stride-1 programs fold the 3 kw taps into N (3 column blocks) unless that makes the accumulators so wide that only one
// M-tile fits in TMEM (2-D layers with 64 output channels): those run 9 entries with shifted A views, like the stride-2 program.
// (2-D layers with 32 output channels as well: measured 1.02 -> 0.98 ms over FeatureNet, the folded epilogue's three TMEM loads and
// shuffles per channel block cost more than the extra MMAs; knob tc_kwfold_max = widest folded N)
bool kwfold_of(const Spec* d) { return mode_of(d) == MODE_S1 && !(d->two_d && 3 * cop_of(d) > mvs_knob(MVS_KNOB_TC_KWFOLD_MAX, 48)); }
int nblk_of(const Spec* d) { return mode_of(d) == MODE_S1 ? (kdfold_of(d) ? 12 : (kwfold_of(d) ? 3 : 1)) : (mode_of(d) == MODE_T2 ? 8 : 1); }
int n_of(const Spec* d) { return (nblk_of(d) * cop_of(d) + 15) / 16 * 16; }
// Thin 2-D layers (N <= 32: the 8-channel full-resolution layers of FeatureNet) take TWO image planes per step: a step costs
// ~1 k cycles of barrier waits / commits and ~1 k of epilogue chain whatever it computes, and 2 x 2 x 4 M-tiles x 32 columns
// still fit TMEM.  The 5x5 program only qualifies with Cin = 8 (13 paired entries per plane; 2 x 25 would not fit the table).
int planes_per_step(const Spec* d) {
    if (mvs_knob(MVS_KNOB_TC_PLANES, 2) == 1) return 1;        // test / tuning knob: 1 disables
    return (d->two_d && n_of(d) <= 32 && (d->ksize == 3 || d->Cin == 8)) ? 2 : 1;
}

// A "fold tap": one A view (slot, sub-plane, row shift) and, per column block, the filter tap it multiplies (-1 = none).
struct FoldTap { int slot, sub, shift, grp, tap[12]; };   // grp: accumulator group (image plane of a two-plane 2-D step)

// Build the entry list + weight-tile sources.  Returns the number of entries.
int build_program(const Spec* d, TcParams& p, TileSrc& src) {
    const int mode = mode_of(d);
    memset(&src, -1, sizeof(src));
    FoldTap ft[54];
    memset(ft, 0, sizeof(ft));
    int nft = 0;
    if (d->two_d && mode == MODE_S1 && !kwfold_of(d)) {
        // 3x3 over one image plane, one entry per tap: row shift kh * 32 + kw from the tile origin (o - 1)
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                FoldTap& t = ft[nft++];
                t.slot = 0; t.sub = 0; t.shift = kh * kPW + kw;
                for (int c = 0; c < 12; ++c) t.tap[c] = c == 0 ? kh * 3 + kw : -1;
            }
    } else if (d->two_d && mode == MODE_S1) {
        // 3x3 over one image plane: one fold tap per kh, column block c = input column offset c; gather form G[kh * 3 + kw]
        for (int kh = 0; kh < 3; ++kh) {
            FoldTap& t = ft[nft++];
            t.slot = 0; t.sub = 0; t.shift = kh * kPW;
            for (int c = 0; c < 12; ++c) t.tap[c] = c < 3 ? kh * 3 + c : -1;
        }
    } else if (d->two_d) {
        // 5x5 stride 2, pad 2: out o reads x[2o - 2 + k] = parity (k & 1) plane at index o - 1 + (k >> 1); the tile origin is o - 1
        for (int kh = 0; kh < 5; ++kh)
            for (int kw = 0; kw < 5; ++kw) {
                FoldTap& t = ft[nft++];
                t.slot = 0; t.sub = (kh & 1) * 2 + (kw & 1); t.shift = (kh >> 1) * kPW + (kw >> 1);
                for (int c = 0; c < 12; ++c) t.tap[c] = c == 0 ? kh * 5 + kw : -1;
            }
    } else if (mode == MODE_S1 && kdfold_of(d)) {
        // one fold tap per kh; column block g * 4 + c: plane group g (output plane p - 1 + g of input plane p) x input column offset c
        for (int kh = 0; kh < 3; ++kh) {
            FoldTap& t = ft[nft++];
            t.slot = 0; t.sub = 0; t.shift = (d->transposed ? 2 - kh : kh) * kPW;
            for (int blk = 0; blk < 12; ++blk) {
                const int g = blk >> 2, c = blk & 3;
                const int kd = d->transposed ? g : 2 - g;       // Conv3d: out q reads x[q - 1 + kd]; transposed: x[q + 1 - kd]
                t.tap[blk] = c < 3 ? (kd * 3 + kh) * 3 + (d->transposed ? 2 - c : c) : -1;
            }
        }
    } else if (mode == MODE_S1) {
        // gather form: Conv3d reads x[o - 1 + k]; stride-1 ConvTranspose3d reads x[o + 1 - k].  Block c = input column offset c.
        for (int kd = 0; kd < 3; ++kd)
            for (int kh = 0; kh < 3; ++kh) {
                FoldTap& t = ft[nft++];
                t.slot = d->transposed ? 2 - kd : kd; t.sub = 0; t.shift = (d->transposed ? 2 - kh : kh) * kPW;
                for (int c = 0; c < 12; ++c) t.tap[c] = c < 3 ? (kd * 3 + kh) * 3 + (d->transposed ? 2 - c : c) : -1;
            }
    } else if (mode == MODE_S2) {
        // x[2o - 1 + k]: k = 1 -> even parity at index o (shift 1 from the tile origin o-1), k = 0 / 2 -> odd parity at o-1 / o
        const int sh[3] = {0, 1, 1};
        for (int tp = 0; tp < 27; ++tp) {
            const int k[3] = {tp / 9, (tp / 3) % 3, tp % 3};
            FoldTap& t = ft[nft++];
            t.slot = k[0]; t.sub = (k[1] == 1 ? 0 : 1) * 2 + (k[2] == 1 ? 0 : 1); t.shift = sh[k[1]] * kPW + sh[k[2]];
            for (int c = 0; c < 12; ++c) t.tap[c] = c == 0 ? tp : -1;
        }
    } else {
        // output parity 0 <- (offset 0, k = 1); parity 1 <- (offset 0, k = 2), (offset 1, k = 0); block c = (pd, ph, pw)
        const int kk[2][2] = {{1, -1}, {2, 0}};   // kk[parity][offset]
        for (int od = 0; od < 2; ++od)
            for (int oh = 0; oh < 2; ++oh)
                for (int ow = 0; ow < 2; ++ow) {
                    FoldTap& t = ft[nft++];
                    t.slot = od; t.sub = 0; t.shift = oh * kPW + ow;
                    for (int c = 0; c < 12; ++c) {
                        if (c >= 8) { t.tap[c] = -1; continue; }
                        const int kd = kk[c >> 2][od], kh = kk[(c >> 1) & 1][oh], kw = kk[c & 1][ow];
                        t.tap[c] = (kd < 0 || kh < 0 || kw < 0) ? -1 : (kd * 3 + kh) * 3 + kw;
                    }
                }
    }
    if (planes_per_step(d) == 2) {
        // two image planes per step: the same taps again on the second live slot, accumulating into the second group
        for (int i = 0; i < nft; ++i) { ft[nft + i] = ft[i]; ft[nft + i].slot = 1; ft[nft + i].grp = 1; }
        nft *= 2;
    }
    int ne = 0;
    bool seen[2] = {false, false};
    auto add = [&](const FoldTap& t, int shift, int lbo_rows) {
        Entry& e = p.prog[ne];
        e.row_shift = (int16_t)shift; e.lbo_rows = (uint16_t)lbo_rows; e.slot_off = (uint8_t)t.slot; e.sub = (uint8_t)t.sub;
        e.group = (uint8_t)t.grp; e.first = seen[t.grp] ? 0 : 1;
        seen[t.grp] = true;
        return ne++;
    };
    if (d->Cin != 8) {
        for (int i = 0; i < nft; ++i) {
            const int e = add(ft[i], ft[i].shift, 0);
            for (int kc = 0; kc < d->Cin / 8; ++kc) { for (int c = 0; c < 12; ++c) src.tap[e][kc][c] = (int8_t)ft[i].tap[c]; src.cib[e][kc] = (int8_t)kc; }
        }
        return ne;
    }
    // Cin = 8: two fold taps of the same (slot, sub-plane) share one K = 16 step; LBO = their row distance
    bool used[54] = {false};
    for (int a = 0; a < nft; ++a) {
        if (used[a]) continue;
        used[a] = true;
        int best = -1;
        for (int c = a + 1; c < nft; ++c)
            if (!used[c] && ft[c].slot == ft[a].slot && ft[c].sub == ft[a].sub && ft[c].shift != ft[a].shift &&
                (best < 0 || abs(ft[c].shift - ft[a].shift) < abs(ft[best].shift - ft[a].shift))) best = c;
        int e;
        if (best >= 0) {
            used[best] = true;
            const int lo = ft[best].shift < ft[a].shift ? best : a, hi = lo == a ? best : a;
            e = add(ft[lo], ft[lo].shift, ft[hi].shift - ft[lo].shift);
            for (int c = 0; c < 12; ++c) { src.tap[e][0][c] = (int8_t)ft[lo].tap[c]; src.tap[e][1][c] = (int8_t)ft[hi].tap[c]; }
        } else if (ft[a].shift > 0) {      // lone tap: zero weights on the row before it
            e = add(ft[a], ft[a].shift - 1, 1);
            for (int c = 0; c < 12; ++c) src.tap[e][1][c] = (int8_t)ft[a].tap[c];
        } else {                            // lone tap at shift 0: zero weights on the row after it
            e = add(ft[a], 0, 1);
            for (int c = 0; c < 12; ++c) src.tap[e][0][c] = (int8_t)ft[a].tap[c];
        }
        src.cib[e][0] = 0; src.cib[e][1] = 0;
    }
    return ne;
}

bool make_plan(const Spec* d, Plan& pl) {
    TcParams& p = pl.p;
    memset(&p, 0, sizeof(p));
    p.mode = mode_of(d);
    p.B = d->B; p.CiB = d->Cin / 8; p.Cout = d->Cout; p.CoP = cop_of(d); p.nblk = nblk_of(d); p.N = n_of(d);
    p.ksteps = d->Cin == 8 ? 1 : d->Cin / 16;
    p.kchunks = 2 * p.ksteps;
    p.Di = d->Din; p.Hi = d->Hin; p.Wi = d->Win; p.Do = d->Dout; p.Ho = d->Hout; p.Wo = d->Wout;
    if (p.mode == MODE_T2) { p.Dt = p.Di; p.Ht = p.Hi; p.Wt = p.Wi; } else { p.Dt = p.Do; p.Ht = p.Ho; p.Wt = p.Wo; }
    p.relu = d->relu; p.is_bf16 = d->dtype_in == MVS_BF16;
    p.slope = d->slope; p.kwfold = kwfold_of(d) ? 1 : 0;
    const int npl = planes_per_step(d);                         // 2-D: image planes per step (1 or 2)
    p.nsub = p.mode == MODE_S2 ? 4 : 1;
    p.sps = d->two_d ? npl : (p.mode == MODE_S2 ? 2 : 1);
    p.live = d->two_d ? npl : (p.mode == MODE_T2 ? 2 : 3);
    p.d_mul = d->two_d ? npl : (p.mode == MODE_S2 ? 2 : 1);
    p.d_org = (d->two_d || p.mode == MODE_T2) ? 0 : -1;
    p.groups = npl;
    if (npl == 2) p.Dt = (p.Do + 1) / 2;                        // the tiles walk steps of two image planes
    p.kdfold = kdfold_of(d) ? 1 : 0;
    // output addressing (voxels): C8 volume [B][CoB][Do][Ho][Wo], plain [B][Do][Ho][Wo] for one channel, or (out_pad) the
    // zero-bordered image-major C8P maps [Do = image][CoB][Ho + 3][Wo + 2] with pixel (0, 0) at row 1, column 1
    {
        const int64_t CoB = (d->Cout + 7) / 8;
        if (d->out_pad) {
            p.ys_h = p.Wo + 2; p.ys_cb = (int64_t)(p.Ho + 3) * p.ys_h; p.ys_d = CoB * p.ys_cb; p.ys_b = 0; p.y_org = p.ys_h + 1;
        } else {
            p.ys_h = p.Wo; p.ys_d = (int64_t)p.Ho * p.Wo; p.ys_cb = d->Cout == 1 ? 0 : (int64_t)p.Do * p.ys_d;
            p.ys_b = (d->Cout == 1 ? 1 : CoB) * (int64_t)p.Do * p.ys_d; p.y_org = 0;
        }
    }
    // ring depth = live slots + the slots of one step prefetched while the current step computes
    p.stages = d->two_d ? 2 * npl : (p.kdfold ? 3 : (p.mode == MODE_S1 ? 4 : (p.mode == MODE_T2 ? 3 : 4)));   // minimum; make_plan adds what fits
    p.nentries = build_program(d, p, pl.src);
    p.btile_bytes = (uint32_t)p.kchunks * p.N * 16;
    int max_reach = 0;   // furthest row an A descriptor touches beyond its 128-row window
    for (int e = 0; e < p.nentries; ++e) max_reach = max(max_reach, (int)p.prog[e].row_shift + (int)p.prog[e].lbo_rows);
    const int halo = (p.mode == MODE_S1 || d->two_d) ? 2 : 1;
    const uint32_t all_b = (uint32_t)p.nentries * p.btile_bytes;
    // largest tile (nM M-tiles of 128 rows = 4 nM x 30 positions) whose slot ring fits beside the weights with 2 accumulator
    // sets in TMEM -- but small volumes take smaller tiles so that at least ~2 waves of CTAs exist
    bool found = false;
    const int forced = mvs_knob(MVS_KNOB_TC_NM, 0);           // test knob: force the M-tile count (when it fits)
    for (int nM = 4; nM >= 1 && !found; --nM) {
        if ((p.kdfold ? nM * kAccRing * kPG : 2 * p.groups * nM * p.N) > 512) continue;
        if (forced >= 1 && forced <= 4 && nM > forced) continue;
        p.nM = nM; p.TH = 4 * nM; p.PH = p.TH + halo;
        const int rows = max(nM * 128 + max_reach + 1, p.PH * kPW);
        p.chunk_bytes = (uint32_t)((rows + 7) / 8 * 8) * 16u;
        p.sub_bytes = p.chunk_bytes * (uint32_t)p.CiB;
        p.slot_bytes = p.sub_bytes * (uint32_t)p.nsub;
        const int min_stages = p.stages;                       // live slots + one step of prefetch
        const size_t ring = (size_t)min_stages * p.slot_bytes;
        const size_t room = ring + kTail <= (size_t)kSmemLimit ? (size_t)kSmemLimit - ring - kTail : 0;
        size_t bbytes;
        if (room >= all_b) { p.b_resident = 1; p.bstages = 1; bbytes = all_b; }
        else if (room >= 4 * (size_t)p.btile_bytes) {
            // streamed weights: half of what is left (at least 4 tiles) for the weight ring, the rest for deeper prefetch
            p.b_resident = 0; p.bstages = (int)min((size_t)kBStages, max((size_t)4, room / 2 / p.btile_bytes)); bbytes = (size_t)p.bstages * p.btile_bytes;
        } else continue;
        // The input ring is what hides HBM latency (one slot = one plane of the tile): the kernel was latency-bound with a
        // single slot in flight per SM (0.8 TB/s), so every byte of shared memory left over goes to more slots in flight.
        int stages_fit = (int)(((size_t)kSmemLimit - kTail - bbytes) / p.slot_bytes);
        if (stages_fit > kMaxStages) stages_fit = kMaxStages;
        // ... up to a point: the ring competes with the L1 for the same 256 KB, and the epilogue's skip loads / stores go through
        // the L1.  Measured per layer at the headline sizes (tools/layer_time.py tc_stages=N, 8 items, us): transposed stride 2
        // conv11 415 (12 slots) -> 336 (3), conv9 142 -> 123; kd-folded conv0 750 -> 727 (3); stride 1 conv4 74 -> 60 (4-5),
        // conv2 171 -> 150 (5) / 138 (8); stride 2 conv1 220 -> 212 (5) but 295 at its minimum of 4; the 2-D feature layers
        // 1.11 -> 1.02 ms per step (4).  So: the minimum for the transposed and kd-folded programs, 5 slots otherwise, 4 in 2-D.
        {
            const int cap = (p.mode == MODE_T2 || p.kdfold) ? min_stages : (d->two_d ? max(min_stages, 4) : max(min_stages, 5));
            if (stages_fit > cap) stages_fit = cap;
        }
        const int fs = mvs_knob(MVS_KNOB_TC_STAGES, 0);        // test / tuning knob
        if (fs >= min_stages && fs <= stages_fit) stages_fit = fs;
        pl.stages_chosen = stages_fit;
        pl.smem = (size_t)stages_fit * p.slot_bytes + bbytes + kTail;
        p.nwt = (p.Wt + kTW - 1) / kTW;
        p.nht = (p.Ht + p.TH - 1) / p.TH;
        const int64_t tiles_nm = (int64_t)p.B * p.nwt * p.nht;
        // (stride 2: 4 waves -- with 2, MVSNet's conv3 took 16 x 30 tiles and 158 us where 8 x 30 tiles take 113 us; the other
        // programs measured equal or slower with the smaller tile)
        found = nM == 1 || nM == forced || tiles_nm * ((p.Dt + 3) / 4) >= (p.mode == MODE_S2 ? 4 : 2) * 148;
    }
    if (!found) return false;
    p.stages = pl.stages_chosen;
    uint32_t cols = 32;
    while ((int)cols < (p.kdfold ? p.nM * kAccRing * kPG : 2 * p.groups * p.nM * p.N)) cols <<= 1;
    p.tmem_cols = cols;
    // depth segments: enough CTAs for >= ~3 waves of 148 SMs, but >= 4 steps each (halo slots are reloaded per segment)
    const int64_t tiles = (int64_t)p.B * p.nwt * p.nht;
    // Work items = tiles x depth segments, dealt round-robin to the persistent CTAs.  The kernel time is the busiest CTA's
    // rounds x (steps per item + the halo planes every item reloads + ~1 step of pipeline bubble), so pick the segment count
    // that minimises exactly that: e.g. 48 tiles x 12 segments = 576 items fill 148 CTAs to 97 % in 4 rounds, where a
    // power-of-two split (16 segments, 768 items) leaves the 6th round 19 % full.
    const int min_ld = p.kdfold ? 4 : (d->two_d ? 1 : 2);
    const int extra = (d->two_d ? 0 : (p.mode == MODE_S1 ? 2 : 1)) + 1;
    const int fseg = mvs_knob(MVS_KNOB_TC_NSEG, 0);          // test / tuning knob
    int best_nseg = 1;
    int64_t best_cost = -1;
    for (int nseg = 1; nseg <= max(1, p.Dt / min_ld); ++nseg) {
        const int ld = (p.Dt + nseg - 1) / nseg;
        if ((int64_t)(nseg - 1) * ld >= p.Dt) continue;          // an empty last segment
        const int64_t rounds = (tiles * nseg + 147) / 148;
        const int64_t cost = rounds * (ld + extra);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_nseg = nseg; }
    }
    if (fseg >= 1 && fseg <= p.Dt) best_nseg = fseg;
    p.LD = (p.Dt + best_nseg - 1) / best_nseg; p.nseg = (p.Dt + p.LD - 1) / p.LD;
    return tiles * p.nseg < (1ll << 31);
}

}  // namespace

static Spec spec3d(const mvs_conv3d_desc* d) {
    Spec sp;
    static_cast<mvs_conv3d_desc&>(sp) = *d;
    sp.two_d = 0; sp.ksize = 3; sp.out_pad = 0; sp.slope = d->relu ? 0.f : 1.f;
    return sp;
}

// A stack of M images as a volume: B = 1, D = M.
static Spec spec2d(const mvs_conv2d_desc* d) {
    Spec sp;
    memset(&sp, 0, sizeof(sp));
    sp.B = 1; sp.Cin = d->Cin; sp.Cout = d->Cout;
    sp.Din = d->M; sp.Hin = d->Hin; sp.Win = d->Win; sp.Dout = d->M; sp.Hout = d->Hout; sp.Wout = d->Wout;
    sp.stride = d->stride; sp.transposed = 0; sp.dtype_in = d->dtype; sp.dtype_out = d->dtype; sp.relu = d->relu; sp.algo = d->ws_packed ? 3 : 2;
    sp.two_d = 1; sp.ksize = d->ksize; sp.out_pad = d->out_padded;
    sp.slope = d->relu == 1 ? 0.f : (d->relu == 2 ? d->leaky_slope : 1.f);
    return sp;
}

static int tc_supported(const Spec* d) {
    if (d->dtype_in != MVS_F16 && d->dtype_in != MVS_BF16) return 0;
    if (d->Cout != 1 && d->dtype_out != d->dtype_in) return 0;
    if (d->Cin != 8 && (d->Cin % 16 != 0 || d->Cin > 64)) return 0;
    if (d->Cout != 1 && (d->Cout % 8 != 0 || d->Cout > 64)) return 0;
    if (mode_of(d) == MODE_T2 && (d->Cout == 1 || 8 * d->Cout > 256)) return 0;
    if (mode_of(d) == MODE_S2 && ((d->Win & 1) || (d->Hin & 1) || (!d->two_d && (d->Din & 1)))) return 0;
    if (d->two_d && !((d->ksize == 3 && d->stride == 1) || (d->ksize == 5 && d->stride == 2))) return 0;
    if (d->two_d && d->Cout == 1) return 0;
    return 1;
}

static int64_t tc_workspace_bytes(const Spec* d) {
    if (!tc_supported(d)) return 0;
    const int kchunks = d->Cin == 8 ? 2 : d->Cin / 8;
    // upper bounds on the entry count (Cin = 8 pairs need fewer)
    const int nentries = d->two_d ? planes_per_step(d) * (d->ksize == 3 ? (kwfold_of(d) ? 3 : 9) : 25)
                                  : (mode_of(d) == MODE_S1 ? 9 : (mode_of(d) == MODE_T2 ? 8 : 27));
    return (int64_t)nentries * kchunks * n_of(d) * 16;  // weight tiles [entry][kchunk][N][8] in the storage dtype
}

static int conv_fwd_tc(const Spec* d, const void* x, const float* g, const float* scale, const float* shift,
                       const void* skip, void* y, void* ws, void* stream);

int mvs_conv3d_tc_supported(const mvs_conv3d_desc* d) { const Spec sp = spec3d(d); return tc_supported(&sp); }
int64_t mvs_conv3d_tc_workspace_bytes(const mvs_conv3d_desc* d) { const Spec sp = spec3d(d); return tc_workspace_bytes(&sp); }
int mvs_conv3d_fwd_tc(const mvs_conv3d_desc* d, const void* x, const float* g, const float* scale, const float* shift,
                      const void* skip, void* y, void* ws, void* stream) {
    const Spec sp = spec3d(d);
    return conv_fwd_tc(&sp, x, g, scale, shift, skip, y, ws, stream);
}

int64_t mvs_conv2d_tc_workspace_bytes(const mvs_conv2d_desc* d) { const Spec sp = spec2d(d); return tc_workspace_bytes(&sp); }
int mvs_conv2d_fwd_tc(const mvs_conv2d_desc* d, const void* x, const float* g, const float* scale, const float* shift, void* y,
                      void* ws, void* stream) {
    const Spec sp = spec2d(d);
    MVS_REQUIRE(tc_supported(&sp), MVS_E_UNSUPPORTED,
                "mvs_conv2d_fwd: needs fp16/bf16 storage, Cin in {8,16,32,48,64}, Cout in {8..64}, 3x3 stride 1 or 5x5 stride 2 (even H, W)");
    return conv_fwd_tc(&sp, x, g, scale, shift, nullptr, y, ws, stream);
}

static int conv_fwd_tc(const Spec* d, const void* x, const float* g, const float* scale, const float* shift,
                       const void* skip, void* y, void* ws, void* stream) {
    MVS_REQUIRE(tc_supported(d), MVS_E_UNSUPPORTED,
                "mvs_conv3d_fwd: tcgen05 path needs fp16/bf16 storage, Cin in {8,16,32,48,64}, Cout in {1,8..64} (<= 32 transposed stride 2)");
    MVS_REQUIRE(ws, MVS_E_ARG, "mvs_conv3d_fwd: the tcgen05 path needs a workspace of mvs_conv3d_workspace_bytes() bytes");
    EncodeTiledFn enc = encode_tiled();
    MVS_REQUIRE(enc, MVS_E_LAUNCH, "mvs_conv3d_fwd: cuTensorMapEncodeTiled is not available from this driver");
    static thread_local Plan pl;
    MVS_REQUIRE(make_plan(d, pl), MVS_E_UNSUPPORTED, "mvs_conv3d_fwd: no tcgen05 tiling fits shared memory for Cin=%d Cout=%d", d->Cin, d->Cout);
    TcParams& p = pl.p;
    p.w = ws; p.scale = scale; p.shift = shift; p.skip = skip; p.y = y;
    cudaStream_t st = (cudaStream_t)stream;
    const int CoutPad = (d->Cout + 7) / 8 * 8;

    // ---- weight tiles, packed from the gather form into the caller's workspace (skipped when algo = 3 says they are there)
    const int nvec = p.nentries * p.kchunks * p.N;
    if (d->algo == 3) {
        // the caller kept the tiles of a frozen weight from an earlier algo = 2 call on the same workspace
    } else if (p.is_bf16) pack_tiles_kernel<__nv_bfloat16><<<mvs_cdiv(nvec, 256), 256, 0, st>>>(g, (__nv_bfloat16*)ws, pl.src, p.nentries, p.kchunks, p.N, d->Cin, d->Cout, CoutPad, p.CoP, p.nblk);
    else pack_tiles_kernel<__half><<<mvs_cdiv(nvec, 256), 256, 0, st>>>(g, (__half*)ws, pl.src, p.nentries, p.kchunks, p.N, d->Cin, d->Cout, CoutPad, p.CoP, p.nblk);

    // ---- tensor maps over x (zero fill outside the volume)
    TensorMaps maps;
    memset(&maps, 0, sizeof(maps));
    const CUtensorMapDataType dt = p.is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const cuuint64_t W = p.Wi, H = p.Hi, D = p.Di, NB = (cuuint64_t)p.B * p.CiB;
    if (p.mode != MODE_S2) {
        // dims (innermost first) {W*8, H, D, B*CiB}, box {256, PH, 1, 1}
        const cuuint64_t gdim[4] = {W * 8, H, D, NB};
        const cuuint64_t gstr[3] = {W * 16, H * W * 16, D * H * W * 16};
        const cuuint32_t box[4] = {(cuuint32_t)kPW * 8, (cuuint32_t)p.PH, 1, 1}, estr[4] = {1, 1, 1, 1};
        const CUresult cr = enc(&maps.m[0], dt, 4, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MVS_REQUIRE(cr == CUDA_SUCCESS, MVS_E_LAUNCH, "mvs_conv3d_fwd: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    } else {
        // one map per (h, w) parity: dims {8, W/2, H/2, D, B*CiB} with doubled h / w strides, box {8, 32, PH, 1, 1}
        for (int s = 0; s < 4; ++s) {
            const int ph = s >> 1, pw = s & 1;
            const cuuint64_t gdim[5] = {8, W / 2, H / 2, D, NB};
            const cuuint64_t gstr[4] = {32, 2 * W * 16, H * W * 16, D * H * W * 16};
            const cuuint32_t box[5] = {8, (cuuint32_t)kPW, (cuuint32_t)p.PH, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
            void* base = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(x)) + ((size_t)ph * W + pw) * 16;
            const CUresult cr = enc(&maps.m[s], dt, 5, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            MVS_REQUIRE(cr == CUDA_SUCCESS, MVS_E_LAUNCH, "mvs_conv3d_fwd: cuTensorMapEncodeTiled (parity %d) failed (%d)", s, (int)cr);
        }
    }
    p.nwork = (int)((int64_t)p.B * p.nwt * p.nht * p.nseg);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned nblocks = (unsigned)min(p.nwork, sms);          // persistent: one CTA per SM walks the work list
    if (p.is_bf16) {
        cudaFuncSetAttribute(conv3d_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        conv3d_tc_kernel<__nv_bfloat16><<<nblocks, kThreads, pl.smem, st>>>(maps, p);
    } else {
        cudaFuncSetAttribute(conv3d_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        conv3d_tc_kernel<__half><<<nblocks, kThreads, pl.smem, st>>>(maps, p);
    }
    return MVS_CHECK_LAUNCH("mvs_conv3d_fwd (tcgen05)");
}

#ifdef MVS_TC_TRACE
extern "C" int mvs_debug_tc_trace(long long* host_out) {   // [10][1024]
    return cudaMemcpyFromSymbol(host_out, g_trace, sizeof(long long) * 10 * 1024) == cudaSuccess ? 0 : -1;
}
#endif
