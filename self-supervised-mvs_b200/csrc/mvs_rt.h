// mvs_rt.h — small runtime layer shared by every kernel file of libmvs_b200.so.
//
// Two build modes:
//   * nvcc, -gencode arch=compute_100a,code=sm_100a  : the product.
//   * g++ -x c++ -DMVS_CPU_EMU                        : tests/emu only.  SIMT kernel bodies are compiled
//     for the host and a launch is a serial loop over (block, thread), so that index arithmetic, layouts
//     and gradient formulas can be checked against the oracle on a box without a GPU.  Kernels that need
//     block-level cooperation (shared memory, shuffles, mbarriers, tcgen05) are excluded from that build.
//     The emulation library is never loaded by the package unless a test binds it explicitly.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include "../../include/mvs_b200.h"

#ifdef MVS_CPU_EMU
// ------------------------------------------------------------------ host emulation shims
#include <cstring>
#include <algorithm>
using std::min;
using std::max;
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
typedef void* cudaStream_t;
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float fmaf_(float a, float b, float c) { return fmaf(a, b, c); }
template <typename F> static inline void mvs_emu_launch(dim3 g, dim3 b, F f) {
    gridDim = g; blockDim = b;
    for (unsigned bz = 0; bz < g.z; ++bz) for (unsigned by = 0; by < g.y; ++by) for (unsigned bx = 0; bx < g.x; ++bx)
        for (unsigned tz = 0; tz < b.z; ++tz) for (unsigned ty = 0; ty < b.y; ++ty) for (unsigned tx = 0; tx < b.x; ++tx) {
            blockIdx = dim3(bx, by, bz); threadIdx = dim3(tx, ty, tz); f();
        }
}
#define MVS_LAUNCH(kern, grid, block, stream, ...) mvs_emu_launch((grid), (block), [&]() { kern(__VA_ARGS__); })
#define MVS_CHECK_LAUNCH(name) (MVS_OK)
#else
// ------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#define MVS_LAUNCH(kern, grid, block, stream, ...) kern<<<(grid), (block), 0, (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define MVS_CHECK_LAUNCH(name) mvs_check_launch(name)
#endif

// ------------------------------------------------------------------ errors (thread-local message)
int mvs_set_error(int code, const char* fmt, ...);
#ifndef MVS_CPU_EMU
static inline int mvs_check_launch(const char* name) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return mvs_set_error(MVS_E_LAUNCH, "%s: %s", name, cudaGetErrorString(e));
    return MVS_OK;
}
#endif
#define MVS_REQUIRE(cond, code, ...) do { if (!(cond)) return mvs_set_error((code), __VA_ARGS__); } while (0)

// ------------------------------------------------------------------ test / tuning knobs (mvs_set_knob; never read from the environment)
enum { MVS_KNOB_WARP_TMA, MVS_KNOB_WARP_DC, MVS_KNOB_WARP_TMA_MINB, MVS_KNOB_WARP_CPT, MVS_KNOB_WARP_MINB, MVS_KNOB_WARP_DZ,
       MVS_KNOB_TC_KDFOLD, MVS_KNOB_TC_PLANES, MVS_KNOB_TC_NM, MVS_KNOB_TC_STAGES, MVS_KNOB_TC_NSEG, MVS_KNOB_WARP_BWD_SPLIT,
       MVS_KNOB_TC_KWFOLD_MAX, MVS_KNOB_COUNT };
extern int g_mvs_knobs[MVS_KNOB_COUNT];   // -1 = unset: the built-in (measured best) default applies
static inline int mvs_knob(int id, int dflt) { return g_mvs_knobs[id] < 0 ? dflt : g_mvs_knobs[id]; }

static inline unsigned mvs_cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// ------------------------------------------------------------------ 8-wide channel-block vectors
// load8/store8 move one (voxel, channel-block) of the C8 layout between memory and 8 fp32 registers.
template <typename T> struct V8;
template <> struct V8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
#ifdef MVS_CPU_EMU
        for (int i = 0; i < 8; ++i) v[i] = p[i];
#else
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#endif
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
#ifdef MVS_CPU_EMU
        for (int i = 0; i < 8; ++i) p[i] = v[i];
#else
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
#endif
    }
};
#ifndef MVS_CPU_EMU
template <> struct V8<__half> {
    static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
        uint4 r; __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    }
};
template <> struct V8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(u[i] << 16); v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 r; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    }
};
#endif

// dtype dispatch: calls `BODY` with `T` bound to the storage type of `code`.
#ifdef MVS_CPU_EMU
#define MVS_DISPATCH_DTYPE(code, T, BODY)                                                          \
    do { if ((code) == MVS_F32) { typedef float T; BODY; }                                        \
         else return mvs_set_error(MVS_E_UNSUPPORTED, "emulation build handles fp32 storage only"); } while (0)
#else
#define MVS_DISPATCH_DTYPE(code, T, BODY)                                                          \
    do { if ((code) == MVS_F32) { typedef float T; BODY; }                                        \
         else if ((code) == MVS_F16) { typedef __half T; BODY; }                                  \
         else if ((code) == MVS_BF16) { typedef __nv_bfloat16 T; BODY; }                          \
         else return mvs_set_error(MVS_E_ARG, "unknown dtype code %d", (int)(code)); } while (0)
#endif

static inline int mvs_dtype_size(int code) { return code == MVS_F32 ? 4 : 2; }
