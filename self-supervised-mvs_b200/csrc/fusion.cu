// fusion.cu — depth-map fusion: the per-pixel consensus kernel of the vendored Gipuma `fusibile`
// (jdacs/fusion/fusibile/fusibile.cu:138-277, helpers :46-133; the same file in jdacs-ms/fusion/fusibile) rebuilt for sm_100a
// as a plain gather kernel over ordinary device arrays: no managed memory, no texture objects, no GlobalState indirection, and
// no camera block passed in the dynamic-shared-memory launch slot (the `<<<grid, block, cam>>>` of fusibile.cu:421).
//
// For every pixel of the reference view: lift it to 3-D with its depth, project the point into every other selected view,
// read that view's (normal, depth) there with bilinear filtering, and accept the view when the two depths agree in disparity
// space (|f b / d - f b / d'| < depth_thresh, b = baseline of the two cameras) and the normals differ by less than
// normal_thresh radians.  Accepted views contribute their own 3-D point (taken at the truncated pixel) and normal (and colour)
// to an average; the fused point is kept when at least num_consistent views agreed.
//
// The original samples float4 textures with cudaFilterModeLinear at (x + 0.5, y + 0.5), un-normalised coordinates (main.cpp:
// 489-493): texel-centre bilinear filtering with clamped addresses and 8-bit fractional weights (CUDA programming guide, "Linear
// Filtering": alpha is stored in 9-bit fixed point with 8 fractional bits).  Reproduced here with ordinary loads: weights rounded
// to 1/256.  Parity: the fusibile binary cannot be built in this container (OpenCV C++ / cmake), so this kernel is checked
// against the restatement in oracle/fusion.py only -- parity unpinned, said so in DESIGN.md.
#include "mvs_rt.h"

namespace {

// camera block, 32 floats: P 3x4 [0,12) | M_inv 3x3 [12,21) | P_col34 [21,24) | C [24,27) | f [27] | pad
struct F4 { float x, y, z, w; };

__device__ __forceinline__ void point_of(const float* __restrict__ cam, float px, float py, float depth, float (&X)[3]) {
    // get3Dpoint_cu (fusibile.cu:56-65): M_inv (depth (x, y, 1) - P[:, 3])
    const float a = depth * px - cam[21], b = depth * py - cam[22], c = depth - cam[23];
    for (int r = 0; r < 3; ++r) X[r] = cam[12 + r * 3] * a + cam[12 + r * 3 + 1] * b + cam[12 + r * 3 + 2] * c;
}

__device__ __forceinline__ F4 tex_linear(const float* __restrict__ img, int H, int W, float x, float y) {
    // tex2D<float4>(tex, x + 0.5, y + 0.5), linear filter, clamp: sample position (x, y) in texel-centre space
    const float fx = floorf(x), fy = floorf(y);
    const float ax = floorf((x - fx) * 256.f + 0.5f) * (1.f / 256.f), ay = floorf((y - fy) * 256.f + 0.5f) * (1.f / 256.f);
    const int x0 = max(0, min((int)fx, W - 1)), x1 = max(0, min((int)fx + 1, W - 1));
    const int y0 = max(0, min((int)fy, H - 1)), y1 = max(0, min((int)fy + 1, H - 1));
    const float* t00 = img + ((int64_t)y0 * W + x0) * 4; const float* t10 = img + ((int64_t)y0 * W + x1) * 4;
    const float* t01 = img + ((int64_t)y1 * W + x0) * 4; const float* t11 = img + ((int64_t)y1 * W + x1) * 4;
    float o[4];
    for (int k = 0; k < 4; ++k)
        o[k] = (1.f - ax) * (1.f - ay) * __ldg(t00 + k) + ax * (1.f - ay) * __ldg(t10 + k) + (1.f - ax) * ay * __ldg(t01 + k) + ax * ay * __ldg(t11 + k);
    return F4{o[0], o[1], o[2], o[3]};
}

__global__ void __launch_bounds__(128)
fusibile_kernel(const float* __restrict__ nd, const float* __restrict__ images, const float* __restrict__ cams,
                const int* __restrict__ subset, int nsub, int H, int W, int ref, float depth_thresh, float normal_thresh,
                int num_consistent, float* __restrict__ points, uint8_t* __restrict__ valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const int py = i / W, px = i - py * W;
    const int64_t HW4 = (int64_t)H * W * 4;
    const float* cref = cams + (int64_t)ref * 32;
    const float* n0 = nd + (int64_t)ref * HW4 + (int64_t)i * 4;
    const float nx = __ldg(n0), ny = __ldg(n0 + 1), nz = __ldg(n0 + 2), depth = __ldg(n0 + 3);
    float X[3];
    point_of(cref, (float)px, (float)py, depth, X);
    float cX[3] = {X[0], X[1], X[2]}, cN[4] = {nx, ny, nz, depth}, cT[4] = {0.f, 0.f, 0.f, 0.f};
    if (images) for (int k = 0; k < 4; ++k) cT[k] = __ldg(images + (int64_t)ref * HW4 + (int64_t)i * 4 + k);
    int n = 0;
    for (int s = 0; s < nsub; ++s) {
        const int v = __ldg(subset + s);
        if (v == ref) continue;
        const float* cv = cams + (int64_t)v * 32;
        // project_on_camera (:127-133)
        const float tx = cv[0] * X[0] + cv[1] * X[1] + cv[2] * X[2] + cv[3];
        const float ty = cv[4] * X[0] + cv[5] * X[1] + cv[6] * X[2] + cv[7];
        const float tz = cv[8] * X[0] + cv[9] * X[1] + cv[10] * X[2] + cv[11];
        const float qx = tx / tz, qy = ty / tz;
        if (!(qx >= 0.f && qx < (float)W && qy >= 0.f && qy < (float)H)) continue;
        const F4 t = tex_linear(nd + (int64_t)v * HW4, H, W, qx, qy);
        // disparityDepthConversion_cu2 (:46-49): focal length of the REFERENCE camera, baseline between the two centres
        const float bx = cref[24] - cv[24], by = cref[25] - cv[25], bz = cref[26] - cv[26];
        const float fb = cref[27] * sqrtf(bx * bx + by * by + bz * bz);
        if (!(fabsf(fb / tz - fb / t.w) < depth_thresh)) continue;
        float angle = acosf(t.x * nx + t.y * ny + t.z * nz);            // getAngle_cu (:118-126)
        if (angle != angle) angle = 0.f;
        if (!(angle < normal_thresh)) continue;
        float Y[3];
        point_of(cv, (float)(int)qx, (float)(int)qy, t.w, Y);
        for (int k = 0; k < 3; ++k) cX[k] += Y[k];
        cN[0] += t.x; cN[1] += t.y; cN[2] += t.z; cN[3] = 0.f;          // the original's float4 operator+ drops w
        if (images) { const F4 c = tex_linear(images + (int64_t)v * HW4, H, W, qx, qy); cT[0] += c.x; cT[1] += c.y; cT[2] += c.z; cT[3] = 0.f; }
        ++n;
    }
    const float inv = 1.f / ((float)n + 1.f);
    const bool keep = n >= num_consistent;
    valid[i] = keep ? 1 : 0;
    float* o = points + (int64_t)i * 12;
    for (int k = 0; k < 3; ++k) { o[k] = keep ? cX[k] * inv : 0.f; o[4 + k] = keep ? cN[k] * inv : 0.f; o[8 + k] = keep ? cT[k] * inv : 0.f; }
    o[3] = 0.f; o[7] = 0.f; o[11] = 0.f;
}

}  // namespace

extern "C" int mvs_fusibile(const float* normals_depths, const float* images, const float* cams, const int* subset, int nsub, int V,
                            int H, int W, int ref, float depth_thresh, float normal_thresh, int num_consistent, float* points,
                            uint8_t* valid, void* stream) {
    MVS_REQUIRE(normals_depths && cams && subset && points && valid, MVS_E_ARG, "mvs_fusibile: null pointer");
    MVS_REQUIRE(V > 0 && H > 0 && W > 0 && nsub > 0 && ref >= 0 && ref < V && (int64_t)H * W < (1ll << 29), MVS_E_SHAPE, "mvs_fusibile: bad dims");
    MVS_LAUNCH(fusibile_kernel, dim3(mvs_cdiv((int64_t)H * W, 128)), dim3(128), stream, normals_depths, images, cams, subset, nsub, H, W, ref,
               depth_thresh, normal_thresh, num_consistent, points, valid);
    return MVS_CHECK_LAUNCH("mvs_fusibile");
}
