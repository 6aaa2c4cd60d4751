/*
 * mvs_b200.h — C ABI of libmvs_b200.so: the B200 (sm_100a) plane-sweep hot path of MVSNet / CVP-MVSNet.
 *
 * The reference (ToughStoneX/Self-Supervised-MVS) has no FFI layer: its boundary for this path is the
 * Python surface train.py imports.  Each entry point below names the reference
 * function (file:line under /root/reference) whose arithmetic it replaces; the Python mirror of that
 * surface lives in self-supervised-mvs_b200/ and binds these symbols with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter is documented "host"; the caller owns all
 *     memory (inputs, outputs, workspaces); kernels never allocate, free or retain pointers.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no internal sync.
 *   - return value: 0 = ok, <0 = error (MVS_E_*); mvs_last_error() returns a thread-local message.
 *   - dtype codes select the STORAGE type of activations; all accumulation is fp32.
 *
 * Activation layout "C8": channels are blocked by 8 and the block is innermost:
 *       maps    [B][C/8][H][W][8]          volumes [B][C/8][D][H][W][8]
 *   so one (voxel, channel-block) is a 16-byte (fp16/bf16) or 32-byte (fp32) vector, consecutive w are
 *   contiguous (coalesced gathers / stores), and a TMA box over (8, W, H, D, C/8) lands in shared memory
 *   as the K-major, row-contiguous operand the tcgen05 implicit-GEMM convolution consumes.
 *   Single-channel volumes (cost_reg, prob) are plain fp32 [B][D][H][W].
 */
#ifndef MVS_B200_H
#define MVS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVS_B200_VERSION 102 /* major*100 + minor */

enum { MVS_F32 = 0, MVS_F16 = 1, MVS_BF16 = 2 };
enum { MVS_OK = 0, MVS_E_ARG = -1, MVS_E_SHAPE = -2, MVS_E_LAUNCH = -3, MVS_E_UNSUPPORTED = -4 };
#define MVS_MAX_SRC 8

int mvs_version(void);
const char* mvs_last_error(void);
/* 1 when this library is the host-emulation build used by the CPU unit tests (tests/emu), else 0. */
int mvs_is_emulation(void);
/* Test / tuning knobs (kernel-variant selection for A/B parity tests and the tools/ sweeps).  The product path never needs this
 * call and the library never reads the environment.  value < 0 restores the built-in default.  Names: warp_tma (0 = gather from
 * global memory instead of TMA-staged windows), warp_dc, warp_tma_minb, warp_cpt, warp_minb, warp_dz, tc_kdfold (0 = off),
 * tc_planes (1 = one image plane per step), tc_nm, tc_stages, tc_nseg, warp_bwd_split (0 = sweep backward with one thread per
 * (pixel, channel block) instead of one per (pixel, channel block, source)). */
int mvs_set_knob(const char* name, int value);

/* ---- layout ------------------------------------------------------------------------------------------- */
/* fp32 [B][C][S] (S = H*W or D*H*W) -> C8 [B][C/8][S][8] in `dtype`.  C % 8 == 0. */
int mvs_pack_c8(const float* src, void* dst, int B, int C, int64_t S, int dtype, void* stream);
/* channels-last [B][S][C] in `dtype` (the layout the library 2-D feature extractor emits) -> C8 [B][C/8][S][8], same dtype */
int mvs_nhwc_to_c8(const void* src, void* dst, int B, int C, int64_t S, int dtype, void* stream);
/* Zero-bordered maps "C8P" for the plane-sweep gather: dst [M][C/8][H+3][W+2][8] in `dtype`, pixel (y,x) at row y+1, column x+1,
 * zeros elsewhere (so that grid_sample's zero padding becomes a plain load).  A C8P buffer handed to mvs_warp_var_fwd / _bwd must be
 * followed by 16 readable bytes holding zeros (one vector): a sample beyond the bottom-right corner reads its zero-weight fourth
 * tap at row H+2, column W+2, which for the last plane lies one vector past the array.  src_layout: 0 = fp32 [M][C][H][W],
 * 1 = channels-last `dtype` [M][H][W][C], 2 = C8 `dtype` [M][C/8][H][W][8]. */
int mvs_pack_c8_padded(const void* src, void* dst, int M, int C, int H, int W, int src_layout, int dtype, void* stream);
/* inverse of mvs_pack_c8 */
int mvs_unpack_c8(const void* src, float* dst, int B, int C, int64_t S, int dtype, void* stream);

/* ---- projection algebra -------------------------------------------------------------------------------- */
/* rt[s][b][0..8] = rot (row major), rt[s][b][9..11] = trans of  proj[b][s+1] @ inverse(proj[b][0]),
 * proj = [B][N][4][4] fp32, rt = [N-1][B][12] fp32.  Replaces jdacs/models/module.py:116-118 (fp64 inside). */
int mvs_compose_proj(const float* proj, float* rt, int B, int N, void* stream);
/* same from intrinsics/extrinsics: P = [[K E[:3]],[0,0,0,1]].  ref_in [B][3][3], src_in [B][nsrc][3][3],
 * ref_ex [B][4][4], src_ex [B][nsrc][4][4]; intrinsics rows 0,1 are divided by `down` first
 * (conditionIntrinsics).  Replaces jdacs-ms/models/modules.py:22-37, 71-80, 222-231. */
int mvs_compose_proj_ke(const float* ref_in, const float* src_in, const float* ref_ex, const float* src_ex,
                        float down, float* rt, int B, int nsrc, void* stream);

/* ---- a1/a2: stand-alone homography warp (API parity with homo_warping) ------------------------------------ */
/* src [B][C][H][W] fp32 -> out [B][C][D][H][W] fp32.  depth is [B][D] (per_pixel=0) or [B][D][H][W] (1).
 * rt = [B][12].  jdacs/models/module.py:105-140, jdacs-ms/models/modules.py:62-104. */
int mvs_homo_warp_fwd(const float* src, const float* rt, const float* depth, int per_pixel, float* out,
                      int B, int C, int D, int H, int W, int align_corners, void* stream);
/* grad_src [B][C][H][W] must be zero-initialised by the caller (4-tap scatter add). */
int mvs_homo_warp_bwd(const float* grad_out, const float* rt, const float* depth, int per_pixel, float* grad_src,
                      int B, int C, int D, int H, int W, int align_corners, void* stream);

/* ---- a1+a3+a4: fused warp + bilinear gather + running variance ------------------------------------------------
 * ref, srcs[i]: C8 maps [B][C/8][H][W][8] (pad=0) or zero-bordered C8P maps [B][C/8][H+3][W+2][8] (pad=1, the fast path
 * for 16-bit storage) in dtype_in; var: C8 volume [B][C/8][D][H][W][8] (dtype_out).
 * srcs = HOST array of nsrc device pointers; rt = [nsrc][B][12].
 * var = S2/N - (S1/N)^2, N = nsrc+1, S1 = ref + sum warped, S2 = ref^2 + sum warped^2;
 * ref_sq_in_sum=1 starts S1 from ref^2 (CVP aliasing, jdacs-ms/models/network.py:114-116).
 * Replaces jdacs/models/mvsnet.py:120-136, jdacs-ms/models/network.py:114-137, modules.py:209-261. */
int mvs_warp_var_fwd(const void* ref, const void* const* srcs, int nsrc, const float* rt, const float* depth,
                     int per_pixel, void* var, int B, int C, int D, int H, int W, int dtype_in, int dtype_out,
                     int align_corners, int ref_sq_in_sum, int pad, void* stream);
/* grad_var: C8 volume (dtype_out).  grad_ref and grad_srcs[i]: fp32 C8 maps (never padded), zero-initialised by the caller. */
int mvs_warp_var_bwd(const void* grad_var, const void* ref, const void* const* srcs, int nsrc, const float* rt,
                     const float* depth, int per_pixel, float* grad_ref, float* const* grad_srcs, int B, int C,
                     int D, int H, int W, int dtype_in, int dtype_out, int align_corners, int ref_sq_in_sum,
                     int pad, void* stream);

/* ---- a5/a6: 3-D convolution stack ------------------------------------------------------------------------------ */
typedef struct {
    int B, Cin, Cout;
    int Din, Hin, Win;     /* input volume  */
    int Dout, Hout, Wout;  /* output volume */
    int stride;            /* 1 or 2 */
    int transposed;        /* 0: Conv3d(k=3,pad=1,stride)   1: ConvTranspose3d(k=3,pad=1,stride,output_padding=stride-1) */
    int dtype_in, dtype_out;
    int relu;              /* apply ReLU after the affine */
    int algo;              /* 0 auto, 1 SIMT fp32 direct, 2 tcgen05 implicit GEMM, 3 tcgen05 with `ws` already packed by an earlier algo-2 call */
} mvs_conv3d_desc;

/* torch weight ([Cout][Cin][27] for Conv3d, [Cin][Cout][27] for ConvTranspose3d) -> gather form
 * G[27][Cin][CoutPad] fp32, CoutPad = round_up(Cout, 8), so that out[o] = sum_tap x[in(o,tap)] . G[tap]. */
int mvs_pack_conv3d_weight(const float* w, float* g, int Cin, int Cout, int transposed, void* stream);
/* y = [relu]( conv(x) * scale + shift ) + skip;  scale/shift [Cout] fp32 or NULL, skip (dtype_out, y's layout) or NULL.
 * x: C8 volume.  y: C8 volume, or plain fp32 [B][D][H][W] when Cout == 1.
 * Replaces ConvBnReLU3D / ConvTranspose3d+BN+ReLU / prob: jdacs/models/module.py:35-42, mvsnet.py:37-74,
 * jdacs-ms/models/network.py:44-74. */
int mvs_conv3d_fwd(const mvs_conv3d_desc* d, const void* x, const float* g, const float* scale, const float* shift,
                   const void* skip, void* y, void* ws, void* stream);
/* bytes of caller-owned scratch `ws` mvs_conv3d_fwd needs for this descriptor (0 for the SIMT path).  The tcgen05
 * path packs the weight tap tiles into it (algo 0 / 2); the tiles depend only on g, Cin, Cout, stride, transposed and
 * the storage dtype, so a caller with frozen weights may keep `ws` and pass algo = 3 to skip the re-pack. */
int64_t mvs_conv3d_workspace_bytes(const mvs_conv3d_desc* d);
/* training: gradient w.r.t. the torch-layout weight ([Cout][Cin][27] or, transposed, [Cin][Cout][27]; zero-initialised
 * by the caller).  x: C8 volume (dtype_in), grad_y: C8 fp32 volume of the un-activated convolution output.
 * (The gradient w.r.t. x needs no entry point of its own: it is mvs_conv3d_fwd of grad_y with the same torch weight
 * packed under the opposite `transposed` flag, exactly how ATen defines conv_transpose3d.) */
int mvs_conv3d_bwd_weight(const mvs_conv3d_desc* d, const void* x, const float* grad_y, float* grad_w, void* stream);

/* ---- f1 (next row): the 2-D feature extractors on the same tcgen05 kernel ---------------------------------------------------
 * A stack of M images is convolved as a volume whose depth axis is the image index (no taps across images):
 *   x: C8 stack [Cin/8][M][Hin][Win][8];  y: C8 stack [Cout/8][M][Hout][Wout][8], or (out_padded) the zero-bordered
 *   image-major maps [M][Cout/8][Hout+3][Wout+2][8] that mvs_warp_var_fwd(pad=1) gathers from (border pre-zeroed by the caller).
 *   y = [relu](conv(x) * scale + shift);  g = gather-form weights G[ksize*ksize][Cin][CoutPad] fp32 (tap = kh * ksize + kw).
 * ksize/stride: 3/1 (pad 1) or 5/2 (pad 2, even Hin/Win).  Replaces ConvBnReLU / Conv2d of FeatureNet
 * (jdacs/models/mvsnet.py:17-34, module.py:13-32) and conv + LeakyReLU of FeaturePyramid (jdacs-ms/models/network.py:16-41,
 * modules.py:15-19) in eval mode with 16-bit storage. */
typedef struct {
    int M, Cin, Cout;      /* Cin in {8,16,32,48,64} (3-channel images are zero-padded to 8), Cout in {8,..,64} */
    int Hin, Win, Hout, Wout;
    int ksize, stride;
    int dtype;             /* MVS_F16 | MVS_BF16 */
    int relu;              /* activation: 0 none, 1 ReLU, 2 LeakyReLU(leaky_slope) */
    int out_padded;
    int ws_packed;         /* 1: `ws` already holds the weight tiles of an earlier call with the same g (frozen weights) */
    float leaky_slope;
} mvs_conv2d_desc;
int64_t mvs_conv2d_workspace_bytes(const mvs_conv2d_desc* d);
int mvs_conv2d_fwd(const mvs_conv2d_desc* d, const void* x, const float* g, const float* scale, const float* shift, void* y,
                   void* ws, void* stream);
/* images [B][N][3][H][W] stored as `src_dtype` (fp32, or fp16 / bf16 as a host pipeline may upload them: the first layer rounds
 * its input to the storage type anyway, so 16-bit images of the volume dtype give bit-identical results at half the H2D bytes)
 * -> C8 stack [1][M = N*B][H][W][8] in `dtype`, image index m = v * B + b, channels 3..7 zero. */
int mvs_pack_images_c8(const void* imgs, int src_dtype, void* dst, int B, int N, int H, int W, int dtype, void* stream);

/* batch-norm helpers for training mode (statistics over B*D*H*W per channel), C8 fp32 volumes.
 * sums = [2][C] (sum, sum of squares), zero-initialised by the caller.  jdacs/models/module.py:39-42. */
int mvs_bn_stats(const float* x, float* sums, int B, int C, int64_t S, void* stream);
/* y = [relu]((x - mean) * invstd * gamma + beta) + skip        (skip may be NULL) */
int mvs_bn_act_fwd(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                   const float* skip, float* y, int B, int C, int64_t S, int relu, void* stream);
/* backward pass 1: red[0][c] = sum g, red[1][c] = sum g*xhat with g = grad_y * relu'(.) (= grad_beta, grad_gamma);
 * red zero-initialised by the caller. */
int mvs_bn_act_bwd_reduce(const float* x, const float* grad_y, const float* mean, const float* invstd,
                          const float* gamma, const float* beta, float* red, int B, int C, int64_t S, int relu,
                          void* stream);
/* backward pass 2: grad_x = gamma*invstd*(g - red[0]/M - xhat*red[1]/M), M = B*S */
int mvs_bn_act_bwd_apply(const float* x, const float* grad_y, const float* mean, const float* invstd,
                         const float* gamma, const float* beta, const float* red, float* grad_x, int B, int C,
                         int64_t S, int relu, void* stream);

/* ---- training with 16-bit activations (fp32 master weights, fp32 / fp64 accumulation): csrc/train.cu ------------------------------
 * The forward convolution and the gradient w.r.t. the input are mvs_conv3d_fwd on the tcgen05 kernel (the latter with the same torch
 * weight packed under the opposite `transposed` flag); the passes below are what BatchNorm3d + ReLU (+ skip) and the weight gradient
 * add.  z / y / grad_* volumes are C8 in `dtype` (any of MVS_F32 / F16 / BF16).  Reference: the autograd of ConvBnReLU3D and
 * ConvTranspose3d + BatchNorm3d + ReLU, jdacs/models/module.py:35-42, mvsnet.py:37-74, jdacs-ms/models/network.py:44-74. */
/* sums = [2][C] doubles (sum, sum of squares over B*S), zero-initialised by the caller */
int mvs_bn_stats_t(const void* z, int dtype, double* sums, int B, int C, int64_t S, void* stream);
/* batch statistics -> a = gamma * invstd, b = beta - mean * a, mean, invstd ([C] fp32 each); running_mean / running_var (may be
 * NULL) are updated like nn.BatchNorm3d.train(): momentum, unbiased variance.  count = B*S. */
int mvs_bn_finalize(const double* sums, const float* gamma, const float* beta, float eps, float momentum, double count, float* a,
                    float* b, float* mean, float* invstd, float* running_mean, float* running_var, int C, void* stream);
/* y = [relu](z * a + b) + skip      (skip may be NULL) */
int mvs_bn_act_fwd_t(const void* z, const float* a, const float* b, const void* skip, void* y, int dtype, int B, int C, int64_t S,
                     int relu, void* stream);
/* red = [2][C] doubles, zero-initialised: red[0] = sum g (= grad beta), red[1] = sum g * xhat (= grad gamma), g = grad_y [z a + b > 0] */
int mvs_bn_act_bwd_reduce_t(const void* z, const void* grad_y, const float* a, const float* b, const float* mean, const float* invstd,
                            double* red, int dtype, int B, int C, int64_t S, int relu, void* stream);
/* grad_z = a (g - red[0]/M - xhat red[1]/M), M = B*S; frozen = 1 (statistics not taken from the batch): grad_z = a g.
 * grad_gamma / grad_beta ([C] fp32, may be NULL) receive red[1] / red[0]. */
int mvs_bn_act_bwd_apply_t(const void* z, const void* grad_y, const float* a, const float* b, const float* mean, const float* invstd,
                           const double* red, void* grad_z, float* grad_gamma, float* grad_beta, int dtype, int B, int C, int64_t S,
                           int relu, int frozen, void* stream);
/* plain fp32 [n] -> C8 block [n][8] in `dtype`, value in channel 0 (lifts the gradient of the single-channel `prob` output) */
int mvs_lift_c1(const float* src, void* dst, int dtype, int64_t n, void* stream);
/* Weight gradient on tensor cores (warp-level mma.sync m16n8k16, fp32 accumulation; why not tcgen05: see csrc/train.cu).
 * x and grad_z: C8 volumes in d->dtype_in (fp16 / bf16), Cin and Cout multiples of 8 (<= 64); grad_w: torch layout [Cout][Cin][27]
 * or, transposed, [Cin][Cout][27], fp32, zero-initialised by the caller, with cout_real <= d->Cout output channels (1 for a lifted
 * single-channel gradient). */
int mvs_conv3d_wgrad_mma(const mvs_conv3d_desc* d, const void* x, const void* grad_z, float* grad_w, int cout_real, void* stream);
/* The same for the 2-D layers of the feature extractor run as zero-kd 3-D layers over an image volume (depth axis = image index,
 * jdacs/models/mvsnet.py:17-34 in training): only the kd = 1 taps are computed; grad_w keeps the [.][.][3][3][3] shape, its
 * kd = 0, 2 planes are left untouched. */
int mvs_conv2d_wgrad_mma(const mvs_conv3d_desc* d, const void* x, const void* grad_z, float* grad_w, int cout_real, void* stream);

/* ---- a7/a8: softmax + soft-argmin + photometric confidence ------------------------------------------------------ */
/* cost [B][D][H][W] fp32; depth [B][D] or [B][D][H][W]; outputs (any may be NULL): depth_out [B][H][W] fp32,
 * index_out [B][H][W] int64 = trunc(sum_d p_d*d), conf_out [B][H][W] = p[i-1]+p[i]+p[i+1]+p[i+2], prob_out [B][D][H][W].
 * jdacs/models/mvsnet.py:141-151, module.py:145-148, jdacs-ms/models/modules.py:324-331, network.py:183-189. */
int mvs_softargmin_fwd(const float* cost, const float* depth, int per_pixel, float* depth_out, int64_t* index_out,
                       float* conf_out, float* prob_out, int B, int D, int H, int W, void* stream);
/* grad_cost[d] = p_d (depth_d - E[depth]) grad_depth */
int mvs_softargmin_bwd(const float* cost, const float* depth, int per_pixel, const float* grad_depth,
                       float* grad_cost, int B, int D, int H, int W, void* stream);

/* ---- a9: CVP per-pixel depth hypotheses ---------------------------------------------------------------------- */
/* hypos[b][k][h][w] = depth_up[b][h][w] + (k-half)*interval_b, interval_b = mean_pixels |delta_d| (fp64).
 * ref_in, src_in0 [B][3][3]; ref_ex, src_ex0 [B][4][4] (source view 0 only); ws = [B] doubles, zero-initialised.
 * jdacs-ms/models/modules.py:107-206. */
int mvs_depth_hypo_refine(const float* depth_up, const float* ref_in, const float* src_in0, const float* ref_ex,
                          const float* src_ex0, float* hypos, double* ws, int B, int H, int W, int half, void* stream);

/* ---- a10: photometric inverse warp (loss side) ------------------------------------------------------------------ */
/* img [B][H][W][C] fp32; left_cam/right_cam [B][2][4][4] ([0]=E, [1][:3][:3]=K); depth [B][H][W];
 * warped [B][H][W][C], mask [B][H][W].  cam_ws = [B][24] floats of scratch.
 * jdacs/losses/homography.py:186-238, 292-374. */
int mvs_invwarp_fwd(const float* img, const float* left_cam, const float* right_cam, const float* depth,
                    float* warped, float* mask, float* cam_ws, int B, int H, int W, int C, void* stream);
/* grad_depth [B][H][W] (written); grad_img [B][H][W][C] or NULL (zero-initialised, scatter add). */
int mvs_invwarp_bwd(const float* img, const float* left_cam, const float* right_cam, const float* depth,
                    const float* grad_warped, float* grad_depth, float* grad_img, float* cam_ws, int B, int H,
                    int W, int C, void* stream);

/* ---- a11 / f2: the self-supervised photometric loss, fused ----------------------------------------------------- */
/* UnSupLoss.forward (jdacs/losses/unsup_loss.py:24-83; jdacs-ms/losses/unsup_loss.py:23-86) on losses/modules.py:17-90 and
 * the warp above: x0.25 resize of the views, inverse warp of every source view, masked smooth-L1 of colours and colour
 * gradients per view, 3x3 SSIM (views 1 and 2), edge-aware depth smoothness, top-3 view selection per pixel.
 *   imgs  [B][N][3][Hi][Wi] fp32 with (Hi, Wi) = (H, W) or 4x that (bilinear x0.25, as the jdacs tree does); cams [B][N][2][4][4];
 *   depth [B][H][W]; N - 1 >= 3 source views (the reference's top-k with k = 3 fails below that: MVS_E_SHAPE).
 *   out   [4] = { 12 rec + 6 ssim + smooth_weight smooth, rec (reconstr_loss), ssim (ssim_loss), smooth (smooth_loss) }.
 * Buffers the backward reuses (caller-owned, written here): small [N][B][H][W][3], warped [N-1][B][H][W][3],
 * mask [N-1][B][H][W], coef [2][B][H][W][9], cam_ws [(N-1) B][24], acc [MVS_LOSS_ACC_DOUBLES] doubles. */
#define MVS_LOSS_ACC_DOUBLES 48
int mvs_unsup_loss_fwd(const float* imgs, const float* cams, const float* depth, int B, int N, int Hi, int Wi, int H, int W,
                       float smooth_lambda, float smooth_weight, float* small, float* warped, float* mask, float* coef,
                       float* cam_ws, double* acc, float* out, void* stream);
/* grad_out [4] = d L / d out (device); grad_depth [B][H][W] written.  The images carry no gradient. */
int mvs_unsup_loss_bwd(const float* grad_out, const float* small, const float* depth, const float* warped, const float* mask,
                       const float* coef, const float* cam_ws, const double* acc, float* grad_depth, int B, int N, int H, int W,
                       float smooth_lambda, float smooth_weight, void* stream);

/* ---- f3: inference output side --------------------------------------------------------------------------------- */
/* F.interpolate(map.unsqueeze(1), size=(Ho, Wo)) in the default 'nearest' mode for M maps [M][H][W] -> [M][Ho][Wo]
 * (jdacs/eval_dense.py:150-153).  flip_rows != 0 stores the rows bottom-up: the order of a .pfm body (data_io.py:53-80). */
int mvs_upsample_nearest(const float* src, float* dst, int M, int H, int W, int Ho, int Wo, int flip_rows, void* stream);
/* write_depth_img (jdacs/eval_dense.py:110-121): out = clamp((depth - offset) / scale, 0, 255) truncated to 8 bits. */
int mvs_depth_preview_u8(const float* depth, uint8_t* out, int64_t n, float offset, float scale, void* stream);
/* reproject_with_depth + check_geometric_consistency (jdacs/eval_dense.py:177-232) for B (reference, source) pairs of depth maps
 * [B][H][W].  cams [B][60] doubles = inv(K_ref) 3x3 | (E_src inv(E_ref))[:3] 3x4 | K_src | inv(K_src) | (E_ref inv(E_src))[:3] | K_ref,
 * row-major, computed by the caller in float32 as the reference does.  Outputs (any may be NULL): mask [B][H][W] bytes
 * (dist < dist_thresh and |d_reproj - d| / d < rel_thresh), depth_reprojected (zeroed outside the mask when apply_mask),
 * the float32 source coordinates, the float32 re-projected coordinates.  cv2.remap's 1/32-pixel bilinear sampling is
 * reproduced exactly. */
int mvs_geo_consistency(const float* depth_ref, const float* depth_src, const double* cams, uint8_t* mask,
                        float* depth_reprojected, float* x_src, float* y_src, float* x_reprojected, float* y_reprojected,
                        int B, int H, int W, float dist_thresh, float rel_thresh, int apply_mask, void* stream);

/* ---- f1: fused full-resolution front of FeatureNet (jdacs/models/mvsnet.py:20-23, 39-40) ------------------------------ */
/* conv0 3x3 (3->8) + BN + ReLU -> conv1 3x3 (8->8) + BN + ReLU -> conv2 5x5 stride 2 (8->16) + BN + ReLU in one kernel (eval mode,
 * BatchNorm folded to per-channel affines): the 8-channel full-resolution maps stay in shared memory.
 *   imgs [B][N][3][H][W] in img_dtype (fp32 or the volume dtype); out = C8 stack [2][M = N*B][H/2][W/2][8] in `dtype` (fp16 / bf16),
 *   image m = v * B + b (the input layout of mvs_conv2d_fwd);  affine [3][32] = per layer scale[16] | shift[16] (fp32);
 *   wfrag = mvs_featnet_front_workspace_bytes() bytes written by mvs_featnet_front_pack from the torch weights
 *   w0 [8][3][3][3], w1 [8][8][3][3], w2 [16][8][5][5] (fp32). */
int64_t mvs_featnet_front_workspace_bytes(void);
int mvs_featnet_front_pack(const float* w0, const float* w1, const float* w2, void* wfrag, int dtype, void* stream);
int mvs_featnet_front(const void* imgs, int img_dtype, const void* wfrag, const float* affine, void* out, int B, int N,
                      int H, int W, int dtype, void* stream);

/* ---- f4: depth-map fusion (vendored Gipuma fusibile, jdacs/fusion/fusibile/fusibile.cu:138-277) ----------------------- */
/* normals_depths [V][H][W][4] = (nx, ny, nz, depth) per view; images [V][H][W][4] colours or NULL; cams [V][32] floats =
 * P 3x4 | inv(P[:, :3]) 3x3 | P[:, 3] | camera centre C | focal length f | pad; subset [nsub] view indices to test against view
 * `ref`.  points [H][W][12] = fused coordinate (xyz0) | normal (xyz0) | colour (rgb0), valid [H][W] = 1 where at least
 * num_consistent views agreed (disparity difference < depth_thresh, normal angle < normal_thresh radians). */
int mvs_fusibile(const float* normals_depths, const float* images, const float* cams, const int* subset, int nsub, int V,
                 int H, int W, int ref, float depth_thresh, float normal_thresh, int num_consistent, float* points,
                 uint8_t* valid, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVS_B200_H */
