// conv3d_tc.cu — tcgen05 implicit-GEMM 3x3x3 convolution (algo = 2).  Placeholder until the kernel lands.
#include "mvs_rt.h"

int mvs_conv3d_tc_supported(const mvs_conv3d_desc* d) { (void)d; return 0; }

int mvs_conv3d_fwd_tc(const mvs_conv3d_desc* d, const void* x, const float* g, const float* scale, const float* shift,
                      const void* skip, void* y, void* stream) {
    (void)d; (void)x; (void)g; (void)scale; (void)shift; (void)skip; (void)y; (void)stream;
    return mvs_set_error(MVS_E_UNSUPPORTED, "mvs_conv3d_fwd: tcgen05 path not built yet");
}
