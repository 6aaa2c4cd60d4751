/* A plain-C99 client of include/mvs_b200.h -- what a cgo / JNI / N-API binding of the reference's host language would compile
 * against.  TEST INFRASTRUCTURE: tests/test_cabi.py builds it with gcc and links it to the host-emulation build of the kernel
 * sources (host pointers); against libmvs_b200.so the same calls take device pointers.  Exercises: version / build query, the
 * layout round trip, soft-argmin + index + confidence on a column with a known answer, the error convention (negative code,
 * thread-local message, nothing written). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mvs_b200.h"

#define CHECK(cond)                                                                                    \
    do {                                                                                               \
        if (!(cond)) {                                                                                 \
            fprintf(stderr, "cabi_client: %s failed at line %d (%s)\n", #cond, __LINE__, mvs_last_error()); \
            return 1;                                                                                  \
        }                                                                                              \
    } while (0)

int main(void) {
    enum { B = 2, C = 16, S = 15, D = 8, H = 1, W = 3 };
    float src[B * C * S], back[B * C * S], packed[B * C * S];
    float cost[B * D * H * W], planes[B * D], depth[B * H * W], conf[B * H * W];
    int64_t index[B * H * W];
    int i, b, d, p;

    CHECK(mvs_version() == MVS_B200_VERSION);
    printf("version %d emulation %d\n", mvs_version(), mvs_is_emulation());

    for (i = 0; i < B * C * S; ++i) src[i] = (float)(i % 97) - 48.0f;
    CHECK(mvs_pack_c8(src, packed, B, C, S, 0, NULL) == 0);
    CHECK(mvs_unpack_c8(packed, back, B, C, S, 0, NULL) == 0);
    CHECK(memcmp(src, back, sizeof src) == 0);
    /* C8: [B][C/8][S][8] -- channel 9 of item 1 at position 4 sits in block 1, lane 1 */
    CHECK(packed[((1 * (C / 8) + 1) * S + 4) * 8 + 1] == src[(1 * C + 9) * S + 4]);

    /* soft-argmin: one overwhelming plane per column -> depth = that plane, index = its number, confidence = 1 */
    for (b = 0; b < B; ++b)
        for (d = 0; d < D; ++d) planes[b * D + d] = 425.0f + 2.5f * (float)d;
    for (b = 0; b < B; ++b)
        for (d = 0; d < D; ++d)
            for (p = 0; p < H * W; ++p) cost[(b * D + d) * H * W + p] = (d == (b + 2 * p) % D) ? 80.0f : 0.0f;
    CHECK(mvs_softargmin_fwd(cost, planes, 0, depth, index, conf, NULL, B, D, H, W, NULL) == 0);
    for (b = 0; b < B; ++b)
        for (p = 0; p < H * W; ++p) {
            int want = (b + 2 * p) % D;
            CHECK(index[b * H * W + p] == want);
            CHECK(fabsf(depth[b * H * W + p] - planes[b * D + want]) < 1e-3f);
            CHECK(fabsf(conf[b * H * W + p] - 1.0f) < 1e-6f);
        }

    /* errors: a negative code and a message, per thread */
    CHECK(mvs_pack_c8(NULL, packed, B, C, S, 0, NULL) == -1 && strstr(mvs_last_error(), "null") != NULL);
    CHECK(mvs_pack_c8(src, packed, B, 7, S, 0, NULL) < 0 && strlen(mvs_last_error()) > 0);
    CHECK(mvs_compose_proj(src, back, 1, 12, NULL) == -2 && strstr(mvs_last_error(), "views") != NULL);
    CHECK(mvs_set_knob("no_such_knob", 1) < 0);
    printf("ok\n");
    return 0;
}
