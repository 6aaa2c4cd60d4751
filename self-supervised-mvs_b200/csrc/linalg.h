// linalg.h - tiny fp64 matrix inverses used by the camera-algebra prep kernels.
#pragma once
#include "mvs_rt.h"

__device__ static inline void inv3(const double* a, double* o) {
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    const double id = 1.0 / det;
    o[0] = c00 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c01 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c02 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}


// 4x4 inverse by Gauss-Jordan with partial pivoting in fp64 (the reference uses an fp32 LU; the fp64 result
// agrees with it to fp32 rounding).
__device__ static inline void inv4(const double* a, double* inv) {
    double m[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { m[r][c] = a[r * 4 + c]; m[r][c + 4] = (r == c) ? 1.0 : 0.0; }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        double best = fabs(m[col][col]);
        for (int r = col + 1; r < 4; ++r) if (fabs(m[r][col]) > best) { best = fabs(m[r][col]); piv = r; }
        if (piv != col) for (int c = 0; c < 8; ++c) { const double t = m[col][c]; m[col][c] = m[piv][c]; m[piv][c] = t; }
        const double d = 1.0 / m[col][col];
        for (int c = 0; c < 8; ++c) m[col][c] *= d;
        for (int r = 0; r < 4; ++r) if (r != col) {
            const double f = m[r][col];
            for (int c = 0; c < 8; ++c) m[r][c] -= f * m[col][c];
        }
    }
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) inv[r * 4 + c] = m[r][c + 4];
}

