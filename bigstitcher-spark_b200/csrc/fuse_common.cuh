// Pieces shared by the two fusion translation units (fuse.cu: generic tile kernel; fuse_tma.cu: TMA-staged
// z-marching kernels).
#pragma once
#include <cmath>

#include "bs_internal.cuh"

#define FUSE_MAX_LUT 256

// cosine blending weight along one axis (l = absolute source coordinate); false when weight is 0
__device__ __forceinline__ bool blend_axis(float l, float dm1, float border, float inv_range, int lut_n,
                                           const float* s_lut, float& w) {
    const float dist = fmaxf(0.f, fminf(l - border, (dm1 - l) - border));
    if (dist == 0.f) return false;
    const float rel = dist * inv_range;
    if (rel < 1.f) {
        float f;
        if (lut_n > 0) {
            const float pos = rel * (float)lut_n;
            const int i = (int)pos;
            const float s = pos - (float)i;
            f = s_lut[i] * (1.0f - s) + s_lut[i + 1] * s;
        } else {
            // (cos((1 - rel) pi) + 1) / 2 == sin^2(pi rel / 2): no cancellation for tiny weights.
            // sin(pi y), y = rel / 2 in [0, 0.5): odd Taylor polynomial to y^11 (rel. error < 1e-7)
            const float yh = 0.5f * rel, y2 = yh * yh;
            float p = -0.0073704309f;              // -pi^11 / 11!
            p = fmaf(p, y2, 0.0821458866f);         //  pi^9 / 9!
            p = fmaf(p, y2, -0.5992645293f);        // -pi^7 / 7!
            p = fmaf(p, y2, 2.5501640399f);         //  pi^5 / 5!
            p = fmaf(p, y2, -5.1677127800f);        // -pi^3 / 3!
            p = fmaf(p, y2, 3.1415926536f);         //  pi
            const float sn = p * yh;
            f = sn * sn;
        }
        w *= f;
    }
    return true;
}


// Same weight as blend_axis, as a factor: 0 when the sample is outside [0, dim-1] or its weight is zero
// (dist == 0), 1 on the plateau.  use_blend == false gives the AVG mask (1 on the closed interval).
__device__ __forceinline__ float blend_factor(float l, float dm1, float border, float inv_range, bool use_blend) {
    if (!(l >= 0.f && l <= dm1)) return 0.f;
    if (!use_blend) return 1.f;
    float w = 1.f;
    if (!blend_axis(l, dm1, border, inv_range, 0, nullptr, w)) return 0.f;
    return w;
}

static inline bool bs_invert34(const double* m, double* inv) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.0 || !std::isfinite(det)) return false;
    const double id = 1.0 / det;
    double A[9] = {(e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id,
                   (f * g - d * i) * id, (a * i - c * g) * id, (c * d - a * f) * id,
                   (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id};
    for (int r = 0; r < 3; ++r) {
        inv[4 * r + 0] = A[3 * r + 0];
        inv[4 * r + 1] = A[3 * r + 1];
        inv[4 * r + 2] = A[3 * r + 2];
        inv[4 * r + 3] = -(A[3 * r + 0] * m[3] + A[3 * r + 1] * m[7] + A[3 * r + 2] * m[11]);
    }
    return true;
}

static inline size_t bs_out_elem_size(int dt) { return dt == BS_DTYPE_F32 ? 4 : dt == BS_DTYPE_U16 ? 2 : 1; }

