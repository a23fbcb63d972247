/*
 * C / OpenMP restatement of oracle/fusion_oracle.py::fuse_block for the weighted-average fusion
 * types (AVG, AVG_BLEND), n-linear interpolation, float32 output.
 *
 * TEST INFRASTRUCTURE ONLY (checker + the CPU arm of bench.py's `cpu_baseline`); never linked or
 * loaded by the product.  PARITY UNPINNED like the numpy oracle it mirrors: the arithmetic of
 * BlkAffineFusion (multiview-reconstruction 8.0.0) is not under /root/reference; this follows the
 * call-site contract src/main/java/net/preibisch/bigstitcher/spark/SparkAffineFusion.java:602-627
 * and SURVEY.md Appendix A.2.  tests/test_fusion_oracle.py checks it against the numpy oracle.
 *
 * Per output voxel and view: world -> source in double (same association as the numpy code), cast
 * to float32; inside test on the closed interval [0, dim-1]; cosine blending weight per axis
 * (dist == 0 -> 0; (cos((1-rel)*pi)+1)/2 in double, product kept in float32); trilinear taps with
 * border clamp, x then y then z as a + f*(b-a) in float32; out = sum w*I / sum w.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static int invert34(const double* m, double* inv) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.0) return -1;
    const double id = 1.0 / det;
    const double A[9] = {(e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id,
                         (f * g - d * i) * id, (a * i - c * g) * id, (c * d - a * f) * id,
                         (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id};
    for (int r = 0; r < 3; ++r) {
        inv[4 * r + 0] = A[3 * r + 0];
        inv[4 * r + 1] = A[3 * r + 1];
        inv[4 * r + 2] = A[3 * r + 2];
        inv[4 * r + 3] = -(A[3 * r + 0] * m[3] + A[3 * r + 1] * m[7] + A[3 * r + 2] * m[11]);
    }
    return 0;
}

static inline float voxel(const void* img, int dtype, size_t i) {
    if (dtype == 0) return (float)((const uint16_t*)img)[i];
    if (dtype == 1) return ((const float*)img)[i];
    return (float)((const uint8_t*)img)[i];
}

static inline long long clampll(long long v, long long hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

/* physical cores beat hyper-threads on these memory-bound loops: the caller picks the count */
void fo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int fo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* dtypes: 0 u16, 1 f32, 2 u8; dims n*3 {x,y,z}; src_to_world n*12; border/range n*3; fusion_type 0 AVG, 1 AVG_BLEND */
int fo_fuse_block(int n_views, const void* const* imgs, const int* dtypes, const long long* dims,
                  const double* src_to_world, const float* border, const float* range, const long long bmin[3],
                  const long long bsize[3], int fusion_type, float* out) {
    if (fusion_type != 0 && fusion_type != 1) return -2;
    double* inv = (double*)malloc(sizeof(double) * 12 * (size_t)(n_views > 0 ? n_views : 1));
    if (!inv) return -3;
    for (int v = 0; v < n_views; ++v)
        if (invert34(src_to_world + 12 * v, inv + 12 * v)) { free(inv); return -1; }
    const long long bx = bsize[0], by = bsize[1], bz = bsize[2];
#pragma omp parallel for collapse(2) schedule(static)
    for (long long k = 0; k < bz; ++k) {
        for (long long j = 0; j < by; ++j) {
            const double Z = (double)(bmin[2] + k), Y = (double)(bmin[1] + j);
            float* row = out + ((size_t)k * by + j) * bx;
            for (long long i = 0; i < bx; ++i) {
                const double X = (double)(bmin[0] + i);
                float sum_i = 0.f, sum_w = 0.f;
                for (int v = 0; v < n_views; ++v) {
                    const double* m = inv + 12 * v;
                    const long long dx = dims[3 * v], dy = dims[3 * v + 1], dz = dims[3 * v + 2];
                    float s[3];
                    for (int r = 0; r < 3; ++r)
                        s[r] = (float)(m[4 * r] * X + (m[4 * r + 1] * Y + (m[4 * r + 2] * Z + m[4 * r + 3])));
                    if (!(s[0] >= 0.f && s[0] <= (float)(dx - 1) && s[1] >= 0.f && s[1] <= (float)(dy - 1) &&
                          s[2] >= 0.f && s[2] <= (float)(dz - 1)))
                        continue;
                    float w = 1.f;
                    if (fusion_type == 1) {
                        int zero = 0;
                        for (int r = 0; r < 3; ++r) {
                            const float dm1 = (float)(dims[3 * v + r] - 1), bo = border[3 * v + r];
                            float dist = fminf(s[r] - bo, (dm1 - s[r]) - bo);
                            if (dist < 0.f) dist = 0.f;
                            if (dist == 0.f) zero = 1;
                            const float rel = dist / range[3 * v + r];
                            if (rel < 1.f) w = (float)((double)w * ((cos((1.0 - (double)rel) * M_PI) + 1.0) / 2.0));
                        }
                        if (zero) w = 0.f;
                    }
                    const float fx = floorf(s[0]), fy = floorf(s[1]), fz = floorf(s[2]);
                    const float rx = s[0] - fx, ry = s[1] - fy, rz = s[2] - fz;
                    const long long x0 = clampll((long long)fx, dx - 1), x1 = clampll((long long)fx + 1, dx - 1);
                    const long long y0 = clampll((long long)fy, dy - 1), y1 = clampll((long long)fy + 1, dy - 1);
                    const long long z0 = clampll((long long)fz, dz - 1), z1 = clampll((long long)fz + 1, dz - 1);
                    const void* img = imgs[v];
                    const int dt = dtypes[v];
#define VX(zz, yy, xx) voxel(img, dt, ((size_t)(zz) * dy + (yy)) * dx + (xx))
                    const float c00 = VX(z0, y0, x0) + rx * (VX(z0, y0, x1) - VX(z0, y0, x0));
                    const float c01 = VX(z0, y1, x0) + rx * (VX(z0, y1, x1) - VX(z0, y1, x0));
                    const float c10 = VX(z1, y0, x0) + rx * (VX(z1, y0, x1) - VX(z1, y0, x0));
                    const float c11 = VX(z1, y1, x0) + rx * (VX(z1, y1, x1) - VX(z1, y1, x0));
#undef VX
                    const float c0 = c00 + ry * (c01 - c00);
                    const float c1 = c10 + ry * (c11 - c10);
                    const float val = c0 + rz * (c1 - c0);
                    sum_i = sum_i + w * val;
                    sum_w = sum_w + w;
                }
                row[i] = sum_w > 0.f ? sum_i / sum_w : 0.f;
            }
        }
    }
    free(inv);
    return 0;
}
