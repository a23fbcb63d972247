/*
 * C / OpenMP restatement of oracle/pcm_oracle.py::pcm_shift (phase correlation of two equal-size crops): blended
 * mirrored extension + zero padding, 3-D real FFT (own batched Stockham transform, radices 2/3/4/5, single precision like
 * ComplexFloatType), unit-magnitude normalisation (|c| < 1e-5 -> 0), conj multiply, inverse, periodic 6-neighbour
 * maxima, top-K, quadratic sub-pixel fit, 2^3 wrap candidates, minimum overlap, two-pass Pearson r in double.
 *
 * TEST INFRASTRUCTURE ONLY (checker + the CPU arm of bench.py's `cpu_baseline` / `--impl reference`); never linked or
 * loaded by the product.  PARITY UNPINNED like the numpy oracle it mirrors (the arithmetic of PairwiseStitching.getShift /
 * PhaseCorrelation2, BigStitcher 2.5.0, is not under /root/reference; call site
 * src/main/java/net/preibisch/bigstitcher/spark/SparkPairwiseStitching.java:247-255).
 * tests/test_pcm_oracle.py checks it against the numpy oracle (same integer shift / peak, r 1e-9, sub-pixel 1e-3).
 *
 * Layout of a transform batch: L lines side by side, element e of line l at buf[e * L + l] (re and im in separate
 * arrays), so every butterfly is a unit-stride loop over l that the compiler vectorises.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define LB 16            /* lines per batch */
#define NORM_THRESHOLD 1e-5f

/* physical cores beat hyper-threads on these memory-bound loops: the caller picks the count */
void po_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int po_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int good_fft_size(int n, int even) {
    int m = n < 2 ? 2 : n;
    for (;; ++m) {
        int k = m;
        while (k % 2 == 0) k /= 2;
        while (k % 3 == 0) k /= 3;
        while (k % 5 == 0) k /= 5;
        if (k == 1 && (!even || m % 2 == 0)) return m;
    }
}

/* ---------------------------------------------------------------- batched complex FFT (forward, unnormalised) */
typedef struct {
    int n, nst, radix[32];
    float *twr, *twi; /* exp(-2 pi i k / n), k = 0..n-1 */
} plan_t;

static int plan_make(plan_t* p, int n) {
    p->n = n;
    p->nst = 0;
    int m = n;
    while (m % 4 == 0) { p->radix[p->nst++] = 4; m /= 4; }
    while (m % 2 == 0) { p->radix[p->nst++] = 2; m /= 2; }
    while (m % 3 == 0) { p->radix[p->nst++] = 3; m /= 3; }
    while (m % 5 == 0) { p->radix[p->nst++] = 5; m /= 5; }
    if (m != 1) return -1;
    p->twr = (float*)malloc(sizeof(float) * (size_t)n);
    p->twi = (float*)malloc(sizeof(float) * (size_t)n);
    if (!p->twr || !p->twi) return -2;
    for (int k = 0; k < n; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)n;
        p->twr[k] = (float)cos(a);
        p->twi[k] = (float)sin(a);
    }
    return 0;
}
static void plan_free(plan_t* p) { free(p->twr); free(p->twi); }

/* one Stockham stage of radix R: in -> out, Ls = product of the previous radices */
static void stage(const plan_t* p, int R, int Ls, const float* restrict ir, const float* restrict ii, float* restrict or_,
                  float* restrict oi) {
    const int N = p->n, m = N / R, twstep = N / (Ls * R);
    for (int j = 0; j < m; ++j) {
        const int k = j % Ls;
        const float* xr[5];
        const float* xi[5];
        float wr[5], wi[5];
        for (int q = 0; q < R; ++q) {
            xr[q] = ir + (size_t)(j + q * m) * LB;
            xi[q] = ii + (size_t)(j + q * m) * LB;
            const int t = (q * k * twstep) % N;
            wr[q] = p->twr[t];
            wi[q] = p->twi[t];
        }
        float* yr[5];
        float* yi[5];
        for (int q = 0; q < R; ++q) {
            yr[q] = or_ + (size_t)((j - k) * R + k + q * Ls) * LB;
            yi[q] = oi + (size_t)((j - k) * R + k + q * Ls) * LB;
        }
        for (int l = 0; l < LB; ++l) {
            float ar[5], ai[5];
            for (int q = 0; q < R; ++q) { /* twiddle */
                const float a = xr[q][l], b = xi[q][l];
                ar[q] = a * wr[q] - b * wi[q];
                ai[q] = a * wi[q] + b * wr[q];
            }
            if (R == 2) {
                yr[0][l] = ar[0] + ar[1]; yi[0][l] = ai[0] + ai[1];
                yr[1][l] = ar[0] - ar[1]; yi[1][l] = ai[0] - ai[1];
            } else if (R == 4) {
                const float t0r = ar[0] + ar[2], t0i = ai[0] + ai[2], t1r = ar[0] - ar[2], t1i = ai[0] - ai[2];
                const float t2r = ar[1] + ar[3], t2i = ai[1] + ai[3], t3r = ar[1] - ar[3], t3i = ai[1] - ai[3];
                yr[0][l] = t0r + t2r; yi[0][l] = t0i + t2i;
                yr[2][l] = t0r - t2r; yi[2][l] = t0i - t2i;
                yr[1][l] = t1r + t3i; yi[1][l] = t1i - t3r; /* -i * t3 */
                yr[3][l] = t1r - t3i; yi[3][l] = t1i + t3r;
            } else if (R == 3) {
                const float c = -0.5f, s = -0.86602540378443865f; /* exp(-2 pi i / 3) */
                const float sr = ar[1] + ar[2], si = ai[1] + ai[2], dr = ar[1] - ar[2], di = ai[1] - ai[2];
                yr[0][l] = ar[0] + sr; yi[0][l] = ai[0] + si;
                const float mr = ar[0] + c * sr, mi = ai[0] + c * si;
                yr[1][l] = mr - s * di; yi[1][l] = mi + s * dr;
                yr[2][l] = mr + s * di; yi[2][l] = mi - s * dr;
            } else { /* R == 5 */
                const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f, s1 = -0.95105651629515357f, s2 = -0.58778525229247313f;
                const float a1r = ar[1] + ar[4], a1i = ai[1] + ai[4], b1r = ar[1] - ar[4], b1i = ai[1] - ai[4];
                const float a2r = ar[2] + ar[3], a2i = ai[2] + ai[3], b2r = ar[2] - ar[3], b2i = ai[2] - ai[3];
                yr[0][l] = ar[0] + a1r + a2r; yi[0][l] = ai[0] + a1i + a2i;
                const float m1r = ar[0] + c1 * a1r + c2 * a2r, m1i = ai[0] + c1 * a1i + c2 * a2i;
                const float m2r = ar[0] + c2 * a1r + c1 * a2r, m2i = ai[0] + c2 * a1i + c1 * a2i;
                const float n1r = s1 * b1r + s2 * b2r, n1i = s1 * b1i + s2 * b2i;
                const float n2r = s2 * b1r - s1 * b2r, n2i = s2 * b1i - s1 * b2i;
                /* y[k] = m + i * n  for exp(-i ...) convention: multiply n by i -> (-n_i, n_r) */
                yr[1][l] = m1r - n1i; yi[1][l] = m1i + n1r;
                yr[4][l] = m1r + n1i; yi[4][l] = m1i - n1r;
                yr[2][l] = m2r - n2i; yi[2][l] = m2i + n2r;
                yr[3][l] = m2r + n2i; yi[3][l] = m2i - n2r;
            }
        }
    }
}

/* forward FFT of LB lines in (ar, ai) using (br, bi) as scratch; returns 0 if the result is in a, 1 if in b */
static int fft_batch(const plan_t* p, float* ar, float* ai, float* br, float* bi) {
    int Ls = 1, in_a = 1;
    for (int s = 0; s < p->nst; ++s) {
        const int R = p->radix[s];
        if (in_a) stage(p, R, Ls, ar, ai, br, bi);
        else stage(p, R, Ls, br, bi, ar, ai);
        in_a = !in_a;
        Ls *= R;
    }
    return in_a ? 0 : 1;
}

/* ---------------------------------------------------------------- blended mirrored extension profile */
static void axis_profile(int d, int ext, int P, int* idx, float* w) {
    const int e = ext < d ? ext : d;
    const int period = 2 * d - 2 > 1 ? 2 * d - 2 : 1;
    for (int p = 0; p < P; ++p) {
        idx[p] = 0;
        w[p] = 0.f;
        const int s = p - e;
        if (s < -e || s > d - 1 + e) continue;
        const int dist = s < 0 ? -s : (s > d - 1 ? s - (d - 1) : 0);
        int m = 0;
        if (d > 1) {
            m = s % period;
            if (m < 0) m += period;
            if (m >= d) m = period - m;
        }
        idx[p] = m;
        w[p] = dist > 0 ? (float)(0.5 * (cos(M_PI * (double)dist / (double)e) + 1.0)) : 1.0f;
    }
}

typedef struct { float val; long long idx; } peak_t;
static int peak_better(float v, long long i, float v2, long long i2) { return v > v2 || (v == v2 && i < i2); }

static void solve3(const double H[3][3], const double r[3], double out[3]) {
    const double a = H[0][0], b = H[0][1], c = H[0][2], d = H[1][0], e = H[1][1], f = H[1][2], g = H[2][0], h = H[2][1], i = H[2][2];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    out[0] = out[1] = out[2] = 0.0;
    if (det == 0.0 || !isfinite(det)) return;
    const double x = (r[0] * (e * i - f * h) - b * (r[1] * i - f * r[2]) + c * (r[1] * h - e * r[2])) / det;
    const double y = (a * (r[1] * i - f * r[2]) - r[0] * (d * i - f * g) + c * (d * r[2] - r[1] * g)) / det;
    const double z = (a * (e * r[2] - r[1] * h) - b * (d * r[2] - r[1] * g) + r[0] * (d * h - e * g)) / det;
    if (isfinite(x) && isfinite(y) && isfinite(z)) { out[0] = x; out[1] = y; out[2] = z; }
}

static double pearson_u16(const uint16_t* a, const uint16_t* b, const long long dims[3], const long long o1[3],
                          const long long o2[3], const long long sz[3]) {
    const long long n = sz[0] * sz[1] * sz[2];
    double s1 = 0.0, s2 = 0.0;
#pragma omp parallel for collapse(2) reduction(+ : s1, s2) schedule(static)
    for (long long z = 0; z < sz[2]; ++z)
        for (long long y = 0; y < sz[1]; ++y) {
            const uint16_t* ra = a + ((o1[2] + z) * dims[1] + o1[1] + y) * dims[0] + o1[0];
            const uint16_t* rb = b + ((o2[2] + z) * dims[1] + o2[1] + y) * dims[0] + o2[0];
            double t1 = 0.0, t2 = 0.0;
            for (long long x = 0; x < sz[0]; ++x) { t1 += ra[x]; t2 += rb[x]; }
            s1 += t1; s2 += t2;
        }
    const double m1 = s1 / (double)n, m2 = s2 / (double)n;
    double s11 = 0.0, s22 = 0.0, s12 = 0.0;
#pragma omp parallel for collapse(2) reduction(+ : s11, s22, s12) schedule(static)
    for (long long z = 0; z < sz[2]; ++z)
        for (long long y = 0; y < sz[1]; ++y) {
            const uint16_t* ra = a + ((o1[2] + z) * dims[1] + o1[1] + y) * dims[0] + o1[0];
            const uint16_t* rb = b + ((o2[2] + z) * dims[1] + o2[1] + y) * dims[0] + o2[0];
            double t11 = 0.0, t22 = 0.0, t12 = 0.0;
            for (long long x = 0; x < sz[0]; ++x) {
                const double da = (double)ra[x] - m1, db = (double)rb[x] - m2;
                t11 += da * da; t22 += db * db; t12 += da * db;
            }
            s11 += t11; s22 += t22; s12 += t12;
        }
    if (s11 == 0.0 || s22 == 0.0) return 0.0;
    return s12 / sqrt(s11 * s22);
}

/* out: [0] found, [1..3] shift_int, [4..6] shift_sub, [7] r, [8] n_overlap_px, [9..11] peak index, [12] pcm value, [13..15] pad */
int po_pcm_shift(const uint16_t* img1, const uint16_t* img2, const long long dims[3], int peaks_to_check, int do_subpixel,
                 double min_overlap_frac, const int ext[3], double* out) {
    for (int i = 0; i < 16; ++i) out[i] = 0.0;
    out[7] = -INFINITY;
    int P[3], d[3];
    for (int a = 0; a < 3; ++a) {
        d[a] = (int)dims[a];
        const int es = d[a] + (d[a] < ext[a] ? 2 * d[a] : 2 * ext[a]);
        P[a] = good_fft_size(es, a == 0);
        out[13 + a] = P[a];
    }
    const int Px = P[0], Py = P[1], Pz = P[2], Hx = Px / 2 + 1;
    const size_t pitch = (size_t)((Hx + LB - 1) / LB) * LB;         /* kx padded to whole batches */
    const size_t nspec = (size_t)Pz * Py * pitch;
    float* sr[2] = {NULL, NULL};
    float* si[2] = {NULL, NULL};
    plan_t px, py, pz;
    if (plan_make(&px, Px) || plan_make(&py, Py) || plan_make(&pz, Pz)) return -1;
    int* ix[3];
    float* wx[3];
    for (int a = 0; a < 3; ++a) {
        ix[a] = (int*)malloc(sizeof(int) * (size_t)P[a]);
        wx[a] = (float*)malloc(sizeof(float) * (size_t)P[a]);
        axis_profile(d[a], ext[a], P[a], ix[a], wx[a]);
    }
    for (int im = 0; im < 2; ++im) {
        sr[im] = (float*)calloc(nspec, sizeof(float));
        si[im] = (float*)calloc(nspec, sizeof(float));
        if (!sr[im] || !si[im]) return -2;
    }
    const size_t maxn = (size_t)(Px > Py ? (Px > Pz ? Px : Pz) : (Py > Pz ? Py : Pz));

    /* ---- x pass (real rows as complex lines), both images */
#pragma omp parallel
    {
        float* b0 = (float*)malloc(sizeof(float) * maxn * LB * 4);
        float *ar = b0, *ai = b0 + maxn * LB, *br = b0 + 2 * maxn * LB, *bi = b0 + 3 * maxn * LB;
        const long long nrows = (long long)Pz * Py;
#pragma omp for schedule(dynamic, 4)
        for (long long r0 = 0; r0 < 2 * ((nrows + LB - 1) / LB); ++r0) {
            const int im = (int)(r0 % 2);
            const long long rb = (r0 / 2) * LB;
            const uint16_t* img = im ? img2 : img1;
            int any = 0;
            for (int l = 0; l < LB; ++l) {
                const long long row = rb + l;
                float wyz = 0.f;
                const uint16_t* src = NULL;
                if (row < nrows) {
                    const int z = (int)(row / Py), y = (int)(row % Py);
                    if (wx[2][z] != 0.f && wx[1][y] != 0.f) {
                        src = img + ((size_t)ix[2][z] * d[1] + ix[1][y]) * d[0];
                        wyz = 1.f; /* weights applied below in upstream's order ((1*wx)*wy)*wz */
                        any = 1;
                    }
                }
                const int z = row < nrows ? (int)(row / Py) : 0, y = row < nrows ? (int)(row % Py) : 0;
                for (int x = 0; x < Px; ++x) {
                    float v = 0.f;
                    if (src && wx[0][x] != 0.f) v = (float)src[ix[0][x]] * ((wx[0][x] * wx[1][y]) * wx[2][z]);
                    ar[(size_t)x * LB + l] = v * wyz;
                    ai[(size_t)x * LB + l] = 0.f;
                }
            }
            if (!any) continue; /* all-zero rows stay zero (calloc) */
            const int inb = fft_batch(&px, ar, ai, br, bi);
            const float* rr = inb ? br : ar;
            const float* ri = inb ? bi : ai;
            for (int l = 0; l < LB; ++l) {
                const long long row = rb + l;
                if (row >= nrows) break;
                float* dr = sr[im] + (size_t)row * pitch;
                float* di = si[im] + (size_t)row * pitch;
                for (int k = 0; k < Hx; ++k) { dr[k] = rr[(size_t)k * LB + l]; di[k] = ri[(size_t)k * LB + l]; }
            }
        }
        free(b0);
    }

    /* ---- strided passes: axis 1 (y) then axis 2 (z); conj = 1 transforms conj(data) (used for the inverse) */
#define STRIDED_PASS(SPEC_R, SPEC_I, PLAN, LEN, OUTER, ESTRIDE, OSTRIDE, CONJ)                                              \
    _Pragma("omp parallel") {                                                                                              \
        float* b0 = (float*)malloc(sizeof(float) * maxn * LB * 4);                                                         \
        float *ar = b0, *ai = b0 + maxn * LB, *br = b0 + 2 * maxn * LB, *bi = b0 + 3 * maxn * LB;                          \
        const long long nb = (long long)(OUTER) * (long long)(pitch / LB);                                                 \
        _Pragma("omp for schedule(dynamic, 4)") for (long long t = 0; t < nb; ++t) {                                       \
            const size_t base = (size_t)(t / (long long)(pitch / LB)) * (OSTRIDE) + (size_t)(t % (long long)(pitch / LB)) * LB; \
            for (int e = 0; e < (LEN); ++e) {                                                                              \
                const float* s_r = (SPEC_R) + base + (size_t)e * (ESTRIDE);                                                \
                const float* s_i = (SPEC_I) + base + (size_t)e * (ESTRIDE);                                                \
                for (int l = 0; l < LB; ++l) { ar[(size_t)e * LB + l] = s_r[l]; ai[(size_t)e * LB + l] = (CONJ) ? -s_i[l] : s_i[l]; } \
            }                                                                                                              \
            const int inb = fft_batch(&(PLAN), ar, ai, br, bi);                                                            \
            const float* rr = inb ? br : ar;                                                                               \
            const float* ri = inb ? bi : ai;                                                                               \
            for (int e = 0; e < (LEN); ++e) {                                                                              \
                float* d_r = (SPEC_R) + base + (size_t)e * (ESTRIDE);                                                      \
                float* d_i = (SPEC_I) + base + (size_t)e * (ESTRIDE);                                                      \
                for (int l = 0; l < LB; ++l) { d_r[l] = rr[(size_t)e * LB + l]; d_i[l] = (CONJ) ? -ri[(size_t)e * LB + l] : ri[(size_t)e * LB + l]; } \
            }                                                                                                              \
        }                                                                                                                  \
        free(b0);                                                                                                          \
    }
    for (int im = 0; im < 2; ++im) {
        STRIDED_PASS(sr[im], si[im], py, Py, Pz, pitch, (size_t)Py * pitch, 0)
        STRIDED_PASS(sr[im], si[im], pz, Pz, Py, (size_t)Py * pitch, pitch, 0)
    }
    /* ---- normalise to unit magnitude, conj(second) * first -> spectrum 0 */
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)nspec; ++i) {
        float a_r = sr[0][i], a_i = si[0][i], b_r = sr[1][i], b_i = si[1][i];
        const float ma = sqrtf(a_r * a_r + a_i * a_i), mb = sqrtf(b_r * b_r + b_i * b_i);
        if (ma >= NORM_THRESHOLD) { a_r /= ma; a_i /= ma; } else { a_r = a_i = 0.f; }
        if (mb >= NORM_THRESHOLD) { b_r /= mb; b_i /= mb; } else { b_r = b_i = 0.f; }
        sr[0][i] = a_r * b_r + a_i * b_i; /* a * conj(b) */
        si[0][i] = a_i * b_r - a_r * b_i;
    }
    free(sr[1]); free(si[1]);
    sr[1] = si[1] = NULL;
    /* ---- inverse: z, y (conjugate trick), then x complex-to-real */
    STRIDED_PASS(sr[0], si[0], pz, Pz, Py, (size_t)Py * pitch, pitch, 1)
    STRIDED_PASS(sr[0], si[0], py, Py, Pz, pitch, (size_t)Py * pitch, 1)
    float* pcm = (float*)malloc(sizeof(float) * (size_t)Pz * Py * Px);
    if (!pcm) return -2;
    const float scale = 1.0f / ((float)Px * (float)Py * (float)Pz);
#pragma omp parallel
    {
        float* b0 = (float*)malloc(sizeof(float) * maxn * LB * 4);
        float *ar = b0, *ai = b0 + maxn * LB, *br = b0 + 2 * maxn * LB, *bi = b0 + 3 * maxn * LB;
        const long long nrows = (long long)Pz * Py;
#pragma omp for schedule(dynamic, 4)
        for (long long rb = 0; rb < nrows; rb += LB) {
            for (int l = 0; l < LB; ++l) {
                const long long row = rb + l < nrows ? rb + l : nrows - 1;
                const float* s_r = sr[0] + (size_t)row * pitch;
                const float* s_i = si[0] + (size_t)row * pitch;
                /* full Hermitian spectrum, conjugated (inverse = conj(fft(conj(X)))) */
                for (int k = 0; k < Hx; ++k) { ar[(size_t)k * LB + l] = s_r[k]; ai[(size_t)k * LB + l] = -s_i[k]; }
                for (int k = Hx; k < Px; ++k) { ar[(size_t)k * LB + l] = s_r[Px - k]; ai[(size_t)k * LB + l] = s_i[Px - k]; }
            }
            const int inb = fft_batch(&px, ar, ai, br, bi);
            const float* rr = inb ? br : ar;
            for (int l = 0; l < LB && rb + l < nrows; ++l) {
                float* dst = pcm + (size_t)(rb + l) * Px;
                for (int x = 0; x < Px; ++x) dst[x] = rr[(size_t)x * LB + l] * scale;
            }
        }
        free(b0);
    }
    free(sr[0]); free(si[0]);

    /* ---- periodic 6-neighbour maxima, top-K by (value desc, index asc) */
    const int K = peaks_to_check;
    int nthreads = po_num_threads();
    peak_t* lists = (peak_t*)malloc(sizeof(peak_t) * (size_t)K * (size_t)nthreads);
    int* counts = (int*)calloc((size_t)nthreads, sizeof(int));
#pragma omp parallel
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        peak_t* mine = lists + (size_t)tid * K;
        int cnt = 0;
#pragma omp for schedule(static)
        for (int z = 0; z < Pz; ++z) {
            const int zm = (z + Pz - 1) % Pz, zp = (z + 1) % Pz;
            for (int y = 0; y < Py; ++y) {
                const int ym = (y + Py - 1) % Py, yp = (y + 1) % Py;
                const float* c = pcm + ((size_t)z * Py + y) * Px;
                const float* cym = pcm + ((size_t)z * Py + ym) * Px;
                const float* cyp = pcm + ((size_t)z * Py + yp) * Px;
                const float* czm = pcm + ((size_t)zm * Py + y) * Px;
                const float* czp = pcm + ((size_t)zp * Py + y) * Px;
                for (int x = 0; x < Px; ++x) {
                    const float v = c[x];
                    if (cnt == K && v < mine[K - 1].val) continue; /* cannot enter the list */
                    if (v < c[(x + Px - 1) % Px] || v < c[(x + 1) % Px] || v < cym[x] || v < cyp[x] || v < czm[x] || v < czp[x]) continue;
                    const long long li = ((long long)z * Py + y) * Px + x;
                    if (cnt == K && !peak_better(v, li, mine[K - 1].val, mine[K - 1].idx)) continue;
                    int pos = cnt < K ? cnt : K - 1;
                    while (pos > 0 && peak_better(v, li, mine[pos - 1].val, mine[pos - 1].idx)) { mine[pos] = mine[pos - 1]; --pos; }
                    mine[pos].val = v;
                    mine[pos].idx = li;
                    if (cnt < K) ++cnt;
                }
            }
        }
        counts[tid] = cnt;
    }
    peak_t best[64];
    int nbest = 0;
    for (int t = 0; t < nthreads; ++t)
        for (int i = 0; i < counts[t]; ++i) {
            const peak_t c = lists[(size_t)t * K + i];
            int pos = nbest < K ? nbest : K - 1;
            if (nbest == K && !peak_better(c.val, c.idx, best[K - 1].val, best[K - 1].idx)) continue;
            while (pos > 0 && peak_better(c.val, c.idx, best[pos - 1].val, best[pos - 1].idx)) { best[pos] = best[pos - 1]; --pos; }
            best[pos] = c;
            if (nbest < K) ++nbest;
        }
    free(lists); free(counts);

    /* ---- candidates: 8 wrap variants per peak, minimum overlap, Pearson; stable best by (r desc, npx desc, order) */
    const long long n_px = dims[0] * dims[1] * dims[2];
    const long long min_px = (long long)((double)n_px * min_overlap_frac);
    double best_r = -INFINITY;
    long long best_npx = -1;
    int have = 0;
    for (int pi = 0; pi < nbest; ++pi) {
        const long long li = best[pi].idx;
        const long long loc[3] = {li % Px, (li / Px) % Py, li / ((long long)Px * Py)};
        double sub[3] = {0, 0, 0};
        if (do_subpixel) {
            double f[3][3][3];
            for (int dz = -1; dz <= 1; ++dz)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx)
                        f[dz + 1][dy + 1][dx + 1] = pcm[(((loc[2] + dz + Pz) % Pz) * Py + (loc[1] + dy + Py) % Py) * Px + (loc[0] + dx + Px) % Px];
            const double c = f[1][1][1];
            const double g[3] = {-(f[1][1][2] - f[1][1][0]) / 2.0, -(f[1][2][1] - f[1][0][1]) / 2.0, -(f[2][1][1] - f[0][1][1]) / 2.0};
            double H[3][3];
            H[0][0] = f[1][1][2] - 2 * c + f[1][1][0];
            H[1][1] = f[1][2][1] - 2 * c + f[1][0][1];
            H[2][2] = f[2][1][1] - 2 * c + f[0][1][1];
            H[0][1] = H[1][0] = (f[1][2][2] - f[1][2][0] - f[1][0][2] + f[1][0][0]) / 4.0;
            H[0][2] = H[2][0] = (f[2][1][2] - f[2][1][0] - f[0][1][2] + f[0][1][0]) / 4.0;
            H[1][2] = H[2][1] = (f[2][2][1] - f[2][0][1] - f[0][2][1] + f[0][0][1]) / 4.0;
            solve3(H, g, sub);
        }
        for (int i = 0; i < 8; ++i) {
            long long s[3], o1[3], o2[3], sz[3], npx = 1;
            int overlap = 1;
            for (int a = 0; a < 3; ++a) {
                s[a] = loc[a];
                if (((i >> a) & 1) == 0) s[a] = s[a] < 0 ? s[a] + P[a] : s[a] - P[a];
                const long long n = dims[a];
                if (s[a] >= 0) {
                    if (s[a] >= n) { overlap = 0; break; }
                    o1[a] = s[a]; o2[a] = 0; sz[a] = n - s[a] < n ? n - s[a] : n;
                } else {
                    if (s[a] <= -n) { overlap = 0; break; }
                    o1[a] = 0; o2[a] = -s[a]; sz[a] = n + s[a] < n ? n + s[a] : n;
                }
                npx *= sz[a];
            }
            double r = -INFINITY;
            if (overlap && npx >= min_px) r = pearson_u16(img1, img2, dims, o1, o2, sz);
            else npx = 0;
            const int better = !have || r > best_r || (r == best_r && npx > best_npx);
            if (better) {
                have = 1;
                best_r = r;
                best_npx = npx;
                for (int a = 0; a < 3; ++a) {
                    out[1 + a] = (double)s[a];
                    out[4 + a] = (double)s[a] + (do_subpixel ? sub[a] : 0.0);
                    out[9 + a] = (double)loc[a];
                }
                out[12] = best[pi].val;
            }
        }
    }
    if (have && isfinite(best_r)) {
        out[0] = 1.0;
        out[7] = best_r;
        out[8] = (double)best_npx;
    } else {
        out[0] = 0.0;
        out[7] = -INFINITY;
    }
    free(pcm);
    for (int a = 0; a < 3; ++a) { free(ix[a]); free(wx[a]); }
    plan_free(&px); plan_free(&py); plan_free(&pz);
    return 0;
}
