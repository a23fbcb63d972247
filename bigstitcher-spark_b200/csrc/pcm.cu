// Hot path 1: phase correlation of one overlap-cropped tile pair, entirely on the device.
// Replaces TransformationTools.computeStitching's numeric core
// (PairwiseStitching.getShift -> PhaseCorrelation2.calculatePCM + getShift; call site
// J/SparkPairwiseStitching.java:247-255).  No cuFFT: the 3-D real FFT is five hand-written
// passes over an x-fastest half spectrum S[Pz][Py][pitch] (complex64, pitch = M+1 rounded up
// to 16 so every row is 128 B aligned):
//
//   k_fft_x_r2c      uint16/float crop -> blended mirrored extension + zero pad -> R2C along x
//   k_fft_strided    (mode 0) forward FFT along y, in place, both spectra in one launch
//   k_fft_strided    (mode 1) forward z FFT of A and B, unit-magnitude normalisation,
//                    conj(A)*B, forward z FFT of the product      [reads 2S, writes S]
//   k_fft_strided    (mode 0) forward y FFT of the product
//   k_fft_x_c2r      conj + C2R along x, in place -> real PCM (row pitch 2*pitch floats)
//   k_peaks          periodic 6-neighbour local maxima, per-CTA top-K
//   k_gather27       3x3x3 neighbourhoods of the K peaks (sub-pixel fit runs on the host)
//   k_pearson        exact integer sums (uint64 atomics) for all surviving wrap candidates
//
// Inverse transforms are forward transforms of conjugated data (conj(F(conj X)) = N F^-1 X);
// the two conjugations between consecutive passes cancel, so only the product step and the
// C2R load conjugate.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "bs_internal.cuh"
#include "pcm_fft.cuh"

#define PCM_THREADS 256
#define PCM_KMAX 32
#define PCM_SMEM_MAX 232448  // 227 KB opt-in limit per CTA on sm_100

extern __shared__ __align__(16) float2 bs_sm[];

// ------------------------------------------------------------------------------------------
// x pass, real -> complex
struct XR2CArgs {
    const void* img[2];
    float2* spec[2];
    int dtype;
    int dx, dy, dz;
    int Px, Py, Pz, M, pitch;
    int ex;        // offset of the crop inside the padded x axis (= min(extension, dx))
    int Ex, Ey, Ez;
    const int* idx_x; const float* w_x;
    const int* idx_y; const float* w_y;
    const int* idx_z; const float* w_z;
    const float2* tw;
    FftPlan plan;
    int lshift;
};

__device__ __forceinline__ float load_voxel(const void* p, int dtype, size_t i) {
    if (dtype == BS_DTYPE_U16) return (float)__ldg((const unsigned short*)p + i);
    if (dtype == BS_DTYPE_F32) return __ldg((const float*)p + i);
    return (float)__ldg((const unsigned char*)p + i);
}

// One warp owns a line at a time: the row's y/z profile is warp-uniform, interior samples are
// read as aligned ushort2 pairs, and each lane keeps XR_UNROLL independent loads in flight.
#define XR_UNROLL 5
template <class F>
__global__ void __launch_bounds__(PCM_THREADS) k_fft_x_r2c(const __grid_constant__ XR2CArgs a) {
    const int M = F::kStatic ? F::N : a.M;
    const int lshift = F::kStatic ? F::LSHIFT : a.lshift;
    const int LB = 1 << lshift, ls = LB + 1;
    const int Px = 2 * M;
    const int pitch = F::kStatic ? ((F::N + 1 + 15) / 16) * 16 : a.pitch;
    float2* tw = bs_sm;
    float2* b0 = bs_sm + Px;
    float2* b1 = b0 + M * ls;
    const int zp = blockIdx.y, y0 = blockIdx.x * LB, im = blockIdx.z;
    float2* __restrict__ spec = a.spec[im];
    const size_t rowbase = ((size_t)zp * a.Py + y0) * pitch;
    const int nlines = min(LB, a.Py - y0);
    if (zp >= a.Ez || y0 >= a.Ey) {  // every line of this CTA lies in the zero padding
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4* d4 = reinterpret_cast<float4*>(spec + rowbase);
        for (int i = threadIdx.x; i < nlines * (pitch >> 1); i += blockDim.x) d4[i] = z4;
        return;
    }
    for (int i = threadIdx.x; i < Px; i += blockDim.x) tw[i] = a.tw[i];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NW = PCM_THREADS / 32;
    const float wz = a.w_z[zp];
    const int sz = a.idx_z[zp];
    const void* img = a.img[im];
    const int e0 = a.ex;
    const bool vec_ok = a.dtype == BS_DTYPE_U16 && !(e0 & 1) && !(a.dx & 1) && !((size_t)img & 3);
    // two lines per warp pass -> 2 * XR_UNROLL independent loads in flight per lane
    for (int lA = wid; lA < LB; lA += 2 * NW) {
        bool live[2];
        float gyz[2], wy[2];
        size_t rb[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int l = lA + h * NW, yp = y0 + l;
            live[h] = l < LB && l < nlines && yp < a.Ey;
            wy[h] = 0.f;
            rb[h] = 0;
            if (live[h]) {
                wy[h] = a.w_y[yp];
                rb[h] = ((size_t)sz * a.dy + a.idx_y[yp]) * a.dx;
            }
            gyz[h] = wy[h] * wz;  // == (1 * wy) * wz for interior samples
        }
        for (int n0 = 0; n0 < M; n0 += 32 * XR_UNROLL) {
            float2 v[2][XR_UNROLL];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int u = 0; u < XR_UNROLL; ++u) {
                    const int n = n0 + u * 32 + lane;
                    v[h][u] = make_float2(0.f, 0.f);
                    if (live[h] && n < M) {
                        const int xp = 2 * n;
                        if (xp >= e0 && xp + 1 < e0 + a.dx) {
                            if (vec_ok) {
                                const ushort2 t = __ldg(reinterpret_cast<const ushort2*>(
                                    (const unsigned short*)img + rb[h] + (xp - e0)));
                                v[h][u] = make_float2((float)t.x * gyz[h], (float)t.y * gyz[h]);
                            } else {
                                v[h][u] = make_float2(load_voxel(img, a.dtype, rb[h] + (xp - e0)) * gyz[h],
                                                      load_voxel(img, a.dtype, rb[h] + (xp - e0) + 1) * gyz[h]);
                            }
                        } else if (xp < a.Ex) {  // blended mirrored margin (a few samples per line)
                            const float g0 = (a.w_x[xp] * wy[h]) * wz, g1 = (a.w_x[xp + 1] * wy[h]) * wz;
                            if (g0 != 0.f) v[h][u].x = load_voxel(img, a.dtype, rb[h] + a.idx_x[xp]) * g0;
                            if (g1 != 0.f) v[h][u].y = load_voxel(img, a.dtype, rb[h] + a.idx_x[xp + 1]) * g1;
                        }
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int l = lA + h * NW;
                if (l < LB) {
#pragma unroll
                    for (int u = 0; u < XR_UNROLL; ++u) {
                        const int n = n0 + u * 32 + lane;
                        if (n < M) b0[n * ls + l] = v[h][u];
                    }
                }
            }
        }
    }
    __syncthreads();
    const float2* res = F::run(b0, b1, tw, a.plan, lshift, ls, 2);
    // untangle the packed half-length transform into the real-input spectrum X[0..M]
    for (int l = wid; l < nlines; l += NW) {
        float2* srow = spec + rowbase + (size_t)l * pitch;
        for (int k = lane; k < pitch; k += 32) {
            float2 X = make_float2(0.f, 0.f);
            if (k <= M) {
                const int k0 = (k == M) ? 0 : k;
                const int k1 = (k == 0 || k == M) ? 0 : M - k;
                const float2 Zk = res[k0 * ls + l];
                float2 Zm = res[k1 * ls + l];
                Zm.y = -Zm.y;
                const float2 E = make_float2(0.5f * (Zk.x + Zm.x), 0.5f * (Zk.y + Zm.y));
                const float2 D = make_float2(0.5f * (Zk.x - Zm.x), 0.5f * (Zk.y - Zm.y));
                const float2 wD = cmulf(tw[k], D);
                X = make_float2(E.x + wD.y, E.y - wD.x);  // E - i*w*D
            }
            __stcg(srow + k, X);
        }
    }
}

// ------------------------------------------------------------------------------------------
// x pass, real -> complex, persistent + TMA: each CTA loops over line groups; the raw uint16 /
// float rows of the NEXT group are pulled into shared memory by the TMA unit (1-D bulk copies
// completing on an mbarrier) while the current group is windowed, transformed and stored.
__device__ __forceinline__ unsigned int smem_u32(const void* p) {
    return (unsigned int)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct XR2CTmaArgs {
    XR2CArgs x;
    int n_groups;      // line groups per plane
    int n_items;       // 2 * Pz * n_groups
    int row_bytes;     // dx * element size (multiple of 16)
    int esize;
};

template <class F>
__global__ void __launch_bounds__(PCM_THREADS) k_fft_x_r2c_tma(const __grid_constant__ XR2CTmaArgs t) {
    const XR2CArgs& a = t.x;
    const int M = F::kStatic ? F::N : a.M;
    const int lshift = F::kStatic ? F::LSHIFT : a.lshift;
    const int LB = 1 << lshift, ls = LB + 1;
    const int Px = 2 * M;
    const int pitch = F::kStatic ? ((F::N + 1 + 15) / 16) * 16 : a.pitch;
    float2* tw = bs_sm;
    float2* b0 = bs_sm + Px;
    float2* b1 = b0 + M * ls;
    unsigned char* raw = reinterpret_cast<unsigned char*>(b1 + M * ls);     // 2 * LB * row_bytes, 16-B aligned
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(raw + 2 * (size_t)LB * t.row_bytes);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NW = PCM_THREADS / 32;

    for (int i = threadIdx.x; i < Px; i += blockDim.x) tw[i] = a.tw[i];
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // item -> (image, plane, line group); consecutive items share the plane
    auto decode = [&](int w, int& im, int& zp, int& y0) {
        const int g = w % t.n_groups;
        const int r = w / t.n_groups;
        zp = r % a.Pz;
        im = r / a.Pz;
        y0 = g * LB;
    };
    // producer: issue the bulk copies of item w into stage buffer `buf` (thread 0 only)
    auto prefetch = [&](int w, int buf) {
        int im, zp, y0;
        decode(w, im, zp, y0);
        if (zp >= a.Ez || y0 >= a.Ey) return;
        const int nl = min(min(LB, a.Py - y0), a.Ey - y0);
        mbar_expect_tx(&bars[buf], (unsigned int)(nl * t.row_bytes));
        const unsigned char* img = reinterpret_cast<const unsigned char*>(a.img[im]);
        const int sz = a.idx_z[zp];
        for (int l = 0; l < nl; ++l) {
            const size_t row = (size_t)sz * a.dy + a.idx_y[y0 + l];
            tma_bulk_g2s(raw + ((size_t)buf * LB + l) * t.row_bytes, img + row * t.row_bytes, (unsigned int)t.row_bytes,
                         &bars[buf]);
        }
    };

    unsigned int phase[2] = {0u, 0u};
    int it = 0;
    if (threadIdx.x == 0 && (int)blockIdx.x < t.n_items) prefetch(blockIdx.x, 0);
    for (int w = blockIdx.x; w < t.n_items; w += gridDim.x, ++it) {
        const int buf = it & 1;
        int im, zp, y0;
        decode(w, im, zp, y0);
        const int wn = w + gridDim.x;
        if (threadIdx.x == 0 && wn < t.n_items) prefetch(wn, buf ^ 1);
        float2* __restrict__ spec = a.spec[im];
        const size_t rowbase = ((size_t)zp * a.Py + y0) * pitch;
        const int nlines = min(LB, a.Py - y0);
        if (zp >= a.Ez || y0 >= a.Ey) {  // zero padding: no loads were issued for this item
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4* d4 = reinterpret_cast<float4*>(spec + rowbase);
            for (int i = threadIdx.x; i < nlines * (pitch >> 1); i += blockDim.x) d4[i] = z4;
            continue;
        }
        mbar_wait(&bars[buf], phase[buf]);
        phase[buf] ^= 1u;
        const float wz = a.w_z[zp];
        const int e0 = a.ex;
        const unsigned char* rb = raw + (size_t)buf * LB * t.row_bytes;
        // window + pack from the staged rows
        for (int l = wid; l < LB; l += NW) {
            const int yp = y0 + l;
            const bool live = l < nlines && yp < a.Ey;
            const float wy = live ? a.w_y[yp] : 0.f;
            const float gyz = wy * wz;
            const unsigned char* row = rb + (size_t)l * t.row_bytes;
            for (int n = lane; n < M; n += 32) {
                float2 v = make_float2(0.f, 0.f);
                if (live) {
                    const int xp = 2 * n;
                    if (xp >= e0 && xp + 1 < e0 + a.dx) {
                        const int xs = xp - e0;
                        float p0, p1;
                        if (a.dtype == BS_DTYPE_U16) {
                            p0 = (float)reinterpret_cast<const unsigned short*>(row)[xs];
                            p1 = (float)reinterpret_cast<const unsigned short*>(row)[xs + 1];
                        } else if (a.dtype == BS_DTYPE_F32) {
                            p0 = reinterpret_cast<const float*>(row)[xs];
                            p1 = reinterpret_cast<const float*>(row)[xs + 1];
                        } else {
                            p0 = (float)row[xs];
                            p1 = (float)row[xs + 1];
                        }
                        v = make_float2(p0 * gyz, p1 * gyz);
                    } else if (xp < a.Ex) {
                        const float g0 = (a.w_x[xp] * wy) * wz, g1 = (a.w_x[xp + 1] * wy) * wz;
                        const int i0 = a.idx_x[xp], i1 = a.idx_x[xp + 1];
                        float p0, p1;
                        if (a.dtype == BS_DTYPE_U16) {
                            p0 = (float)reinterpret_cast<const unsigned short*>(row)[i0];
                            p1 = (float)reinterpret_cast<const unsigned short*>(row)[i1];
                        } else if (a.dtype == BS_DTYPE_F32) {
                            p0 = reinterpret_cast<const float*>(row)[i0];
                            p1 = reinterpret_cast<const float*>(row)[i1];
                        } else {
                            p0 = (float)row[i0];
                            p1 = (float)row[i1];
                        }
                        v = make_float2(g0 != 0.f ? p0 * g0 : 0.f, g1 != 0.f ? p1 * g1 : 0.f);
                    }
                }
                b0[n * ls + l] = v;
            }
        }
        __syncthreads();
        const float2* res = F::run(b0, b1, tw, a.plan, lshift, ls, 2);
        for (int l = wid; l < nlines; l += NW) {
            float2* srow = spec + rowbase + (size_t)l * pitch;
            for (int k = lane; k < pitch; k += 32) {
                float2 X = make_float2(0.f, 0.f);
                if (k <= M) {
                    const int k0 = (k == M) ? 0 : k;
                    const int k1 = (k == 0 || k == M) ? 0 : M - k;
                    const float2 Zk = res[k0 * ls + l];
                    float2 Zm = res[k1 * ls + l];
                    Zm.y = -Zm.y;
                    const float2 E = make_float2(0.5f * (Zk.x + Zm.x), 0.5f * (Zk.y + Zm.y));
                    const float2 D = make_float2(0.5f * (Zk.x - Zm.x), 0.5f * (Zk.y - Zm.y));
                    const float2 wD = cmulf(tw[k], D);
                    X = make_float2(E.x + wD.y, E.y - wD.x);
                }
                __stcg(srow + k, X);
            }
        }
        __syncthreads();  // b0/b1 and the consumed stage buffer may be overwritten from here on
    }
}

// ------------------------------------------------------------------------------------------
struct XC2RArgs {
    float2* spec;
    int Px, Py, Pz, M, pitch;
    const float2* tw;
    FftPlan plan;
    int lshift;
    float scale;
};

// Warp-private x passes: one warp owns one line at a time (window, FFT, untangle, store) with
// only __syncwarp() between steps; the raw row of the warp's NEXT line is fetched by a 1-D bulk
// TMA copy into the warp's own double buffer (own mbarrier), so there is no block-wide barrier
// in the steady state and the 8 warps of a CTA overlap each other's memory and math phases.
struct XWArgs {
    XR2CArgs x;
    int row_bytes;     // dx * element size; multiple of 16 when use_tma
    int use_tma;
    long long n_lines; // 2 * Pz * Py
};

__device__ __forceinline__ float raw_elem(const unsigned char* row, int dtype, int i) {
    if (dtype == BS_DTYPE_U16) return (float)reinterpret_cast<const unsigned short*>(row)[i];
    if (dtype == BS_DTYPE_F32) return reinterpret_cast<const float*>(row)[i];
    return (float)row[i];
}

template <class F>
__global__ void __launch_bounds__(PCM_THREADS, 4) k_fft_x_r2c_w(const __grid_constant__ XWArgs t) {
    const XR2CArgs& a = t.x;
    const int M = F::kStatic ? F::N : a.M;
    const int Px = 2 * M;
    const int pitch = F::kStatic ? ((F::N + 1 + 15) / 16) * 16 : a.pitch;
    constexpr int NW = PCM_THREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // smem: tw[Px] | per warp: A[M], B[M] | per warp raw[2][row_bytes] | per warp mbar[2]
    float2* tw = bs_sm;
    float2* A = bs_sm + Px + (size_t)wid * 2 * M;
    float2* B = A + M;
    unsigned char* raw = reinterpret_cast<unsigned char*>(bs_sm + Px + (size_t)NW * 2 * M) + (size_t)wid * 2 * t.row_bytes;
    unsigned long long* bars =
        reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(bs_sm + Px + (size_t)NW * 2 * M) +
                                              (size_t)NW * 2 * t.row_bytes) + 2 * wid;
    constexpr int UTW = F::N / 2 + 1;        // untangle twiddles tw[0 .. M/2]
    if constexpr (F::kSmemTw) {
        // the Px-entry table area holds: untangle twiddles | compact per-stage tables (conflict-free reads)
        static_assert(UTW + F::Tw::total <= 2 * F::N, "stage tables must fit the twiddle area");
        for (int i = threadIdx.x; i < UTW; i += blockDim.x) tw[i] = a.tw[i];
        F::Tw::build(tw + UTW, a.tw, threadIdx.x, blockDim.x);
    } else {
        for (int i = threadIdx.x; i < Px; i += blockDim.x) tw[i] = a.tw[i];
    }
    if (t.use_tma && lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // static plan: FFT twiddles and the untangle twiddles of this lane live in registers for the whole kernel
    typename F::Tw twr;
    twr.init(F::kSmemTw ? tw + UTW : tw, lane);
    constexpr bool kRegU = F::kStatic && !F::kSmemTw;
    constexpr int UIT = F::kStatic ? (F::N / 2 + 1 + 31) / 32 : 1;
    float2 utw[kRegU ? UIT : 1];
    if constexpr (kRegU) {
#pragma unroll
        for (int i = 0; i < UIT; ++i) utw[i] = tw[min(lane + 32 * i, M)];
    }

    const long long gw = (long long)blockIdx.x * NW + wid, gstride = (long long)gridDim.x * NW;
    const int e0 = a.ex;
    // line -> (image, plane, row); returns the source row index or -1 for an all-zero line
    auto src_row = [&](long long line, int& im, int& zp, int& yp) -> long long {
        yp = (int)(line % a.Py);
        const long long r = line / a.Py;
        zp = (int)(r % a.Pz);
        im = (int)(r / a.Pz);
        if (zp >= a.Ez || yp >= a.Ey) return -1;
        return (long long)a.idx_z[zp] * a.dy + a.idx_y[yp];
    };
    auto prefetch = [&](long long line, int buf) {  // lane 0 only
        int im, zp, yp;
        const long long row = src_row(line, im, zp, yp);
        if (row < 0) return;
        const unsigned char* img = reinterpret_cast<const unsigned char*>(a.img[im]);
        mbar_expect_tx(&bars[buf], (unsigned int)t.row_bytes);
        tma_bulk_g2s(raw + (size_t)buf * t.row_bytes, img + (size_t)row * t.row_bytes, (unsigned int)t.row_bytes, &bars[buf]);
    };

    unsigned int phase0 = 0u, phase1 = 0u;
    int it = 0;
    if (t.use_tma && lane == 0 && gw < t.n_lines) prefetch(gw, 0);
    for (long long line = gw; line < t.n_lines; line += gstride, ++it) {
        const int buf = it & 1;
        if (t.use_tma && lane == 0 && line + gstride < t.n_lines) prefetch(line + gstride, buf ^ 1);
        int im, zp, yp;
        const long long row = src_row(line, im, zp, yp);
        float2* srow = a.spec[im] + ((size_t)zp * a.Py + yp) * pitch;
        if (row < 0) {
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = lane; i < (pitch >> 1); i += 32) reinterpret_cast<float4*>(srow)[i] = z4;
            continue;
        }
        const float wy = a.w_y[yp], wz = a.w_z[zp];
        const float gyz = wy * wz;
        const unsigned char* rrow;
        if (t.use_tma) {
            mbar_wait(&bars[buf], buf ? phase1 : phase0);
            if (buf) phase1 ^= 1u; else phase0 ^= 1u;
            rrow = raw + (size_t)buf * t.row_bytes;
        } else {
            rrow = reinterpret_cast<const unsigned char*>(a.img[im]) + (size_t)row * t.row_bytes;  // global
        }
        if (a.dtype == BS_DTYPE_U16 && !(e0 & 1) && !((size_t)rrow & 3)) {
            // interior pairs are aligned 32-bit words of the staged row
            const unsigned int* r32 = reinterpret_cast<const unsigned int*>(rrow);
            const int n_lo = e0 >> 1, n_hi = (e0 + a.dx) >> 1;   // pairs fully inside [e0, e0 + dx)
            for (int n = lane; n < M; n += 32) {
                float2 v = make_float2(0.f, 0.f);
                if (n >= n_lo && n < n_hi) {
                    const unsigned int w = r32[n - n_lo];
                    v = make_float2((float)(w & 0xffffu) * gyz, (float)(w >> 16) * gyz);
                } else if (2 * n < a.Ex) {
                    const int xp = 2 * n;
                    const float g0 = (a.w_x[xp] * wy) * wz, g1 = (a.w_x[xp + 1] * wy) * wz;
                    const float p0 = raw_elem(rrow, BS_DTYPE_U16, a.idx_x[xp]), p1 = raw_elem(rrow, BS_DTYPE_U16, a.idx_x[xp + 1]);
                    v = make_float2(g0 != 0.f ? p0 * g0 : 0.f, g1 != 0.f ? p1 * g1 : 0.f);
                }
                A[n] = v;
            }
        } else {
            for (int n = lane; n < M; n += 32) {
                float2 v = make_float2(0.f, 0.f);
                const int xp = 2 * n;
                if (xp >= e0 && xp + 1 < e0 + a.dx) {
                    v = make_float2(raw_elem(rrow, a.dtype, xp - e0) * gyz, raw_elem(rrow, a.dtype, xp - e0 + 1) * gyz);
                } else if (xp < a.Ex) {
                    const float g0 = (a.w_x[xp] * wy) * wz, g1 = (a.w_x[xp + 1] * wy) * wz;
                    const float p0 = raw_elem(rrow, a.dtype, a.idx_x[xp]), p1 = raw_elem(rrow, a.dtype, a.idx_x[xp + 1]);
                    v = make_float2(g0 != 0.f ? p0 * g0 : 0.f, g1 != 0.f ? p1 * g1 : 0.f);
                }
                A[n] = v;
            }
        }
        __syncwarp();
        const float2* res = F::run(A, B, tw, twr, a.plan, 2, lane);
        // untangle: X[k] and X[M-k] share E, D and the twiddle (w_{M-k} = -conj(w_k))
        auto untangle = [&](int k, float2 w) {
            const float2 Zk = res[k];
            const float2 Zm = p_mul(res[k == 0 ? 0 : M - k], make_float2(1.f, -1.f));   // conj
            const float2 S = caddf(Zk, Zm), Dd = csubf(Zk, Zm);        // 2E, 2D
            const float2 wD = cmulf(w, Dd);
            __stcg(srow + k, cscale(0.5f, cadd_mi(S, wD)));            // E - i w D
            if (2 * k != M) __stcg(srow + (M - k), p_mul(cadd_pi(S, wD), make_float2(0.5f, -0.5f)));   // conj(E + i w D)
        };
        if (F::kStatic) {
#pragma unroll
            for (int i = 0; i < UIT; ++i) {
                const int k = lane + 32 * i;
                if (2 * k <= M) untangle(k, kRegU ? utw[kRegU ? i : 0] : tw[k]);
            }
        } else {
            for (int k = lane; 2 * k <= M; k += 32) untangle(k, tw[k]);
        }
        for (int k = M + 1 + lane; k < pitch; k += 32) __stcg(srow + k, make_float2(0.f, 0.f));
        __syncwarp();  // A/B and the consumed raw buffer are free again
    }
}

#define XW_MAXV 10   // float2 per lane covering a row of up to 32 * XW_MAXV spectrum entries
template <class F>
__global__ void __launch_bounds__(PCM_THREADS, 3) k_fft_x_c2r_w(const __grid_constant__ XC2RArgs a) {
    const int M = F::kStatic ? F::N : a.M;
    const int Px = 2 * M;
    const int pitch = F::kStatic ? ((F::N + 1 + 15) / 16) * 16 : a.pitch;
    constexpr int NW = PCM_THREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float2* tw = bs_sm;
    float2* A = bs_sm + Px + (size_t)wid * 2 * (M + 1);   // M + 1 entries each
    float2* B = A + (M + 1);
    for (int i = threadIdx.x; i < Px; i += blockDim.x) tw[i] = a.tw[i];
    __syncthreads();
    typename F::Tw twr;
    twr.init(tw, lane);
    constexpr int UIT = F::kStatic ? (F::N / 2 + 1 + 31) / 32 : 1;
    const long long n_lines = (long long)a.Py * a.Pz;
    const long long gw = (long long)blockIdx.x * NW + wid, gstride = (long long)gridDim.x * NW;
    float2 nxt[XW_MAXV];
    auto load_row = [&](long long line) {
        const float2* srow = a.spec + (size_t)line * pitch;
#pragma unroll
        for (int u = 0; u < XW_MAXV; ++u) {
            const int k = lane + 32 * u;
            nxt[u] = (k <= M) ? __ldcg(srow + k) : make_float2(0.f, 0.f);
        }
    };
    if (gw < n_lines) load_row(gw);
    for (long long line = gw; line < n_lines; line += gstride) {
        // spectrum row -> B (conjugated: partial inverse along y,z = conj of forward transforms of conj data)
#pragma unroll
        for (int u = 0; u < XW_MAXV; ++u) {
            const int k = lane + 32 * u;
            if (k <= M) B[k] = make_float2(nxt[u].x, -nxt[u].y);
        }
        __syncwarp();
        if (line + gstride < n_lines) load_row(line + gstride);   // prefetch into registers
        // tangle: A[k] and A[M-k] share E, D and the twiddle (E' = conj E, D' = -conj D, conj w' = -w):
        //   A[k] = conj(E + i O), A[M-k] = (E.x + O.y, E.y - O.x) with O = D conj(w_k)
        auto tangle = [&](int k, float2 wc) {
            const float2 Xk = B[k];
            const float2 Xm = p_mul(B[M - k], make_float2(1.f, -1.f));
            const float2 S = caddf(Xk, Xm), Dd = csubf(Xk, Xm);       // 2E, 2D
            const float2 O = cmulf(Dd, wc);                            // 2 O
            A[k] = p_mul(cadd_pi(S, O), make_float2(0.5f, -0.5f));     // conj(E + i O)
            if (k != 0 && 2 * k != M) A[M - k] = cscale(0.5f, cadd_mi(S, O));   // E - i O = (E.x + O.y, E.y - O.x)
        };
        if (F::kStatic) {
#pragma unroll
            for (int i = 0; i < UIT; ++i) {
                const int k = lane + 32 * i;
                if (2 * k <= M) tangle(k, p_mul(tw[k], make_float2(1.f, -1.f)));   // contiguous table read, conflict-free
            }
        } else {
            for (int k = lane; 2 * k <= M; k += 32) tangle(k, p_mul(tw[k], make_float2(1.f, -1.f)));
        }
        __syncwarp();
        const float2* res = F::run(A, B, tw, twr, a.plan, 2, lane);
        float2* row = a.spec + (size_t)line * pitch;
        for (int n = lane; n < M; n += 32) {
            const float2 r = res[n];
            __stcg(row + n, make_float2(r.x * a.scale, -r.y * a.scale));
        }
        __syncwarp();
    }
}

typedef FftWStatic<270, 2, 9, 6, 5> FftW270;     // register twiddles (c2r: 3 CTAs / SM)
typedef FftWStaticS<270, 2, 9, 6, 5> FftW270S;   // shared-memory stage tables (r2c: 4 CTAs / SM)

// ------------------------------------------------------------------------------------------
// strided passes (y and z)
struct StridedArgs {
    float2* a;
    float2* b;
    long long estride;   // float2 units between consecutive elements along the FFT axis
    long long ostride;   // float2 units between consecutive blockIdx.y
    const float2* tw;
    FftPlan plan;
    int tshift;          // log2(tile width in float2)
    int mode;            // 0: forward in place on (blockIdx.z ? b : a); 1: cross-power (a,b) -> a
    float thresh;        // normalisation threshold
};

template <int UNR>
__device__ __forceinline__ void tile_load(float2* dst, const float2* g, long long estride, int N, int tshift) {
    const int vshift = tshift - 1;           // float4 vectors per row = TW/2
    const int vmask = (1 << vshift) - 1;
    const int nvec = N << vshift;
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i0 = threadIdx.x; i0 < nvec; i0 += blockDim.x * UNR) {
        float4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < nvec) v[u] = __ldcg(reinterpret_cast<const float4*>(g + (long long)(i >> vshift) * estride) + (i & vmask));
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < nvec) d4[i] = v[u];
        }
    }
}

__device__ __forceinline__ void tile_store(float2* g, const float2* src, long long estride, int N, int tshift) {
    const int vshift = tshift - 1;
    const int vmask = (1 << vshift) - 1;
    const int nvec = N << vshift;
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll 4
    for (int i = threadIdx.x; i < nvec; i += blockDim.x)
        __stcg(reinterpret_cast<float4*>(g + (long long)(i >> vshift) * estride) + (i & vmask), s4[i]);
}

__device__ __forceinline__ float2 unit_or_zero(float2 x, float thresh) {
    const float m2 = x.x * x.x + x.y * x.y;
    if (m2 < thresh * thresh) return make_float2(0.f, 0.f);   // |x| < threshold
    const float inv = rsqrtf(m2);
    return make_float2(x.x * inv, x.y * inv);
}

template <class F>
__global__ void __launch_bounds__(PCM_THREADS, 3) k_fft_strided(const __grid_constant__ StridedArgs a) {
    const int tshift = F::kStatic ? F::LSHIFT : a.tshift;
    const int TW = 1 << tshift, N = F::kStatic ? F::N : a.plan.n;
    const int twpad = (N + 1) & ~1;
    float2* tw = bs_sm;
    float2* B0 = bs_sm + twpad;
    float2* B1 = B0 + (size_t)N * TW;
    const size_t base = (size_t)blockIdx.y * a.ostride + (size_t)blockIdx.x * TW;
    for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = a.tw[i];
    if (a.mode == 0) {
        float2* g = (blockIdx.z ? a.b : a.a) + base;
        tile_load<9>(B0, g, a.estride, N, tshift);
        __syncthreads();
        const float2* res = F::run(B0, B1, tw, a.plan, tshift, TW, 1);
        tile_store(g, res, a.estride, N, tshift);
    } else {
        float2* B2 = B1 + (size_t)N * TW;
        tile_load<9>(B0, a.a + base, a.estride, N, tshift);
        tile_load<9>(B1, a.b + base, a.estride, N, tshift);
        __syncthreads();
        float2* rA = F::run(B0, B2, tw, a.plan, tshift, TW, 1);
        float2* freeA = (rA == B0) ? B2 : B0;
        float2* rB = F::run(B1, freeA, tw, a.plan, tshift, TW, 1);
        float2* free2 = (rB == B1) ? freeA : B1;
        const int tot = N * TW;
        for (int i = threadIdx.x; i < tot; i += blockDim.x) {
            const float2 x = unit_or_zero(rA[i], a.thresh);
            const float2 y = unit_or_zero(rB[i], a.thresh);
            rA[i] = make_float2(x.x * y.x + x.y * y.y, x.x * y.y - x.y * y.x);  // conj(x) * y
        }
        __syncthreads();
        const float2* rQ = F::run(rA, free2, tw, a.plan, tshift, TW, 1);
        tile_store(a.a + base, rQ, a.estride, N, tshift);
    }
}

// Persistent, software-pipelined variant of mode 0: each CTA walks its tiles with three rotating
// shared-memory buffers; the NEXT tile is pulled in with cp.async (LDGSTS, 16 B per thread and
// request, no register staging) while the current one is transformed and stored.
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct StridedPipeArgs {
    StridedArgs s;
    int tiles_x;     // pitch / TW
    int n_other;     // blockIdx.y extent of the non-persistent kernel
    int n_tiles;     // tiles_x * n_other * n_img
};

template <class F>
__global__ void __launch_bounds__(PCM_THREADS, 2) k_fft_strided_pipe(const __grid_constant__ StridedPipeArgs p) {
    const StridedArgs& a = p.s;
    const int tshift = F::kStatic ? F::LSHIFT : a.tshift;
    const int TW = 1 << tshift, N = F::kStatic ? F::N : a.plan.n;
    const int twpad = (N + 1) & ~1;
    float2* tw = bs_sm;
    float2* B[3];
    B[0] = bs_sm + twpad;
    B[1] = B[0] + (size_t)N * TW;
    B[2] = B[1] + (size_t)N * TW;
    for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = a.tw[i];
    const int vshift = tshift - 1, vmask = (1 << vshift) - 1, nvec = N << vshift;

    auto tile_ptr = [&](int t) -> float2* {
        const int tx = t % p.tiles_x;
        const int r = t / p.tiles_x;
        const int o = r % p.n_other, im = r / p.n_other;
        return (im ? a.b : a.a) + (size_t)o * a.ostride + (size_t)tx * TW;
    };
    auto prefetch = [&](int t, float2* dst) {
        const float2* g = tile_ptr(t);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = threadIdx.x; i < nvec; i += blockDim.x)
            cp_async16(d4 + i, reinterpret_cast<const float4*>(g + (long long)(i >> vshift) * a.estride) + (i & vmask));
        cp_async_commit();
    };

    int t = blockIdx.x;
    if (t < p.n_tiles) prefetch(t, B[0]);
    for (int it = 0; t < p.n_tiles; t += gridDim.x, ++it) {
        float2* cur = B[(2 * it) % 3];
        float2* pong = B[(2 * it + 1) % 3];
        float2* nxt = B[(2 * it + 2) % 3];
        const int tn = t + gridDim.x;
        if (tn < p.n_tiles) {
            prefetch(tn, nxt);
            cp_async_wait<1>();   // everything but the newest group (the next tile) has landed
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float2* res = F::run(cur, pong, tw, a.plan, tshift, TW, 1);
        tile_store(tile_ptr(t), res, a.estride, N, tshift);
        __syncthreads();   // res / cur may be overwritten by the next iteration's prefetch and FFT
    }
}

// Persistent, software-pipelined cross-power pass (z): same three tile buffers as k_fft_strided mode 1 (so two
// CTAs still fit an SM), but the loads are cp.async groups that overlap the transforms:
//   A(t) lands -> FFT A   | B(t) still in flight
//   B(t) lands -> FFT B -> normalise, conj(A) * B
//   B's buffer is free    -> prefetch A(t+1) under the transform of the product
//   the transform's scratch buffer is free -> prefetch B(t+1) under the store and FFT A(t+1)
template <class F>
__global__ void __launch_bounds__(PCM_THREADS, 2) k_fft_xpower_pipe(const __grid_constant__ StridedPipeArgs p) {
    const StridedArgs& a = p.s;
    const int tshift = F::kStatic ? F::LSHIFT : a.tshift;
    const int TW = 1 << tshift, N = F::kStatic ? F::N : a.plan.n;
    const int twpad = (N + 1) & ~1;
    float2* tw = bs_sm;
    float2* bufA = bs_sm + twpad;
    float2* bufB = bufA + (size_t)N * TW;
    float2* pong = bufB + (size_t)N * TW;
    for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = a.tw[i];
    const int vshift = tshift - 1, vmask = (1 << vshift) - 1, nvec = N << vshift;
    auto tile_off = [&](int t) -> size_t {
        const int tx = t % p.tiles_x, o = t / p.tiles_x;
        return (size_t)o * a.ostride + (size_t)tx * TW;
    };
    auto prefetch = [&](const float2* g, float2* dst) {
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = threadIdx.x; i < nvec; i += blockDim.x)
            cp_async16(d4 + i, reinterpret_cast<const float4*>(g + (long long)(i >> vshift) * a.estride) + (i & vmask));
        cp_async_commit();
    };
    int t = blockIdx.x;
    if (t < p.n_tiles) {
        prefetch(a.a + tile_off(t), bufA);
        prefetch(a.b + tile_off(t), bufB);
    }
    for (; t < p.n_tiles; t += gridDim.x) {
        const int tn = t + gridDim.x;
        const bool has_next = tn < p.n_tiles;
        cp_async_wait<1>();            // A(t) landed (B(t) is the newest group)
        __syncthreads();
        float2* rA = F::run(bufA, pong, tw, a.plan, tshift, TW, 1);
        float2* freeA = (rA == bufA) ? pong : bufA;
        cp_async_wait<0>();            // B(t) landed
        __syncthreads();
        float2* rB = F::run(bufB, freeA, tw, a.plan, tshift, TW, 1);
        float2* free2 = (rB == bufB) ? freeA : bufB;
        const int tot = N * TW;
        for (int i = threadIdx.x; i < tot; i += blockDim.x) {
            const float2 x = unit_or_zero(rA[i], a.thresh);
            const float2 y = unit_or_zero(rB[i], a.thresh);
            rA[i] = make_float2(x.x * y.x + x.y * y.y, x.x * y.y - x.y * y.x);  // conj(x) * y
        }
        __syncthreads();
        if (has_next) prefetch(a.a + tile_off(tn), rB);          // rB's buffer is free from here on
        float2* rQ = F::run(rA, free2, tw, a.plan, tshift, TW, 1);
        float2* otherQ = (rQ == rA) ? free2 : rA;
        if (has_next) prefetch(a.b + tile_off(tn), otherQ);      // scratch of the last transform is free
        tile_store(a.a + tile_off(t), rQ, a.estride, N, tshift);
        __syncthreads();               // rQ is read out: it becomes the next iteration's scratch
        bufA = rB; bufB = otherQ; pong = rQ;
    }
}

// ------------------------------------------------------------------------------------------
// x pass, complex -> real, in place

template <class F>
__global__ void __launch_bounds__(PCM_THREADS) k_fft_x_c2r(const __grid_constant__ XC2RArgs a) {
    const int M = F::kStatic ? F::N : a.M;
    const int lshift = F::kStatic ? F::LSHIFT : a.lshift;
    const int LB = 1 << lshift, ls = LB + 1;
    const int Px = 2 * M;
    const int pitch = F::kStatic ? ((F::N + 1 + 15) / 16) * 16 : a.pitch;
    float2* tw = bs_sm;
    float2* T0 = bs_sm + Px;                 // (M+1) * ls
    float2* T1 = T0 + (size_t)(M + 1) * ls;  // M * ls
    const int zp = blockIdx.y, y0 = blockIdx.x * LB;
    const size_t rowbase = ((size_t)zp * a.Py + y0) * pitch;
    const int nlines = min(LB, a.Py - y0);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NW = PCM_THREADS / 32;
    for (int i = threadIdx.x; i < Px; i += blockDim.x) tw[i] = a.tw[i];
    for (int l = wid; l < LB; l += NW) {
        const float2* srow = a.spec + rowbase + (size_t)l * pitch;
        for (int k0 = 0; k0 <= M; k0 += 32 * XR_UNROLL) {
            float2 v[XR_UNROLL];
#pragma unroll
            for (int u = 0; u < XR_UNROLL; ++u) {
                const int k = k0 + u * 32 + lane;
                v[u] = make_float2(0.f, 0.f);
                if (l < nlines && k <= M) v[u] = __ldcg(srow + k);
            }
#pragma unroll
            for (int u = 0; u < XR_UNROLL; ++u) {
                const int k = k0 + u * 32 + lane;
                // partial inverse along y,z = conj of the forward transforms of conj data
                if (k <= M) T0[k * ls + l] = make_float2(v[u].x, -v[u].y);
            }
        }
    }
    __syncthreads();
    for (int item = threadIdx.x; item < M * LB; item += blockDim.x) {
        const int k = item >> lshift, l = item & (LB - 1);
        const float2 Xk = T0[k * ls + l];
        float2 Xm = T0[(M - k) * ls + l];
        Xm.y = -Xm.y;
        const float2 E = make_float2(0.5f * (Xk.x + Xm.x), 0.5f * (Xk.y + Xm.y));
        const float2 D = make_float2(0.5f * (Xk.x - Xm.x), 0.5f * (Xk.y - Xm.y));
        float2 w = tw[k];
        w.y = -w.y;  // e^{+2 pi i k / P}
        const float2 O = cmulf(D, w);
        // Z = E + i*O ; store conj(Z) (inverse via forward transform of the conjugate)
        T1[k * ls + l] = make_float2(E.x - O.y, -(E.y + O.x));
    }
    __syncthreads();
    const float2* res = F::run(T1, T0, tw, a.plan, lshift, ls, 2);
    for (int l = wid; l < nlines; l += NW) {
        float2* row = a.spec + rowbase + (size_t)l * pitch;
        for (int n = lane; n < M; n += 32) {
            const float2 r = res[n * ls + l];
            __stcg(row + n, make_float2(r.x * a.scale, -r.y * a.scale));
        }
    }
}

// ------------------------------------------------------------------------------------------
// peak search: periodic axis-neighbour local maxima, top-K per CTA
struct PeakEntry {
    float val;
    int pad;
    long long idx;
};

struct PeakArgs {
    const float* pcm;
    int Px, Py, Pz;
    long long rowpitch;  // floats (multiple of 4, >= Px rounded up to 4)
    int K;
    PeakEntry* out;      // gridDim.x * K
};

__device__ __forceinline__ bool peak_better(float v, long long i, float v2, long long i2) {
    return v > v2 || (v == v2 && i < i2);
}

// A warp streams one row at a time with PK_UNROLL float4 loads in flight per lane.  The warp
// keeps ONE sorted top-K list in registers (lane i holds the i-th best; K <= 32) and its K-th
// value as a warp-uniform threshold, so after a few rows almost every voxel is rejected by a
// single compare; only survivors fetch their six periodic neighbours and are inserted with a
// ballot + shuffle shift.  Per-CTA merge of the 8 warp lists, then K entries per CTA go out.
#define PK_UNROLL 5

struct WarpTopK {
    float v;        // lane i: value of the i-th best (or -inf)
    long long i;    // its linear index (or LLONG_MAX)
    float thr;      // warp-uniform: value of the K-th best (-inf until the list is full)
};

__device__ __forceinline__ void warp_topk_insert(WarpTopK& t, int K, float nv, long long ni, int lane) {
    // rank of the new entry = number of current entries that are better
    const bool mine_better = peak_better(t.v, t.i, nv, ni);
    const unsigned better = __ballot_sync(0xffffffffu, mine_better && lane < K);
    const int pos = __popc(better);
    if (pos >= K) return;  // warp-uniform
    const float upv = __shfl_up_sync(0xffffffffu, t.v, 1);
    const long long upi = __shfl_up_sync(0xffffffffu, t.i, 1);
    if (lane == pos) { t.v = nv; t.i = ni; }
    else if (lane > pos && lane < K) { t.v = upv; t.i = upi; }
    t.thr = __shfl_sync(0xffffffffu, t.v, K - 1);
}

__global__ void __launch_bounds__(PCM_THREADS) k_peaks(const __grid_constant__ PeakArgs a) {
    const int K = a.K;
    const long long nrows = (long long)a.Py * a.Pz;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NW = PCM_THREADS / 32;
    const int nvec = (a.Px + 3) >> 2;
    WarpTopK top;
    top.v = -INFINITY;
    top.i = 0x7fffffffffffffffLL;
    top.thr = -INFINITY;
    for (long long row = (long long)blockIdx.x * NW + wid; row < nrows; row += (long long)gridDim.x * NW) {
        const float* rp = a.pcm + row * a.rowpitch;
        const float4* rp4 = reinterpret_cast<const float4*>(rp);
        const int z = (int)(row / a.Py), y = (int)(row - (long long)z * a.Py);
        for (int v0 = 0; v0 < nvec; v0 += 32 * PK_UNROLL) {
            float4 q[PK_UNROLL];
#pragma unroll
            for (int u = 0; u < PK_UNROLL; ++u) {
                const int vi = v0 + u * 32 + lane;
                q[u] = vi < nvec ? __ldcs(rp4 + vi) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            }
            // phase 1 (unrolled, tiny): which of my 4 * PK_UNROLL values reach the warp threshold?
            unsigned int mask = 0u;
#pragma unroll
            for (int u = 0; u < PK_UNROLL; ++u) {
                const int vi = v0 + u * 32 + lane;
                const float c4[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c4[c] >= top.thr && c4[c] > -INFINITY && 4 * vi + c < a.Px) mask |= 1u << (4 * u + c);
            }
            // phase 2 (rolled, rare): neighbour test + insertion, one candidate per lane per round
            while (__any_sync(0xffffffffu, mask != 0u)) {
                bool cand = mask != 0u;
                int x = 0;
                float v = -INFINITY;
                if (cand) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1u;
                    x = 4 * (v0 + (b >> 2) * 32 + lane) + (b & 3);
                    v = rp[x];
                    cand = v >= top.thr;   // the threshold may have risen since phase 1
                    if (cand) cand = !(v < rp[x == 0 ? a.Px - 1 : x - 1] || v < rp[x == a.Px - 1 ? 0 : x + 1]);
                    if (cand) {
                        const float* rym = a.pcm + ((long long)z * a.Py + (y == 0 ? a.Py - 1 : y - 1)) * a.rowpitch;
                        const float* ryp = a.pcm + ((long long)z * a.Py + (y == a.Py - 1 ? 0 : y + 1)) * a.rowpitch;
                        const float* rzm = a.pcm + ((long long)(z == 0 ? a.Pz - 1 : z - 1) * a.Py + y) * a.rowpitch;
                        const float* rzp = a.pcm + ((long long)(z == a.Pz - 1 ? 0 : z + 1) * a.Py + y) * a.rowpitch;
                        cand = !(v < rym[x] || v < ryp[x] || v < rzm[x] || v < rzp[x]);
                    }
                }
                unsigned m = __ballot_sync(0xffffffffu, cand);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float nv = __shfl_sync(0xffffffffu, v, src);
                    const int nx = __shfl_sync(0xffffffffu, x, src);
                    warp_topk_insert(top, K, nv, row * a.Px + nx, lane);
                }
            }
        }
    }
    // merge the warp lists of this CTA into warp 0's list, then write K entries
    __shared__ float s_v[NW * PCM_KMAX];
    __shared__ long long s_i[NW * PCM_KMAX];
    if (lane < K) { s_v[wid * PCM_KMAX + lane] = top.v; s_i[wid * PCM_KMAX + lane] = top.i; }
    __syncthreads();
    if (wid == 0) {
        for (int w = 1; w < NW; ++w)
            for (int e = 0; e < K; ++e) {
                const float nv = s_v[w * PCM_KMAX + e];
                if (nv == -INFINITY) break;  // lists are sorted; uniform across the warp
                warp_topk_insert(top, K, nv, s_i[w * PCM_KMAX + e], lane);
            }
        if (lane < K) {
            PeakEntry e;
            e.val = top.v;
            e.pad = 0;
            e.idx = (top.v == -INFINITY) ? -1 : top.i;
            a.out[(size_t)blockIdx.x * K + lane] = e;
        }
    }
}

// compile-time specialisations: padded 512^3 overlaps -> 540^3 (x half-length 270)
typedef FftStatic<270, 4, 17, 2, PCM_THREADS, 9, 6, 5> FftX270;
typedef FftStatic<270, 3, 9, 2, PCM_THREADS, 9, 6, 5> FftX270L8;
typedef FftStatic<540, 3, 8, 1, PCM_THREADS, 9, 10, 6> FftS540;
typedef FftStatic<540, 2, 4, 1, PCM_THREADS, 9, 10, 6> FftS540T4;
typedef FftStatic<540, 3, 8, 1, PCM_THREADS, 27, 20> FftS540R2;   // two-stage variant (radix 27 x 20)

// ------------------------------------------------------------------------------------------
// Pearson sums for all candidate shifts in one launch
struct PearsonCand {
    int o1[3], o2[3], sz[3];
    int pad;
};

struct PearsonArgs {
    const void* img1;
    const void* img2;
    int dtype;
    int dx, dy, dz;
    const PearsonCand* cands;
    unsigned long long* sums_u;  // 5 per candidate: sa, sb, saa, sbb, sab
    double* sums_d;              // same for float input
};

// Row-chunk-major traversal: a warp owns PR_ROWS consecutive rows of image 1 and evaluates EVERY
// candidate on them before moving on, so all candidates stream through the volumes in lockstep and
// the second..K-th read of a row is an L1/L2 hit (DRAM traffic ~ one sweep instead of K sweeps).
#define PR_ROWS 16
#define PR_MLP 8

__device__ __forceinline__ void pr_acc(unsigned int va, unsigned int vb, unsigned int& ra, unsigned int& rb,
                                       unsigned long long& saa, unsigned long long& sbb, unsigned long long& sab) {
    ra += va;
    rb += vb;
    saa += (unsigned long long)va * va;   // one IMAD.WIDE.U32 with 64-bit accumulate each
    sbb += (unsigned long long)vb * vb;
    sab += (unsigned long long)va * vb;
}

__device__ __forceinline__ void pearson_u16(const PearsonArgs& a, int ncand, unsigned long long* s_acc) {
    const unsigned short* __restrict__ i1 = (const unsigned short*)a.img1;
    const unsigned short* __restrict__ i2 = (const unsigned short*)a.img2;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long nrows = (long long)a.dy * a.dz;
    const long long nchunks = (nrows + PR_ROWS - 1) / PR_ROWS;
    const bool even_rows = !(a.dx & 1) && !((size_t)i1 & 3) && !((size_t)i2 & 3);
    for (long long ch = (long long)blockIdx.x * nw + wid; ch < nchunks; ch += (long long)gridDim.x * nw) {
        const long long r0 = ch * PR_ROWS;
        const int z0 = (int)(r0 / a.dy), y0 = (int)(r0 - (long long)z0 * a.dy);
        for (int c = 0; c < ncand; ++c) {
            const PearsonCand cd = a.cands[c];
            unsigned long long sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
            bool any = false;
            // both x offsets even -> aligned ushort2 loads on both images
            const bool vec = even_rows && !((cd.o1[0] | cd.o2[0]) & 1);
            for (int rr = 0; rr < PR_ROWS; ++rr) {
                const long long r = r0 + rr;
                if (r >= nrows) break;
                int z = z0, y = y0 + rr;          // (z, y) of row r0 + rr without a 64-bit division per row
                while (y >= a.dy) { y -= a.dy; ++z; }
                const int yy = y - cd.o1[1], zz = z - cd.o1[2];
                if (yy < 0 || yy >= cd.sz[1] || zz < 0 || zz >= cd.sz[2]) continue;
                any = true;
                const unsigned short* p1 = i1 + (size_t)r * a.dx + cd.o1[0];
                const unsigned short* p2 = i2 + ((size_t)(zz + cd.o2[2]) * a.dy + (yy + cd.o2[1])) * a.dx + cd.o2[0];
                unsigned int ra = 0, rb = 0;
                const int n = cd.sz[0];
                if (vec) {
                    const unsigned int* q1 = reinterpret_cast<const unsigned int*>(p1);
                    const unsigned int* q2 = reinterpret_cast<const unsigned int*>(p2);
                    const int nv = n >> 1;
                    for (int x0 = lane; x0 < nv; x0 += 32 * PR_MLP) {   // PR_MLP words per image in flight per lane
                        unsigned int w1[PR_MLP], w2[PR_MLP];
#pragma unroll
                        for (int u = 0; u < PR_MLP; ++u) {
                            const int x = x0 + 32 * u;
                            w1[u] = x < nv ? __ldg(q1 + x) : 0u;
                            w2[u] = x < nv ? __ldg(q2 + x) : 0u;
                        }
#pragma unroll
                        for (int u = 0; u < PR_MLP; ++u) {
                            pr_acc(w1[u] & 0xffffu, w2[u] & 0xffffu, ra, rb, saa, sbb, sab);
                            pr_acc(w1[u] >> 16, w2[u] >> 16, ra, rb, saa, sbb, sab);
                        }
                    }
                    if ((n & 1) && lane == 0) pr_acc(__ldg(p1 + n - 1), __ldg(p2 + n - 1), ra, rb, saa, sbb, sab);
                } else if (even_rows && n >= 4) {
                    // exactly one x offset is odd: aligned words on one side, funnel-shifted pairs of
                    // aligned words on the other (element -1 and the following words stay inside the row)
                    // the aligned side is accumulated as "a", the funnel-shifted side as "b"; the
                    // a/b statistics are swapped once per row when image 1 is the shifted side
                    const bool odd1 = cd.o1[0] & 1;
                    const unsigned int* qa = reinterpret_cast<const unsigned int*>(odd1 ? p2 : p1);
                    const unsigned int* qm = reinterpret_cast<const unsigned int*>((odd1 ? p1 : p2) - 1);
                    const int nv = (n >> 1) - 1;  // last pair(s) handled below: qm[x + 1] must not leave the row
                    unsigned int ta = 0, tb = 0;
                    unsigned long long taa = 0, tbb = 0;
                    for (int x0 = lane; x0 < nv; x0 += 32 * PR_MLP) {
                        unsigned int wa[PR_MLP], m0[PR_MLP], m1[PR_MLP];
#pragma unroll
                        for (int u = 0; u < PR_MLP; ++u) {
                            const int x = x0 + 32 * u;
                            wa[u] = x < nv ? __ldg(qa + x) : 0u;
                            m0[u] = x < nv ? __ldg(qm + x) : 0u;
                            m1[u] = x < nv ? __ldg(qm + x + 1) : 0u;
                        }
#pragma unroll
                        for (int u = 0; u < PR_MLP; ++u) {
                            const unsigned int wm = __funnelshift_r(m0[u], m1[u], 16);
                            pr_acc(wa[u] & 0xffffu, wm & 0xffffu, ta, tb, taa, tbb, sab);
                            pr_acc(wa[u] >> 16, wm >> 16, ta, tb, taa, tbb, sab);
                        }
                    }
                    if (odd1) { ra += tb; rb += ta; saa += tbb; sbb += taa; }
                    else { ra += ta; rb += tb; saa += taa; sbb += tbb; }
                    for (int x = 2 * nv + lane; x < n; x += 32) pr_acc(__ldg(p1 + x), __ldg(p2 + x), ra, rb, saa, sbb, sab);
                } else {
#pragma unroll 4
                    for (int x = lane; x < n; x += 32) pr_acc(__ldg(p1 + x), __ldg(p2 + x), ra, rb, saa, sbb, sab);
                }
                sa += ra;
                sb += rb;
            }
            if (!any) continue;  // warp-uniform
            unsigned long long v[5] = {sa, sb, saa, sbb, sab};
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                unsigned long long t = v[k];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
                if (lane == 0 && t) atomicAdd(&s_acc[5 * c + k], t);
            }
        }
    }
}

template <typename T, typename ACC>
__device__ __forceinline__ void pearson_generic(const PearsonArgs& a, int ncand, ACC* s_acc) {
    const T* __restrict__ i1 = (const T*)a.img1;
    const T* __restrict__ i2 = (const T*)a.img2;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long nrows = (long long)a.dy * a.dz;
    const long long nchunks = (nrows + PR_ROWS - 1) / PR_ROWS;
    for (long long ch = (long long)blockIdx.x * nw + wid; ch < nchunks; ch += (long long)gridDim.x * nw) {
        const long long r0 = ch * PR_ROWS;
        for (int c = 0; c < ncand; ++c) {
            const PearsonCand cd = a.cands[c];
            ACC v[5] = {0, 0, 0, 0, 0};
            bool any = false;
            for (int rr = 0; rr < PR_ROWS; ++rr) {
                const long long r = r0 + rr;
                if (r >= nrows) break;
                const int z = (int)(r / a.dy), y = (int)(r - (long long)z * a.dy);
                const int yy = y - cd.o1[1], zz = z - cd.o1[2];
                if (yy < 0 || yy >= cd.sz[1] || zz < 0 || zz >= cd.sz[2]) continue;
                any = true;
                const T* p1 = i1 + (size_t)r * a.dx + cd.o1[0];
                const T* p2 = i2 + ((size_t)(zz + cd.o2[2]) * a.dy + (yy + cd.o2[1])) * a.dx + cd.o2[0];
                for (int x = lane; x < cd.sz[0]; x += 32) {
                    const ACC va = (ACC)__ldg(p1 + x), vb = (ACC)__ldg(p2 + x);
                    v[0] += va; v[1] += vb; v[2] += va * va; v[3] += vb * vb; v[4] += va * vb;
                }
            }
            if (!any) continue;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                ACC t = v[k];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
                if (lane == 0) atomicAdd(&s_acc[5 * c + k], t);
            }
        }
    }
}

// dynamic smem: 5 accumulators per candidate (8 bytes each)
__global__ void __launch_bounds__(PCM_THREADS) k_pearson(const __grid_constant__ PearsonArgs a, const int* __restrict__ ncand_ptr) {
    const int ncand = *ncand_ptr;     // written by k_pcm_select (no host round trip between peaks and Pearson)
    if (ncand <= 0) return;
    unsigned long long* s_u = reinterpret_cast<unsigned long long*>(bs_sm);
    double* s_d = reinterpret_cast<double*>(bs_sm);
    for (int i = threadIdx.x; i < 5 * ncand; i += blockDim.x) s_u[i] = 0ull;  // 0.0 has the same bit pattern
    __syncthreads();
    if (a.dtype == BS_DTYPE_U16) pearson_u16(a, ncand, s_u);
    else if (a.dtype == BS_DTYPE_U8) pearson_generic<unsigned char, unsigned long long>(a, ncand, s_u);
    else pearson_generic<float, double>(a, ncand, s_d);
    __syncthreads();
    for (int i = threadIdx.x; i < 5 * ncand; i += blockDim.x) {
        if (a.dtype == BS_DTYPE_F32) { if (s_d[i] != 0.0) atomicAdd(a.sums_d + i, s_d[i]); }
        else if (s_u[i]) atomicAdd(a.sums_u + i, s_u[i]);
    }
}

// ------------------------------------------------------------------------------------------
// Device-side glue between peak search and Pearson verification: merge the per-CTA top-K lists, expand every
// peak into its 2^3 wrap candidates (PhaseCorrelation2Util.expandPeakToPossibleShifts for equal-size crops),
// keep those with enough overlap, gather the 3x3x3 neighbourhoods for the sub-pixel fit.  One small CTA; the
// host reads ONE result block per pair, after the Pearson kernel, and can do so a pair late.
struct PcmSelect {
    int np;                       // peaks kept (<= K)
    int nslots;                   // Pearson-verified candidates
    long long idx[PCM_KMAX];      // linear PCM index of peak i
    float val[PCM_KMAX];
    float nb[27 * PCM_KMAX];      // periodic 3x3x3 neighbourhoods
};

// candidate i (0..7) of a peak at PCM location loc: shift per axis is loc or loc - P; returns true when the
// candidate overlaps by at least min_px voxels (then pc / npx describe the overlap boxes)
__host__ __device__ inline bool pcm_expand_candidate(const long long loc[3], const int P[3], const int d[3], int i,
                                                     long long min_px, long long shift[3], PearsonCand* pc, long long* npx_out) {
    bool overlap = true;
    long long npx = 1;
    for (int a = 0; a < 3; ++a) {
        long long s = loc[a];
        if (((i >> a) & 1) == 0) s = s < 0 ? s + P[a] : s - P[a];
        shift[a] = s;
        const long long n = d[a];
        if (s >= 0) {
            if (s >= n) { overlap = false; continue; }
            pc->o1[a] = (int)s; pc->o2[a] = 0; pc->sz[a] = (int)(n - s < n ? n - s : n);
        } else {
            if (s <= -n) { overlap = false; continue; }
            pc->o1[a] = 0; pc->o2[a] = (int)-s; pc->sz[a] = (int)(n + s < n ? n + s : n);
        }
        npx *= pc->sz[a];
    }
    pc->pad = 0;
    *npx_out = npx;
    return overlap && npx >= min_px;
}

struct SelectArgs {
    const PeakEntry* peaks;       // n_entries per-CTA candidates (idx < 0: empty)
    int n_entries;
    int K;
    int P[3], d[3];
    long long rowpitch;
    const float* pcm;
    long long min_px;
    int do_subpixel;
    PcmSelect* sel;
    PearsonCand* cands;           // 8 * K
    unsigned long long* sums;     // 5 * 8 * K, zeroed here
    int* ncand;
};

__global__ void __launch_bounds__(256) k_pcm_select(const __grid_constant__ SelectArgs a) {
    __shared__ float s_v[8];
    __shared__ long long s_i[8];
    __shared__ float s_pv[PCM_KMAX];
    __shared__ long long s_pi[PCM_KMAX];
    __shared__ int s_np, s_nslots;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float lastv = INFINITY;
    long long lasti = -1;
    int np = 0;
    for (int j = 0; j < a.K; ++j) {
        // best entry that comes strictly after the previous pick in (value desc, index asc) order
        float bv = -INFINITY;
        long long bi = LLONG_MAX;
        for (int e = tid; e < a.n_entries; e += blockDim.x) {
            const PeakEntry pe = a.peaks[e];
            if (pe.idx < 0) continue;
            if (!peak_better(lastv, lasti, pe.val, pe.idx)) continue;
            if (bi == LLONG_MAX || peak_better(pe.val, pe.idx, bv, bi)) { bv = pe.val; bi = pe.idx; }
        }
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_down_sync(0xffffffffu, bv, o);
            const long long oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (oi != LLONG_MAX && (bi == LLONG_MAX || peak_better(ov, oi, bv, bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_v[wid] = bv; s_i[wid] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
                if (s_i[w] != LLONG_MAX && (bi == LLONG_MAX || peak_better(s_v[w], s_i[w], bv, bi))) { bv = s_v[w]; bi = s_i[w]; }
            s_v[0] = bv; s_i[0] = bi;
        }
        __syncthreads();
        bv = s_v[0]; bi = s_i[0];
        __syncthreads();
        if (bi == LLONG_MAX) break;
        if (tid == 0) { s_pv[np] = bv; s_pi[np] = bi; }
        lastv = bv; lasti = bi;
        ++np;
    }
    __syncthreads();
    if (tid == 0) {
        int nslots = 0;
        for (int pi = 0; pi < np; ++pi) {
            const long long li = s_pi[pi];
            const long long loc[3] = {li % a.P[0], (li / a.P[0]) % a.P[1], li / ((long long)a.P[0] * a.P[1])};
            for (int i = 0; i < 8; ++i) {
                long long shift[3], npx;
                PearsonCand pc;
                pc.o1[0] = pc.o1[1] = pc.o1[2] = pc.o2[0] = pc.o2[1] = pc.o2[2] = pc.sz[0] = pc.sz[1] = pc.sz[2] = 0;
                if (pcm_expand_candidate(loc, a.P, a.d, i, a.min_px, shift, &pc, &npx)) a.cands[nslots++] = pc;
            }
        }
        a.sel->np = np;
        a.sel->nslots = nslots;
        *a.ncand = nslots;
        s_np = np; s_nslots = nslots;
    }
    __syncthreads();
    np = s_np;
    for (int i = tid; i < PCM_KMAX; i += blockDim.x) {
        a.sel->idx[i] = i < np ? s_pi[i] : -1;
        a.sel->val[i] = i < np ? s_pv[i] : 0.f;
    }
    for (int i = tid; i < 5 * s_nslots; i += blockDim.x) a.sums[i] = 0ull;
    if (a.do_subpixel)
        for (int t = tid; t < 27 * np; t += blockDim.x) {
            const int p = t / 27, o = t - p * 27;
            const long long li = s_pi[p];
            const int x = (int)(li % a.P[0]);
            const long long r = li / a.P[0];
            const int y = (int)(r % a.P[1]), z = (int)(r / a.P[1]);
            const int dz = o / 9 - 1, dy = (o / 3) % 3 - 1, dx = o % 3 - 1;
            const int xx = (x + dx + a.P[0]) % a.P[0], yy = (y + dy + a.P[1]) % a.P[1], zz = (z + dz + a.P[2]) % a.P[2];
            a.sel->nb[t] = a.pcm[((long long)zz * a.P[1] + yy) * a.rowpitch + xx];
        }
}

// ==========================================================================================
// host side
// ==========================================================================================
static const int kRadices[] = {16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2};

static void plan_search(int n, int depth, int* cur, int* best, int* best_len) {
    if (n == 1) {
        if (depth < *best_len) {
            *best_len = depth;
            memcpy(best, cur, sizeof(int) * depth);
        }
        return;
    }
    if (depth + 1 >= *best_len || depth >= BS_FFT_MAX_STAGES) return;
    for (int r : kRadices) {
        if (n % r) continue;
        if (depth > 0 && r > cur[depth - 1]) continue;  // non-increasing: canonical order
        cur[depth] = r;
        plan_search(n / r, depth + 1, cur, best, best_len);
    }
}

static bool make_plan(int n, FftPlan* p) {
    memset(p, 0, sizeof(*p));
    p->n = n;
    if (n == 1) { p->nst = 0; return true; }
    int cur[BS_FFT_MAX_STAGES], best[BS_FFT_MAX_STAGES], best_len = BS_FFT_MAX_STAGES + 1;
    plan_search(n, 0, cur, best, &best_len);
    if (best_len > BS_FFT_MAX_STAGES) return false;
    // odd radices first (conflict-free first-stage scatter), then descending
    std::stable_sort(best, best + best_len, [](int x, int y) { return (x & 1) > (y & 1); });
    p->nst = best_len;
    for (int i = 0; i < best_len; ++i) p->radix[i] = best[i];
    return true;
}

extern "C" int bs_good_fft_size(int n, int even) {
    int m = n < 2 ? 2 : n;
    for (;; ++m) {
        int k = m;
        for (int p : {2, 3, 5})
            while (k % p == 0) k /= p;
        if (k == 1 && (!even || m % 2 == 0)) return m;
    }
}

static int ext_size(int d, int ext) { return d + (d < ext ? 2 * d : 2 * ext); }

// per-axis source index + blending weight for every padded position (mirrors
// oracle/pcm_oracle.py:_axis_profile; BlendedExtendedMirroredRandomAccesible2 semantics)
static void axis_profile(int d, int ext, int P, std::vector<int>& idx, std::vector<float>& w) {
    const int e = std::min(ext, d);
    idx.assign(P, 0);
    w.assign(P, 0.f);
    const int period = std::max(2 * d - 2, 1);
    for (int p = 0; p < P; ++p) {
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
        w[p] = dist > 0 ? (float)(0.5 * (std::cos(M_PI * (double)dist / (double)e) + 1.0)) : 1.0f;
    }
}

struct PcmGeometry {
    int d[3], ext[3], P[3], E[3];
    int M, pitch;
    FftPlan plan_x, plan_y, plan_z;
    int lshift_x;   // log2 lines per CTA in the c2r kernel
    int lshift_r2c; // log2 lines per CTA in the r2c kernel (fewer lines -> more CTAs/SM to hide load latency)
    int tshift_y, tshift_z;
    size_t smem_x_r2c, smem_x_c2r, smem_y, smem_z;
    bool static_x, static_y, static_z;  // compile-time specialised kernels apply
};

struct PcmDeviceTables {
    int key[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // d[3], ext[3], P[3]
    int* idx[3] = {nullptr, nullptr, nullptr};
    float* w[3] = {nullptr, nullptr, nullptr};
    float2* tw[3] = {nullptr, nullptr, nullptr};
    int cap[3] = {0, 0, 0};
};

// one table set per context
static PcmDeviceTables* tables_of(bs_ctx* ctx) {
    if (!ctx->ws.tables) ctx->ws.tables = new PcmDeviceTables();
    return (PcmDeviceTables*)ctx->ws.tables;
}

void bs_pcm_slots_free(bs_ctx* ctx);
void bs_pcm_workspace_free(bs_ctx* ctx) {
    bs_pcm_slots_free(ctx);
    bs_pcm_workspace& ws = ctx->ws;
    if (ws.spec_a) cudaFree(ws.spec_a);
    if (ws.spec_b) cudaFree(ws.spec_b);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            if (ws.crop[i][j]) cudaFree(ws.crop[i][j]);
    if (ws.small) cudaFree(ws.small);
    if (ws.small_host) cudaFreeHost(ws.small_host);
    for (int i = 0; i < 2; ++i) {
        if (ws.crop_ready[i]) cudaEventDestroy(ws.crop_ready[i]);
        if (ws.crop_free[i]) cudaEventDestroy(ws.crop_free[i]);
    }
    if (ws.tables) {
        PcmDeviceTables* t = (PcmDeviceTables*)ws.tables;
        for (int d = 0; d < 3; ++d) {
            if (t->idx[d]) cudaFree(t->idx[d]);
            if (t->w[d]) cudaFree(t->w[d]);
            if (t->tw[d]) cudaFree(t->tw[d]);
        }
        delete t;
    }
    ws = bs_pcm_workspace();
}

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

static int pcm_geometry(bs_ctx* ctx, const long long dims[3], const int ext[3], PcmGeometry* g) {
    for (int d = 0; d < 3; ++d) {
        if (dims[d] <= 0 || dims[d] > 16384) return bs_set_error(ctx, BS_ERR_ARG, "pcm: dims[%d]=%lld out of range", d, dims[d]);
        if (ext[d] < 1 || ext[d] > 4096) return bs_set_error(ctx, BS_ERR_ARG, "pcm: extension[%d]=%d out of range", d, ext[d]);
        g->d[d] = (int)dims[d];
        g->ext[d] = ext[d];
        g->E[d] = ext_size(g->d[d], ext[d]);
        g->P[d] = bs_good_fft_size(g->E[d], d == 0);
    }
    g->M = g->P[0] / 2;
    g->pitch = ((g->M + 1 + 15) / 16) * 16;
    if (!make_plan(g->M, &g->plan_x) || !make_plan(g->P[1], &g->plan_y) || !make_plan(g->P[2], &g->plan_z))
        return bs_set_error(ctx, BS_ERR_UNSUPPORTED, "pcm: cannot plan FFT for padded size %dx%dx%d", g->P[0], g->P[1], g->P[2]);
    // lines per CTA for the x kernels: largest power of two <= 16 that fits shared memory
    int ls = env_int("BS_FFT_XLINES_LOG2", 4);
    for (;; --ls) {
        const int LB = 1 << ls;
        g->smem_x_r2c = ((size_t)g->P[0] + 2 * (size_t)g->M * (LB + 1)) * sizeof(float2);
        g->smem_x_c2r = ((size_t)g->P[0] + (size_t)(2 * g->M + 1) * (LB + 1)) * sizeof(float2);
        if (g->smem_x_c2r <= PCM_SMEM_MAX && g->smem_x_r2c <= PCM_SMEM_MAX) break;
        if (ls == 0) return bs_set_error(ctx, BS_ERR_UNSUPPORTED, "pcm: x size %d too large for shared memory", g->P[0]);
    }
    g->lshift_x = ls;
    g->lshift_r2c = std::min(ls, env_int("BS_FFT_R2C_LINES_LOG2", 3));
    g->smem_x_r2c = ((size_t)g->P[0] + 2 * (size_t)g->M * ((1 << g->lshift_r2c) + 1)) * sizeof(float2);
    auto strided = [&](int N, int nbuf, int pref, int* tshift, size_t* smem) -> bool {
        for (int ts = pref; ts >= 1; --ts) {
            const size_t b = ((size_t)((N + 1) & ~1) + (size_t)nbuf * N * (1 << ts)) * sizeof(float2);
            if (b <= PCM_SMEM_MAX) { *tshift = ts; *smem = b; return true; }
        }
        return false;
    };
    if (!strided(g->P[1], 2, env_int("BS_FFT_YTILE_LOG2", 3), &g->tshift_y, &g->smem_y) ||
        !strided(g->P[2], 3, env_int("BS_FFT_ZTILE_LOG2", 3), &g->tshift_z, &g->smem_z))
        return bs_set_error(ctx, BS_ERR_UNSUPPORTED, "pcm: y/z size %dx%d too large for shared memory", g->P[1], g->P[2]);
    const bool allow_static = env_int("BS_FFT_STATIC", 1) != 0;
    g->static_x = allow_static && g->M == FftX270::N && (g->lshift_x == FftX270::LSHIFT || g->lshift_x == FftX270L8::LSHIFT) &&
                  (g->lshift_r2c == 3 || g->lshift_r2c == 4);
    g->static_y = allow_static && g->P[1] == FftS540::N && g->tshift_y == FftS540::LSHIFT;
    g->static_z = allow_static && g->P[2] == FftS540::N && (g->tshift_z == FftS540::LSHIFT || g->tshift_z == FftS540T4::LSHIFT);
    return BS_OK;
}

static int pcm_tables(bs_ctx* ctx, const PcmGeometry& g, PcmDeviceTables** out) {
    PcmDeviceTables* t = tables_of(ctx);
    *out = t;
    int key[9] = {g.d[0], g.d[1], g.d[2], g.ext[0], g.ext[1], g.ext[2], g.P[0], g.P[1], g.P[2]};
    if (!memcmp(key, t->key, sizeof(key))) return BS_OK;
    BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int d = 0; d < 3; ++d) {
        const int P = g.P[d];
        if (t->cap[d] < P) {
            if (t->idx[d]) { cudaFree(t->idx[d]); cudaFree(t->w[d]); cudaFree(t->tw[d]); }
            BS_CUDA(ctx, cudaMalloc(&t->idx[d], sizeof(int) * P));
            BS_CUDA(ctx, cudaMalloc(&t->w[d], sizeof(float) * P));
            BS_CUDA(ctx, cudaMalloc(&t->tw[d], sizeof(float2) * P));
            t->cap[d] = P;
        }
        std::vector<int> idx;
        std::vector<float> w;
        axis_profile(g.d[d], g.ext[d], P, idx, w);
        std::vector<float2> tw(P);
        for (int k = 0; k < P; ++k) {
            const double ang = -2.0 * M_PI * (double)k / (double)P;
            tw[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
        BS_CUDA(ctx, cudaMemcpy(t->idx[d], idx.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
        BS_CUDA(ctx, cudaMemcpy(t->w[d], w.data(), sizeof(float) * P, cudaMemcpyHostToDevice));
        BS_CUDA(ctx, cudaMemcpy(t->tw[d], tw.data(), sizeof(float2) * P, cudaMemcpyHostToDevice));
    }
    memcpy(t->key, key, sizeof(key));
    return BS_OK;
}

static int pcm_workspace(bs_ctx* ctx, const PcmGeometry& g) {
    bs_pcm_workspace& ws = ctx->ws;
    const size_t need = (size_t)g.P[2] * g.P[1] * g.pitch * sizeof(float2);
    if (ws.spec_bytes < need) {
        BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ws.spec_a) cudaFree(ws.spec_a);
        if (ws.spec_b) cudaFree(ws.spec_b);
        ws.spec_a = ws.spec_b = nullptr;
        ws.spec_bytes = 0;
        BS_CUDA(ctx, cudaMalloc(&ws.spec_a, need));
        BS_CUDA(ctx, cudaMalloc(&ws.spec_b, need));
        ws.spec_bytes = need;
    }
    const size_t small_need = 4 << 20;   // PCM_SLOTS result slots (per-CTA peak lists, select block, candidates, sums)
    if (!ws.small) {
        BS_CUDA(ctx, cudaMalloc(&ws.small, small_need));
        BS_CUDA(ctx, cudaHostAlloc(&ws.small_host, small_need, cudaHostAllocDefault));
        ws.small_bytes = small_need;
    }
    return BS_OK;
}

static int set_smem(bs_ctx* ctx, const void* fn, size_t bytes) {
    BS_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PCM_SMEM_MAX));
    (void)bytes;
    return BS_OK;
}

// forward pipeline up to the real PCM in ws.spec_a (row pitch 2*pitch floats)
static int pcm_compute_pcm(bs_ctx* ctx, const void* d1, const void* d2, int dtype, const PcmGeometry& g,
                           PcmDeviceTables* t) {
    if (!ctx->pcm_attr_done) {
        int rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_r2c<FftGeneric>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_strided<FftGeneric>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_c2r<FftGeneric>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_r2c<FftX270>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_r2c<FftX270L8>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_r2c_tma<FftX270L8>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_r2c_w<FftW270S>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_r2c_w<FftWGeneric>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_c2r_w<FftW270>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_c2r_w<FftWGeneric>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_r2c_tma<FftX270>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_r2c_tma<FftGeneric>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_c2r<FftX270L8>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_strided<FftS540>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_strided<FftS540T4>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_strided<FftS540R2>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_strided_pipe<FftS540>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_strided_pipe<FftS540R2>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_strided_pipe<FftGeneric>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_xpower_pipe<FftS540R2>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_xpower_pipe<FftS540>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_xpower_pipe<FftGeneric>, 0))) return rc;
        if ((rc = set_smem(ctx, (const void*)k_fft_x_c2r<FftX270>, 0))) return rc;
        ctx->pcm_attr_done = true;
    }
    bs_pcm_workspace& ws = ctx->ws;
    float2* sa = (float2*)ws.spec_a;
    float2* sb = (float2*)ws.spec_b;
    {
        XR2CArgs a;
        a.img[0] = d1; a.img[1] = d2;
        a.spec[0] = sa; a.spec[1] = sb;
        a.dtype = dtype;
        a.dx = g.d[0]; a.dy = g.d[1]; a.dz = g.d[2];
        a.Px = g.P[0]; a.Py = g.P[1]; a.Pz = g.P[2]; a.M = g.M; a.pitch = g.pitch;
        a.ex = std::min(g.ext[0], g.d[0]);
        a.Ex = g.E[0]; a.Ey = g.E[1]; a.Ez = g.E[2];
        a.idx_x = t->idx[0]; a.w_x = t->w[0];
        a.idx_y = t->idx[1]; a.w_y = t->w[1];
        a.idx_z = t->idx[2]; a.w_z = t->w[2];
        a.tw = t->tw[0];
        a.plan = g.plan_x;
        a.lshift = g.lshift_r2c;
        const int LB = 1 << g.lshift_r2c;
        dim3 grid((g.P[1] + LB - 1) / LB, g.P[2], 2);
        bs_launch_scope sc(ctx, "fft_x_r2c");
        const int esize = dtype == BS_DTYPE_U16 ? 2 : dtype == BS_DTYPE_F32 ? 4 : 1;
        const int row_bytes = g.d[0] * esize;
        const size_t smem_tma = g.smem_x_r2c + 2 * (size_t)LB * row_bytes + 16;
        const bool tma_ok = env_int("BS_FFT_R2C_TMA", 1) != 0 && (row_bytes % 16) == 0 && ((size_t)d1 % 16) == 0 &&
                            ((size_t)d2 % 16) == 0 && smem_tma <= PCM_SMEM_MAX && (g.smem_x_r2c % 16) == 0;
        const int xmode = env_int("BS_FFT_X_WARP", 1);
        const size_t smem_w = ((size_t)g.P[0] + (size_t)(PCM_THREADS / 32) * 2 * g.M) * sizeof(float2) +
                              (tma_ok ? (size_t)(PCM_THREADS / 32) * (2 * (size_t)row_bytes + 16) : 0);
        if (xmode && g.M <= 32 * XW_MAXV - 1 && smem_w <= PCM_SMEM_MAX && (((size_t)g.P[0] + 16 * (size_t)g.M) * 8) % 16 == 0) {
            XWArgs t;
            t.x = a;
            t.row_bytes = row_bytes;
            t.use_tma = tma_ok ? 1 : 0;
            t.n_lines = 2LL * g.P[2] * g.P[1];
            const int per_sm = std::max(1, std::min(6, (int)(PCM_SMEM_MAX / (smem_w + 1024))));
            const int nctas = (int)std::min<long long>((t.n_lines + 7) / 8, (long long)ctx->sm_count * per_sm);
            if (g.M == FftW270S::N && env_int("BS_FFT_STATIC", 1)) k_fft_x_r2c_w<FftW270S><<<nctas, PCM_THREADS, smem_w, ctx->stream>>>(t);
            else k_fft_x_r2c_w<FftWGeneric><<<nctas, PCM_THREADS, smem_w, ctx->stream>>>(t);
        } else if (tma_ok) {
            XR2CTmaArgs t;
            t.x = a;
            t.n_groups = (g.P[1] + LB - 1) / LB;
            t.n_items = 2 * g.P[2] * t.n_groups;
            t.row_bytes = row_bytes;
            t.esize = esize;
            const int per_sm = std::max(1, (int)(PCM_SMEM_MAX / (smem_tma + 1024)));
            const int nctas = std::min(t.n_items, ctx->sm_count * std::min(per_sm, 4));
            if (g.static_x && g.lshift_r2c == 3) k_fft_x_r2c_tma<FftX270L8><<<nctas, PCM_THREADS, smem_tma, ctx->stream>>>(t);
            else if (g.static_x && g.lshift_r2c == 4) k_fft_x_r2c_tma<FftX270><<<nctas, PCM_THREADS, smem_tma, ctx->stream>>>(t);
            else k_fft_x_r2c_tma<FftGeneric><<<nctas, PCM_THREADS, smem_tma, ctx->stream>>>(t);
        } else if (g.static_x && g.lshift_r2c == 3) k_fft_x_r2c<FftX270L8><<<grid, PCM_THREADS, g.smem_x_r2c, ctx->stream>>>(a);
        else if (g.static_x && g.lshift_r2c == 4) k_fft_x_r2c<FftX270><<<grid, PCM_THREADS, g.smem_x_r2c, ctx->stream>>>(a);
        else k_fft_x_r2c<FftGeneric><<<grid, PCM_THREADS, g.smem_x_r2c, ctx->stream>>>(a);
    }
    BS_CUDA(ctx, cudaGetLastError());
    {
        StridedArgs a;
        a.a = sa; a.b = sb;
        a.estride = g.pitch;
        a.ostride = (long long)g.P[1] * g.pitch;
        a.tw = t->tw[1];
        a.plan = g.plan_y;
        a.tshift = g.tshift_y;
        a.mode = 0;
        a.thresh = 0.f;
        dim3 grid(g.pitch >> g.tshift_y, g.P[2], 2);
        bs_launch_scope sc(ctx, "fft_y");
        const size_t smem_pipe = ((size_t)((g.P[1] + 1) & ~1) + 3 * (size_t)g.P[1] * (1 << g.tshift_y)) * sizeof(float2);
        if (env_int("BS_FFT_Y_PIPE", 1) && smem_pipe <= PCM_SMEM_MAX) {
            StridedPipeArgs pp;
            pp.s = a;
            pp.tiles_x = g.pitch >> g.tshift_y;
            pp.n_other = g.P[2];
            pp.n_tiles = pp.tiles_x * pp.n_other * 2;
            const int per_sm = std::max(1, std::min(2, (int)(PCM_SMEM_MAX / (smem_pipe + 1024))));
            const int nctas = std::min(pp.n_tiles, ctx->sm_count * per_sm);
            if (g.static_y && env_int("BS_FFT_Y_R2", 1)) k_fft_strided_pipe<FftS540R2><<<nctas, PCM_THREADS, smem_pipe, ctx->stream>>>(pp);
            else if (g.static_y) k_fft_strided_pipe<FftS540><<<nctas, PCM_THREADS, smem_pipe, ctx->stream>>>(pp);
            else k_fft_strided_pipe<FftGeneric><<<nctas, PCM_THREADS, smem_pipe, ctx->stream>>>(pp);
        } else if (g.static_y) k_fft_strided<FftS540><<<grid, PCM_THREADS, g.smem_y, ctx->stream>>>(a);
        else k_fft_strided<FftGeneric><<<grid, PCM_THREADS, g.smem_y, ctx->stream>>>(a);
    }
    BS_CUDA(ctx, cudaGetLastError());
    {
        StridedArgs a;
        a.a = sa; a.b = sb;
        a.estride = (long long)g.P[1] * g.pitch;
        a.ostride = g.pitch;
        a.tw = t->tw[2];
        a.plan = g.plan_z;
        a.tshift = g.tshift_z;
        a.mode = 1;
        a.thresh = 1e-5f;  // PhaseCorrelation2Util.normalizeInterval threshold
        dim3 grid(g.pitch >> g.tshift_z, g.P[1], 1);
        bs_launch_scope sc(ctx, "fft_z_xpower");
        if (env_int("BS_FFT_Z_PIPE", 1) && g.tshift_z >= 1 && !(g.static_z && g.tshift_z == 2)) {
            StridedPipeArgs pp;
            pp.s = a;
            pp.tiles_x = g.pitch >> g.tshift_z;
            pp.n_other = g.P[1];
            pp.n_tiles = pp.tiles_x * pp.n_other;
            const int per_sm = std::max(1, std::min(2, (int)(PCM_SMEM_MAX / (g.smem_z + 1024))));
            const int nctas = std::min(pp.n_tiles, ctx->sm_count * per_sm);
            if (g.static_z && g.tshift_z == 3 && env_int("BS_FFT_Z_R2", 1)) k_fft_xpower_pipe<FftS540R2><<<nctas, PCM_THREADS, g.smem_z, ctx->stream>>>(pp);
            else if (g.static_z && g.tshift_z == 3) k_fft_xpower_pipe<FftS540><<<nctas, PCM_THREADS, g.smem_z, ctx->stream>>>(pp);
            else k_fft_xpower_pipe<FftGeneric><<<nctas, PCM_THREADS, g.smem_z, ctx->stream>>>(pp);
        } else if (g.static_z && g.tshift_z == 3 && env_int("BS_FFT_Z_R2", 1)) k_fft_strided<FftS540R2><<<grid, PCM_THREADS, g.smem_z, ctx->stream>>>(a);
        else if (g.static_z && g.tshift_z == 2) k_fft_strided<FftS540T4><<<grid, PCM_THREADS, g.smem_z, ctx->stream>>>(a);
        else if (g.static_z) k_fft_strided<FftS540><<<grid, PCM_THREADS, g.smem_z, ctx->stream>>>(a);
        else k_fft_strided<FftGeneric><<<grid, PCM_THREADS, g.smem_z, ctx->stream>>>(a);
    }
    BS_CUDA(ctx, cudaGetLastError());
    {
        StridedArgs a;
        a.a = sa; a.b = sb;
        a.estride = g.pitch;
        a.ostride = (long long)g.P[1] * g.pitch;
        a.tw = t->tw[1];
        a.plan = g.plan_y;
        a.tshift = g.tshift_y;
        a.mode = 0;
        a.thresh = 0.f;
        dim3 grid(g.pitch >> g.tshift_y, g.P[2], 1);
        bs_launch_scope sc(ctx, "fft_y_inv");
        const size_t smem_pipe = ((size_t)((g.P[1] + 1) & ~1) + 3 * (size_t)g.P[1] * (1 << g.tshift_y)) * sizeof(float2);
        if (env_int("BS_FFT_Y_PIPE", 1) && smem_pipe <= PCM_SMEM_MAX) {
            StridedPipeArgs pp;
            pp.s = a;
            pp.tiles_x = g.pitch >> g.tshift_y;
            pp.n_other = g.P[2];
            pp.n_tiles = pp.tiles_x * pp.n_other * 1;
            const int per_sm = std::max(1, std::min(2, (int)(PCM_SMEM_MAX / (smem_pipe + 1024))));
            const int nctas = std::min(pp.n_tiles, ctx->sm_count * per_sm);
            if (g.static_y && env_int("BS_FFT_Y_R2", 1)) k_fft_strided_pipe<FftS540R2><<<nctas, PCM_THREADS, smem_pipe, ctx->stream>>>(pp);
            else if (g.static_y) k_fft_strided_pipe<FftS540><<<nctas, PCM_THREADS, smem_pipe, ctx->stream>>>(pp);
            else k_fft_strided_pipe<FftGeneric><<<nctas, PCM_THREADS, smem_pipe, ctx->stream>>>(pp);
        } else if (g.static_y) k_fft_strided<FftS540><<<grid, PCM_THREADS, g.smem_y, ctx->stream>>>(a);
        else k_fft_strided<FftGeneric><<<grid, PCM_THREADS, g.smem_y, ctx->stream>>>(a);
    }
    BS_CUDA(ctx, cudaGetLastError());
    {
        XC2RArgs a;
        a.spec = sa;
        a.Px = g.P[0]; a.Py = g.P[1]; a.Pz = g.P[2]; a.M = g.M; a.pitch = g.pitch;
        a.tw = t->tw[0];
        a.plan = g.plan_x;
        a.lshift = g.lshift_x;
        a.scale = (float)(1.0 / ((double)g.M * (double)g.P[1] * (double)g.P[2]));
        const int LB = 1 << g.lshift_x;
        dim3 grid((g.P[1] + LB - 1) / LB, g.P[2], 1);
        bs_launch_scope sc(ctx, "fft_x_c2r");
        const size_t smem_w = ((size_t)g.P[0] + (size_t)(PCM_THREADS / 32) * 2 * (g.M + 1)) * sizeof(float2);
        if (env_int("BS_FFT_X_WARP", 1) && g.M <= 32 * XW_MAXV - 1 && smem_w <= PCM_SMEM_MAX) {
            const long long n_lines = (long long)g.P[1] * g.P[2];
            const int per_sm = std::max(1, std::min(6, (int)(PCM_SMEM_MAX / (smem_w + 1024))));
            const int nctas = (int)std::min<long long>((n_lines + 7) / 8, (long long)ctx->sm_count * per_sm);
            if (g.M == FftW270::N && env_int("BS_FFT_STATIC", 1)) k_fft_x_c2r_w<FftW270><<<nctas, PCM_THREADS, smem_w, ctx->stream>>>(a);
            else k_fft_x_c2r_w<FftWGeneric><<<nctas, PCM_THREADS, smem_w, ctx->stream>>>(a);
        } else if (g.static_x && g.lshift_x == 3) k_fft_x_c2r<FftX270L8><<<grid, PCM_THREADS, g.smem_x_c2r, ctx->stream>>>(a);
        else if (g.static_x) k_fft_x_c2r<FftX270><<<grid, PCM_THREADS, g.smem_x_c2r, ctx->stream>>>(a);
        else k_fft_x_c2r<FftGeneric><<<grid, PCM_THREADS, g.smem_x_c2r, ctx->stream>>>(a);
    }
    BS_CUDA(ctx, cudaGetLastError());
    return BS_OK;
}

static void solve3(const double H[3][3], const double rhs[3], double out[3]) {
    const double a = H[0][0], b = H[0][1], c = H[0][2], d = H[1][0], e = H[1][1], f = H[1][2], g = H[2][0],
                 h = H[2][1], i = H[2][2];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    out[0] = out[1] = out[2] = 0.0;
    if (det == 0.0 || !std::isfinite(det)) return;
    const double r0 = rhs[0], r1 = rhs[1], r2 = rhs[2];
    const double x = (r0 * (e * i - f * h) - b * (r1 * i - f * r2) + c * (r1 * h - e * r2)) / det;
    const double y = (a * (r1 * i - f * r2) - r0 * (d * i - f * g) + c * (d * r2 - r1 * g)) / det;
    const double z = (a * (e * r2 - r1 * h) - b * (d * r2 - r1 * g) + r0 * (d * h - e * g)) / det;
    if (std::isfinite(x) && std::isfinite(y) && std::isfinite(z)) { out[0] = x; out[1] = y; out[2] = z; }
}

// quadratic sub-pixel fit on a 3x3x3 neighbourhood nb[dz+1][dy+1][dx+1]
// (imglib2 SubpixelLocalization as used by PhaseCorrelationPeak2.calculateSubpixelLocalization)
static void subpixel_offset(const float* nb, double out[3]) {
    auto f = [&](int z, int y, int x) { return (double)nb[(z * 3 + y) * 3 + x]; };
    const double c = f(1, 1, 1);
    const double g[3] = {(f(1, 1, 2) - f(1, 1, 0)) / 2.0, (f(1, 2, 1) - f(1, 0, 1)) / 2.0, (f(2, 1, 1) - f(0, 1, 1)) / 2.0};
    double H[3][3];
    H[0][0] = f(1, 1, 2) - 2 * c + f(1, 1, 0);
    H[1][1] = f(1, 2, 1) - 2 * c + f(1, 0, 1);
    H[2][2] = f(2, 1, 1) - 2 * c + f(0, 1, 1);
    H[0][1] = H[1][0] = (f(1, 2, 2) - f(1, 2, 0) - f(1, 0, 2) + f(1, 0, 0)) / 4.0;
    H[0][2] = H[2][0] = (f(2, 1, 2) - f(2, 1, 0) - f(0, 1, 2) + f(0, 1, 0)) / 4.0;
    H[1][2] = H[2][1] = (f(2, 2, 1) - f(2, 0, 1) - f(0, 2, 1) + f(0, 0, 1)) / 4.0;
    const double rhs[3] = {-g[0], -g[1], -g[2]};
    solve3(H, rhs, out);
}

struct HostCand {
    long long shift[3];
    int peak;      // index into the peak list
    int order;     // upstream enumeration order
    long long npx;
    double r;
    int slot;      // slot in the device candidate list, -1 when below min overlap
};

static double pearson_from_int_sums(const unsigned long long s[5], long long n) {
    // n*Sxy - Sx*Sy etc. in exact 128-bit integer arithmetic, final ratio in double
    const __int128 N = n;
    const __int128 sa = s[0], sb = s[1], saa = s[2], sbb = s[3], sab = s[4];
    const __int128 cov = N * sab - sa * sb;
    const __int128 va = N * saa - sa * sa;
    const __int128 vb = N * sbb - sb * sb;
    if (va == 0 || vb == 0) return 0.0;  // getCorrelation: constant overlap -> 0
    return (double)cov / (std::sqrt((double)va) * std::sqrt((double)vb));
}

static double pearson_from_dbl_sums(const double s[5], long long n) {
    const double N = (double)n;
    const double cov = s[4] - s[0] * s[1] / N;
    const double va = s[2] - s[0] * s[0] / N;
    const double vb = s[3] - s[1] * s[1] / N;
    if (!(va > 0.0) || !(vb > 0.0)) return 0.0;
    return cov / std::sqrt(va * vb);
}

// full pipeline on device-resident crops
// Result slots: a pair's device-side result block (peaks, neighbourhoods, candidates, Pearson sums) is read
// back by ONE async copy; the host math of pair i (r from integer sums, candidate sort, sub-pixel solve) runs
// after pair i+1 has been enqueued, so the stream never drains between pairs.
#define PCM_SLOTS 2
struct PcmPending {
    bool active = false;
    PcmGeometry g;
    bs_pcm_params p;
    int dtype = 0;
    long long min_px = 0;
    size_t off_sel = 0, off_cands = 0, off_sums = 0, slot_base = 0;
};

struct PcmSlotState {
    PcmPending pend[PCM_SLOTS];
    cudaEvent_t done[PCM_SLOTS] = {};
};
static std::mutex g_slot_mu;
static std::unordered_map<bs_ctx*, PcmSlotState*> g_slot_states;
static PcmSlotState* slots_of(bs_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_slot_mu);
    auto it = g_slot_states.find(ctx);
    if (it != g_slot_states.end()) return it->second;
    PcmSlotState* s = new PcmSlotState();
    g_slot_states[ctx] = s;
    return s;
}
void bs_pcm_slots_free(bs_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_slot_mu);
    auto it = g_slot_states.find(ctx);
    if (it == g_slot_states.end()) return;
    for (int i = 0; i < PCM_SLOTS; ++i)
        if (it->second->done[i]) cudaEventDestroy(it->second->done[i]);
    delete it->second;
    g_slot_states.erase(it);
}

static int pcm_check_params(bs_ctx* ctx, const bs_pcm_params* p, int dtype) {
    if (p->peaks_to_check < 1 || p->peaks_to_check > PCM_KMAX)
        return bs_set_error(ctx, BS_ERR_ARG, "pcm: peaks_to_check must be in [1,%d]", PCM_KMAX);
    if (p->interpolate_xcorr)
        return bs_set_error(ctx, BS_ERR_UNSUPPORTED, "pcm: interpolate_xcorr is not supported (reference default false)");
    if (dtype != BS_DTYPE_U16 && dtype != BS_DTYPE_F32 && dtype != BS_DTYPE_U8)
        return bs_set_error(ctx, BS_ERR_ARG, "pcm: bad dtype %d", dtype);
    return BS_OK;
}

// everything of one pair that runs on the device, asynchronously, into result slot `slot`
static int pcm_enqueue(bs_ctx* ctx, const void* d1, const void* d2, const long long dims[3], int dtype,
                       const bs_pcm_params* p, int slot) {
    int rc = pcm_check_params(ctx, p, dtype);
    if (rc) return rc;
    PcmSlotState* S = slots_of(ctx);
    PcmPending& pd = S->pend[slot];
    if ((rc = pcm_geometry(ctx, dims, p->extension, &pd.g))) return rc;
    const PcmGeometry& g = pd.g;
    PcmDeviceTables* t;
    if ((rc = pcm_tables(ctx, g, &t))) return rc;
    if ((rc = pcm_workspace(ctx, g))) return rc;
    if ((rc = pcm_compute_pcm(ctx, d1, d2, dtype, g, t))) return rc;

    bs_pcm_workspace& ws = ctx->ws;
    const int K = p->peaks_to_check;
    // one resident wave of persistent CTAs (the kernel is latency bound: a second, partial wave halves the bytes in flight
    // for the tail of the launch)
    static int peak_occ = 0;
    if (!peak_occ) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&peak_occ, k_peaks, PCM_THREADS, 0) != cudaSuccess || peak_occ < 1) peak_occ = 4;
        peak_occ = std::min(peak_occ, env_int("BS_PEAKS_CTAS_PER_SM", 8));
    }
    const int peak_ctas = std::max(1, std::min(ctx->sm_count * peak_occ, (g.P[1] * g.P[2] + 7) / 8));
    // small-buffer layout of one slot (device and pinned mirror share offsets)
    const size_t slot_bytes = ws.small_bytes / PCM_SLOTS;
    pd.slot_base = (size_t)slot * slot_bytes;
    size_t off = 0;
    pd.off_sel = off;   off += (sizeof(PcmSelect) + 15) & ~(size_t)15;
    pd.off_cands = off; off += sizeof(PearsonCand) * 8 * PCM_KMAX;
    pd.off_sums = off;  off += sizeof(unsigned long long) * 5 * 8 * PCM_KMAX;
    const size_t readback = off;
    const size_t off_ncand = off; off += 16;
    const size_t off_peaks = off; off += sizeof(PeakEntry) * (size_t)peak_ctas * K;
    if (off > slot_bytes) return bs_set_error(ctx, BS_ERR_NOMEM, "pcm: scratch too small");
    unsigned char* dsmall = (unsigned char*)ws.small + pd.slot_base;
    unsigned char* hsmall = (unsigned char*)ws.small_host + pd.slot_base;
    {
        PeakArgs a;
        a.pcm = (const float*)ws.spec_a;
        a.Px = g.P[0]; a.Py = g.P[1]; a.Pz = g.P[2];
        a.rowpitch = 2LL * g.pitch;
        a.K = K;
        a.out = (PeakEntry*)(dsmall + off_peaks);
        bs_launch_scope sc(ctx, "peaks");
        k_peaks<<<peak_ctas, PCM_THREADS, 0, ctx->stream>>>(a);
    }
    BS_CUDA(ctx, cudaGetLastError());
    const long long n_px = (long long)g.d[0] * g.d[1] * g.d[2];
    pd.min_px = (long long)((double)n_px * p->min_overlap_frac);
    pd.p = *p;
    pd.dtype = dtype;
    {
        SelectArgs a;
        a.peaks = (const PeakEntry*)(dsmall + off_peaks);
        a.n_entries = peak_ctas * K;
        a.K = K;
        for (int d = 0; d < 3; ++d) { a.P[d] = g.P[d]; a.d[d] = g.d[d]; }
        a.rowpitch = 2LL * g.pitch;
        a.pcm = (const float*)ws.spec_a;
        a.min_px = pd.min_px;
        a.do_subpixel = p->do_subpixel ? 1 : 0;
        a.sel = (PcmSelect*)(dsmall + pd.off_sel);
        a.cands = (PearsonCand*)(dsmall + pd.off_cands);
        a.sums = (unsigned long long*)(dsmall + pd.off_sums);
        a.ncand = (int*)(dsmall + off_ncand);
        bs_launch_scope sc(ctx, "select");
        k_pcm_select<<<1, 256, 0, ctx->stream>>>(a);
    }
    BS_CUDA(ctx, cudaGetLastError());
    {
        PearsonArgs a;
        a.img1 = d1; a.img2 = d2;
        a.dtype = dtype;
        a.dx = g.d[0]; a.dy = g.d[1]; a.dz = g.d[2];
        a.cands = (const PearsonCand*)(dsmall + pd.off_cands);
        a.sums_u = (unsigned long long*)(dsmall + pd.off_sums);
        a.sums_d = (double*)(dsmall + pd.off_sums);
        const long long rows = (long long)g.d[1] * g.d[2];
        const long long chunks = (rows + PR_ROWS - 1) / PR_ROWS;
        const int ctas = (int)std::max<long long>(1, std::min<long long>((chunks + 7) / 8, (long long)ctx->sm_count * 8));
        bs_launch_scope sc(ctx, "pearson");
        k_pearson<<<ctas, PCM_THREADS, sizeof(unsigned long long) * 5 * 8 * K, ctx->stream>>>(a, (const int*)(dsmall + off_ncand));
    }
    BS_CUDA(ctx, cudaGetLastError());
    BS_CUDA(ctx, cudaMemcpyAsync(hsmall, dsmall, readback, cudaMemcpyDeviceToHost, ctx->stream));
    if (!S->done[slot]) BS_CUDA(ctx, cudaEventCreateWithFlags(&S->done[slot], cudaEventDisableTiming));
    BS_CUDA(ctx, cudaEventRecord(S->done[slot], ctx->stream));
    pd.active = true;
    return BS_OK;
}

// host half of a pair: wait for its result block, derive r, sort the candidates, solve the sub-pixel fit
static int pcm_finish(bs_ctx* ctx, int slot, bs_pcm_result* out) {
    PcmSlotState* S = slots_of(ctx);
    PcmPending& pd = S->pend[slot];
    memset(out, 0, sizeof(*out));
    out->r = -INFINITY;
    if (!pd.active) return bs_set_error(ctx, BS_ERR_ARG, "pcm: no pending pair in slot %d", slot);
    pd.active = false;
    BS_CUDA(ctx, cudaEventSynchronize(S->done[slot]));
    const PcmGeometry& g = pd.g;
    for (int d = 0; d < 3; ++d) out->pad[d] = g.P[d];
    const unsigned char* hsmall = (const unsigned char*)ctx->ws.small_host + pd.slot_base;
    const PcmSelect* sel = (const PcmSelect*)(hsmall + pd.off_sel);
    const unsigned long long* hsums = (const unsigned long long*)(hsmall + pd.off_sums);
    const int np = sel->np;
    if (np == 0) return BS_OK;  // found = 0
    // same enumeration as k_pcm_select: slots are numbered in (peak, candidate) order
    std::vector<HostCand> cands;
    int nslots = 0;
    for (int pi = 0; pi < np; ++pi) {
        const long long li = sel->idx[pi];
        const long long loc[3] = {li % g.P[0], (li / g.P[0]) % g.P[1], li / ((long long)g.P[0] * g.P[1])};
        for (int i = 0; i < 8; ++i) {
            HostCand c;
            c.peak = pi;
            c.order = pi * 8 + i;
            c.r = -INFINITY;
            c.npx = 0;
            c.slot = -1;
            PearsonCand pc;
            memset(&pc, 0, sizeof(pc));
            long long npx = 0;
            if (pcm_expand_candidate(loc, g.P, g.d, i, pd.min_px, c.shift, &pc, &npx)) {
                out->pearson_px += npx;
                c.npx = npx;
                c.slot = nslots++;
            }
            cands.push_back(c);
        }
    }
    if (nslots != sel->nslots)
        return bs_set_error(ctx, BS_ERR_CUDA, "pcm: candidate enumeration mismatch (host %d, device %d)", nslots, sel->nslots);
    out->n_candidates = nslots;
    for (auto& c : cands) {
        if (c.slot < 0) continue;
        if (pd.dtype == BS_DTYPE_F32) c.r = pearson_from_dbl_sums((const double*)hsums + 5 * c.slot, c.npx);
        else c.r = pearson_from_int_sums(hsums + 5 * c.slot, c.npx);
    }
    // Collections.sort(peaks, reverseOrder(by crossCorr, then nPixel)) -- stable
    std::stable_sort(cands.begin(), cands.end(), [](const HostCand& x, const HostCand& y) {
        if (x.r != y.r) return x.r > y.r;
        return x.npx > y.npx;
    });
    const HostCand& best = cands[0];
    if (std::isinf(best.r)) return BS_OK;  // found = 0
    out->found = 1;
    out->r = best.r;
    out->n_overlap_px = best.npx;
    const long long li = sel->idx[best.peak];
    out->peak_index[0] = li % g.P[0];
    out->peak_index[1] = (li / g.P[0]) % g.P[1];
    out->peak_index[2] = li / ((long long)g.P[0] * g.P[1]);
    out->pcm_value = sel->val[best.peak];
    double sub[3] = {0, 0, 0};
    if (pd.p.do_subpixel) subpixel_offset(sel->nb + 27 * best.peak, sub);
    for (int d = 0; d < 3; ++d) {
        out->shift_int[d] = best.shift[d];
        out->shift_sub[d] = (double)best.shift[d] + sub[d];
    }
    return BS_OK;
}

static size_t crop_bytes_of(const long long dims[3], int dtype) {
    const size_t es = dtype == BS_DTYPE_U16 ? 2 : dtype == BS_DTYPE_F32 ? 4 : 1;
    return (size_t)dims[0] * dims[1] * dims[2] * es;
}

static int ensure_crop_buffers(bs_ctx* ctx, size_t bytes, int nbuf) {
    bs_pcm_workspace& ws = ctx->ws;
    if (ws.crop_bytes >= bytes && ws.crop[0][0] && (nbuf < 2 || ws.crop[1][0])) return BS_OK;
    BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    BS_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    const size_t nb = std::max(bytes, ws.crop_bytes);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            if (ws.crop[i][j]) cudaFree(ws.crop[i][j]);
            ws.crop[i][j] = nullptr;
        }
    ws.crop_bytes = 0;
    for (int i = 0; i < nbuf; ++i)
        for (int j = 0; j < 2; ++j) BS_CUDA(ctx, cudaMalloc(&ws.crop[i][j], nb));
    for (int i = 0; i < 2; ++i) {
        if (!ws.crop_ready[i]) BS_CUDA(ctx, cudaEventCreateWithFlags(&ws.crop_ready[i], cudaEventDisableTiming));
        if (!ws.crop_free[i]) BS_CUDA(ctx, cudaEventCreateWithFlags(&ws.crop_free[i], cudaEventDisableTiming));
    }
    ws.crop_bytes = nb;
    return BS_OK;
}

extern "C" {

void bs_pcm_default_params(bs_pcm_params* p) {
    if (!p) return;
    p->peaks_to_check = 5;
    p->do_subpixel = 1;
    p->interpolate_xcorr = 0;
    p->min_overlap_frac = 0.25;
    p->extension[0] = p->extension[1] = p->extension[2] = 10;
}

int bs_pcm_pair(bs_ctx* ctx, const void* img1, const void* img2, const long long dims[3], int dtype,
                const bs_pcm_params* params, int on_device, bs_pcm_result* out) {
    if (!ctx) return BS_ERR_ARG;
    const void* a1[1] = {img1};
    const void* a2[1] = {img2};
    return bs_pcm_batch(ctx, 1, a1, a2, dims, dtype, params, on_device, out);
}

int bs_pcm_batch(bs_ctx* ctx, int n, const void* const* img1, const void* const* img2, const long long* dims,
                 int dtype, const bs_pcm_params* params, int on_device, bs_pcm_result* out) {
    if (!ctx) return BS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (n < 0 || (n > 0 && (!img1 || !img2 || !dims || !params || !out)))
        return bs_set_error(ctx, BS_ERR_ARG, "bs_pcm_batch: NULL argument");
    if (n == 0) return BS_OK;
    BS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n > 0 && (params->peaks_to_check < 1 || params->peaks_to_check > PCM_KMAX))
        return bs_set_error(ctx, BS_ERR_ARG, "pcm: peaks_to_check must be in [1,%d]", PCM_KMAX);
    for (int i = 0; i < n; ++i) {
        if (!img1[i] || !img2[i]) return bs_set_error(ctx, BS_ERR_ARG, "bs_pcm_batch: pair %d has a NULL image", i);
        for (int d = 0; d < 3; ++d)
            if (dims[3 * i + d] <= 0) return bs_set_error(ctx, BS_ERR_ARG, "bs_pcm_batch: pair %d has dims[%d] <= 0", i, d);
        // validate every pair's geometry BEFORE any byte is copied or any kernel is launched
        PcmGeometry g;
        const int rc = pcm_geometry(ctx, dims + 3 * i, params->extension, &g);
        if (rc) return rc;
    }
    // pair i's host half (result read-back, candidate sort, sub-pixel solve) runs after pair i+1 was enqueued.
    // A change of geometry rebuilds tables / workspace, so the previous pair is finished first.
    auto same_dims = [&](int i, int j) {
        return dims[3 * i] == dims[3 * j] && dims[3 * i + 1] == dims[3 * j + 1] && dims[3 * i + 2] == dims[3 * j + 2];
    };
    if (on_device) {
        int pending = -1;
        for (int i = 0; i < n; ++i) {
            if (pending >= 0 && !same_dims(i, pending)) {
                int rc = pcm_finish(ctx, pending & 1, out + pending);
                if (rc) return rc;
                pending = -1;
            }
            int rc = pcm_enqueue(ctx, img1[i], img2[i], dims + 3 * i, dtype, params, i & 1);
            if (rc) return rc;
            if (pending >= 0 && (rc = pcm_finish(ctx, pending & 1, out + pending))) return rc;
            pending = i;
        }
        if (pending >= 0) return pcm_finish(ctx, pending & 1, out + pending);
        return BS_OK;
    }
    // host inputs: double-buffered H2D on the copy stream, overlapped with the previous pair
    size_t maxb = 0;
    for (int i = 0; i < n; ++i) maxb = std::max(maxb, crop_bytes_of(dims + 3 * i, dtype));
    int rc = ensure_crop_buffers(ctx, maxb, n > 1 ? 2 : 1);
    if (rc) return rc;
    bs_pcm_workspace& ws = ctx->ws;
    auto enqueue_copy = [&](int i) -> int {
        const int b = i & 1;
        const size_t bytes = crop_bytes_of(dims + 3 * i, dtype);
        if (i >= 2) BS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ws.crop_free[b], 0));
        BS_CUDA(ctx, cudaMemcpyAsync(ws.crop[b][0], img1[i], bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        BS_CUDA(ctx, cudaMemcpyAsync(ws.crop[b][1], img2[i], bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        BS_CUDA(ctx, cudaEventRecord(ws.crop_ready[b], ctx->copy_stream));
        return BS_OK;
    };
    if (n > 0 && (rc = enqueue_copy(0))) return rc;
    int pending = -1;
    for (int i = 0; i < n; ++i) {
        const int b = i & 1;
        if (i + 1 < n && (rc = enqueue_copy(i + 1))) return rc;
        if (pending >= 0 && !same_dims(i, pending)) {
            if ((rc = pcm_finish(ctx, pending & 1, out + pending))) return rc;
            pending = -1;
        }
        BS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ws.crop_ready[b], 0));
        rc = pcm_enqueue(ctx, ws.crop[b][0], ws.crop[b][1], dims + 3 * i, dtype, params, b);
        if (rc) return rc;
        BS_CUDA(ctx, cudaEventRecord(ws.crop_free[b], ctx->stream));
        if (pending >= 0 && (rc = pcm_finish(ctx, pending & 1, out + pending))) return rc;
        pending = i;
    }
    if (pending >= 0) return pcm_finish(ctx, pending & 1, out + pending);
    return BS_OK;
}

int bs_pcm_volumes_batch(bs_ctx* ctx, int n, const bs_pcm_job* jobs, const bs_pcm_params* params, bs_pcm_result* out) {
    if (!ctx) return BS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (n < 0 || (n > 0 && (!jobs || !params || !out))) return bs_set_error(ctx, BS_ERR_ARG, "bs_pcm_volumes_batch: NULL argument");
    if (n == 0) return BS_OK;
    BS_CUDA(ctx, cudaSetDevice(ctx->device));
    // validate everything before any launch
    size_t maxb = 0;
    bool need_crop = false;
    int dtype = -1;
    for (int i = 0; i < n; ++i) {
        const bs_pcm_job& j = jobs[i];
        auto i1 = ctx->vols.find(j.vol1), i2 = ctx->vols.find(j.vol2);
        if (i1 == ctx->vols.end() || i2 == ctx->vols.end())
            return bs_set_error(ctx, BS_ERR_ARG, "bs_pcm_volumes_batch: job %d has an unknown volume handle", i);
        if (i1->second.dtype != i2->second.dtype) return bs_set_error(ctx, BS_ERR_ARG, "bs_pcm_volumes_batch: job %d mixes dtypes", i);
        if (dtype >= 0 && i1->second.dtype != dtype) return bs_set_error(ctx, BS_ERR_ARG, "bs_pcm_volumes_batch: mixed dtypes in one batch");
        dtype = i1->second.dtype;
        for (int d = 0; d < 3; ++d) {
            if (j.dims[d] <= 0 || j.min1[d] < 0 || j.min2[d] < 0 || j.min1[d] + j.dims[d] > i1->second.dims[d] ||
                j.min2[d] + j.dims[d] > i2->second.dims[d])
                return bs_set_error(ctx, BS_ERR_ARG, "bs_pcm_volumes_batch: job %d: overlap interval outside the volume (axis %d)", i, d);
        }
        PcmGeometry g;
        int rc = pcm_geometry(ctx, j.dims, params->extension, &g);
        if (rc) return rc;
        const bool whole1 = j.dims[0] == i1->second.dims[0] && j.dims[1] == i1->second.dims[1] && j.dims[2] == i1->second.dims[2];
        const bool whole2 = j.dims[0] == i2->second.dims[0] && j.dims[1] == i2->second.dims[1] && j.dims[2] == i2->second.dims[2];
        if (!whole1 || !whole2) {
            need_crop = true;
            maxb = std::max(maxb, crop_bytes_of(j.dims, dtype));
        }
    }
    int rc = pcm_check_params(ctx, params, dtype);
    if (rc) return rc;
    if (need_crop && (rc = ensure_crop_buffers(ctx, maxb, 2))) return rc;
    bs_pcm_workspace& ws = ctx->ws;
    const size_t es = dtype == BS_DTYPE_U16 ? 2 : dtype == BS_DTYPE_F32 ? 4 : 1;
    // the overlap crop of a volume: the volume itself when the interval covers it, else a strided device copy
    auto crop = [&](bs_volume& v, const long long mn[3], const long long dims[3], void* dst, const void** out_ptr) -> int {
        if (dims[0] == v.dims[0] && dims[1] == v.dims[1] && dims[2] == v.dims[2]) { *out_ptr = v.dev; return BS_OK; }
        cudaMemcpy3DParms p;
        memset(&p, 0, sizeof(p));
        p.srcPtr = make_cudaPitchedPtr(v.dev, (size_t)v.dims[0] * es, (size_t)v.dims[0], (size_t)v.dims[1]);
        p.srcPos = make_cudaPos((size_t)mn[0] * es, (size_t)mn[1], (size_t)mn[2]);
        p.dstPtr = make_cudaPitchedPtr(dst, (size_t)dims[0] * es, (size_t)dims[0], (size_t)dims[1]);
        p.extent = make_cudaExtent((size_t)dims[0] * es, (size_t)dims[1], (size_t)dims[2]);
        p.kind = cudaMemcpyDeviceToDevice;
        BS_CUDA(ctx, cudaMemcpy3DAsync(&p, ctx->stream));
        ctx->launches++;
        *out_ptr = dst;
        return BS_OK;
    };
    auto same_dims = [&](int i, int j) {
        return jobs[i].dims[0] == jobs[j].dims[0] && jobs[i].dims[1] == jobs[j].dims[1] && jobs[i].dims[2] == jobs[j].dims[2];
    };
    int pending = -1;
    for (int i = 0; i < n; ++i) {
        const bs_pcm_job& j = jobs[i];
        bs_volume& v1 = ctx->vols.find(j.vol1)->second;
        bs_volume& v2 = ctx->vols.find(j.vol2)->second;
        if ((rc = bs_volume_acquire(ctx, v1)) || (rc = bs_volume_acquire(ctx, v2))) return rc;
        if (pending >= 0 && !same_dims(i, pending)) {
            if ((rc = pcm_finish(ctx, pending & 1, out + pending))) return rc;
            pending = -1;
        }
        const int b = i & 1;
        const void *p1 = nullptr, *p2 = nullptr;
        // crop buffers of slot b were last read by pair i-2's Pearson kernel: same stream, already ordered
        if ((rc = crop(v1, j.min1, j.dims, need_crop ? ws.crop[b][0] : nullptr, &p1))) return rc;
        if ((rc = crop(v2, j.min2, j.dims, need_crop ? ws.crop[b][1] : nullptr, &p2))) return rc;
        if ((rc = pcm_enqueue(ctx, p1, p2, j.dims, dtype, params, b))) return rc;
        if (pending >= 0 && (rc = pcm_finish(ctx, pending & 1, out + pending))) return rc;
        pending = i;
    }
    if (pending >= 0) return pcm_finish(ctx, pending & 1, out + pending);
    return BS_OK;
}

int bs_pcm_debug_pcm(bs_ctx* ctx, const void* img1, const void* img2, const long long dims[3], int dtype,
                     const int extension[3], float* out_pcm, int pad_out[3]) {
    if (!ctx) return BS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!img1 || !img2 || !dims || !extension || !out_pcm)
        return bs_set_error(ctx, BS_ERR_ARG, "bs_pcm_debug_pcm: NULL argument");
    BS_CUDA(ctx, cudaSetDevice(ctx->device));
    PcmGeometry g;
    int rc = pcm_geometry(ctx, dims, extension, &g);
    if (rc) return rc;
    PcmDeviceTables* t;
    if ((rc = pcm_tables(ctx, g, &t))) return rc;
    if ((rc = pcm_workspace(ctx, g))) return rc;
    const size_t bytes = crop_bytes_of(dims, dtype);
    if ((rc = ensure_crop_buffers(ctx, bytes, 1))) return rc;
    bs_pcm_workspace& ws = ctx->ws;
    BS_CUDA(ctx, cudaMemcpyAsync(ws.crop[0][0], img1, bytes, cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(ctx, cudaMemcpyAsync(ws.crop[0][1], img2, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = pcm_compute_pcm(ctx, ws.crop[0][0], ws.crop[0][1], dtype, g, t))) return rc;
    BS_CUDA(ctx, cudaMemcpy2DAsync(out_pcm, sizeof(float) * g.P[0], ws.spec_a, sizeof(float2) * g.pitch,
                                   sizeof(float) * g.P[0], (size_t)g.P[1] * g.P[2], cudaMemcpyDeviceToHost, ctx->stream));
    BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (pad_out) for (int d = 0; d < 3; ++d) pad_out[d] = g.P[d];
    return BS_OK;
}

}  // extern "C"
