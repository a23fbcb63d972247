// Hot path 2: fused affine-warp + n-linear sample + blend/content weight + accumulate
// (+ dtype convert) for one output block.  Replaces BlkAffineFusion.init... +
// BlockAlgoUtils.arrayImg at J/SparkAffineFusion.java:602-627.
//
// Layout: every source volume is a dense x-fastest array resident in HBM (uint16 / float32 /
// uint8); the output block is dense x-fastest.  One CTA renders a 32 x 8 x FUSE_ZT tile: a
// warp spans 32 consecutive x so that, for the (typical) near-axis-aligned registrations,
// the 8 taps of a warp fall into a handful of 64-128 B source segments served by L1/L2.
// Views are culled per CTA (tile corners -> source AABB) so only overlapping views are
// evaluated; per view the thread keeps the inverse affine in registers and walks z.
#include <cmath>
#include <cstring>

#include "bs_internal.cuh"

#define FUSE_TX 32
#define FUSE_TY 8
#define FUSE_ZT 8
#define FUSE_CHUNK 256
#define FUSE_MAX_LUT 256

struct FuseViewDev {
    double inv[12];        // world -> source pixel
    const void* data;
    const float* content;
    int dims[3];
    int dtype;
    float border[3];
    float range[3];
};

struct FuseArgs {
    long long bmin[3];
    int size[3];
    int fusion_type;
    int lut_n;
    const float* lut;      // device, lut_n + 2 entries
    double cmin, cscale;   // integer output conversion: (v - cmin) * cscale
    double ctop;
    void* out;
    float* acc_wi;         // accumulate mode
    float* acc_w;
};

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p, size_t i) {
    return (float)__ldg(p + i);
}

template <typename T, bool LINEAR>
__device__ __forceinline__ float sample(const T* __restrict__ d, int dx, int dy, int dz, float sx, float sy,
                                        float sz) {
    if (LINEAR) {
        float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
        float rx = sx - fx, ry = sy - fy, rz = sz - fz;
        int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
        int x1 = min(x0 + 1, dx - 1), y1 = min(y0 + 1, dy - 1), z1 = min(z0 + 1, dz - 1);
        size_t r00 = ((size_t)z0 * dy + y0) * dx, r01 = ((size_t)z0 * dy + y1) * dx;
        size_t r10 = ((size_t)z1 * dy + y0) * dx, r11 = ((size_t)z1 * dy + y1) * dx;
        float a000 = ld_as_float(d, r00 + x0), a001 = ld_as_float(d, r00 + x1);
        float a010 = ld_as_float(d, r01 + x0), a011 = ld_as_float(d, r01 + x1);
        float a100 = ld_as_float(d, r10 + x0), a101 = ld_as_float(d, r10 + x1);
        float a110 = ld_as_float(d, r11 + x0), a111 = ld_as_float(d, r11 + x1);
        float c00 = a000 + rx * (a001 - a000);
        float c01 = a010 + rx * (a011 - a010);
        float c10 = a100 + rx * (a101 - a100);
        float c11 = a110 + rx * (a111 - a110);
        float c0 = c00 + ry * (c01 - c00);
        float c1 = c10 + ry * (c11 - c10);
        return c0 + rz * (c1 - c0);
    } else {
        int xi = min(max((int)floorf(sx + 0.5f), 0), dx - 1);
        int yi = min(max((int)floorf(sy + 0.5f), 0), dy - 1);
        int zi = min(max((int)floorf(sz + 0.5f), 0), dz - 1);
        return ld_as_float(d, ((size_t)zi * dy + yi) * dx + xi);
    }
}

template <bool LINEAR>
__device__ __forceinline__ float sample_any(const void* d, int dtype, int dx, int dy, int dz, float sx,
                                            float sy, float sz) {
    if (dtype == BS_DTYPE_U16) return sample<unsigned short, LINEAR>((const unsigned short*)d, dx, dy, dz, sx, sy, sz);
    if (dtype == BS_DTYPE_F32) return sample<float, LINEAR>((const float*)d, dx, dy, dz, sx, sy, sz);
    return sample<unsigned char, LINEAR>((const unsigned char*)d, dx, dy, dz, sx, sy, sz);
}

// cosine blending weight along one axis; returns false when the total weight is 0
__device__ __forceinline__ bool blend_axis(float l, float dm1, float border, float range, int lut_n,
                                           const float* s_lut, float& w) {
    float dist = fmaxf(0.f, fminf(l - border, (dm1 - l) - border));
    if (dist == 0.f) return false;
    float rel = dist / range;
    if (rel < 1.f) {
        float f;
        if (lut_n > 0) {
            float pos = rel * (float)lut_n;
            int i = (int)pos;
            float s = pos - (float)i;
            f = s_lut[i] * (1.0f - s) + s_lut[i + 1] * s;
        } else {
            f = 0.5f * (cospif(1.0f - rel) + 1.0f);
        }
        w *= f;
    }
    return true;
}

// KIND 0: weighted average family (AVG, AVG_BLEND, *_CONTENT); KIND 1: winner family.
// ACCUM: add partial sums into acc_wi/acc_w instead of producing the final voxel.
template <int KIND, bool LINEAR, int OUT, bool ACCUM>
__global__ void __launch_bounds__(FUSE_TX* FUSE_TY)
fuse_kernel(const FuseViewDev* __restrict__ views, int nviews, FuseArgs a) {
    __shared__ int s_active[FUSE_CHUNK];
    __shared__ int s_nactive;
    __shared__ float s_lut[FUSE_MAX_LUT + 2];

    const int tid = threadIdx.y * FUSE_TX + threadIdx.x;
    const int x = blockIdx.x * FUSE_TX + threadIdx.x;
    const int y = blockIdx.y * FUSE_TY + threadIdx.y;
    const int z0 = blockIdx.z * FUSE_ZT;
    const bool valid = x < a.size[0] && y < a.size[1];
    const int ft = a.fusion_type;
    const bool use_blend = ft == BS_FUSE_AVG_BLEND || ft == BS_FUSE_AVG_BLEND_CONTENT || ft == BS_FUSE_CLOSEST_PIXEL_WINS;
    const bool use_content = ft == BS_FUSE_AVG_CONTENT || ft == BS_FUSE_AVG_BLEND_CONTENT;

    if (a.lut_n > 0)
        for (int i = tid; i < a.lut_n + 2; i += FUSE_TX * FUSE_TY) s_lut[i] = a.lut[i];

    float acc0[FUSE_ZT];  // KIND0: sum w*I ; KIND1: best value
    float acc1[FUSE_ZT];  // KIND0: sum w   ; KIND1: best weight (CLOSEST) / have flag
#pragma unroll
    for (int k = 0; k < FUSE_ZT; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; }

    const double wx = (double)(a.bmin[0] + x);
    const double wy = (double)(a.bmin[1] + y);

    for (int chunk = 0; chunk < nviews; chunk += FUSE_CHUNK) {
        __syncthreads();
        if (tid == 0) s_nactive = 0;
        __syncthreads();
        // ---- cull: source AABB of the tile's 8 corners against [0, dim-1] (+-1e-3 guard)
        for (int vi = chunk + tid; vi < min(nviews, chunk + FUSE_CHUNK); vi += FUSE_TX * FUSE_TY) {
            const FuseViewDev& v = views[vi];
            double cx0 = (double)(a.bmin[0] + (long long)blockIdx.x * FUSE_TX);
            double cy0 = (double)(a.bmin[1] + (long long)blockIdx.y * FUSE_TY);
            double cz0 = (double)(a.bmin[2] + z0);
            double cx1 = cx0 + (double)(min(FUSE_TX, a.size[0] - (int)blockIdx.x * FUSE_TX) - 1);
            double cy1 = cy0 + (double)(min(FUSE_TY, a.size[1] - (int)blockIdx.y * FUSE_TY) - 1);
            double cz1 = cz0 + (double)(min(FUSE_ZT, a.size[2] - z0) - 1);
            bool hit = true;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double lo = v.inv[4 * r + 3], hi = lo;
                double m0 = v.inv[4 * r], m1 = v.inv[4 * r + 1], m2 = v.inv[4 * r + 2];
                lo += fmin(m0 * cx0, m0 * cx1) + fmin(m1 * cy0, m1 * cy1) + fmin(m2 * cz0, m2 * cz1);
                hi += fmax(m0 * cx0, m0 * cx1) + fmax(m1 * cy0, m1 * cy1) + fmax(m2 * cz0, m2 * cz1);
                if (hi < -1e-3 || lo > (double)(v.dims[r] - 1) + 1e-3) hit = false;
            }
            if (hit) s_active[atomicAdd(&s_nactive, 1)] = vi;
        }
        __syncthreads();
        const int nact = s_nactive;
        // deterministic view order (ascending ViewId): sort the short active list
        if (tid == 0) {
            for (int i = 1; i < nact; ++i) {
                int key = s_active[i], j = i - 1;
                while (j >= 0 && s_active[j] > key) { s_active[j + 1] = s_active[j]; --j; }
                s_active[j + 1] = key;
            }
        }
        __syncthreads();
        if (!valid) continue;

        for (int ai = 0; ai < nact; ++ai) {
            const FuseViewDev& v = views[s_active[ai]];
            const double m02 = v.inv[2], m12 = v.inv[6], m22 = v.inv[10];
            const double bx = fma(v.inv[0], wx, fma(v.inv[1], wy, v.inv[3]));
            const double by = fma(v.inv[4], wx, fma(v.inv[5], wy, v.inv[7]));
            const double bz = fma(v.inv[8], wx, fma(v.inv[9], wy, v.inv[11]));
            const int dx = v.dims[0], dy = v.dims[1], dz = v.dims[2];
            const float dm1x = (float)(dx - 1), dm1y = (float)(dy - 1), dm1z = (float)(dz - 1);
            const void* data = v.data;
            const int dtype = v.dtype;
#pragma unroll
            for (int k = 0; k < FUSE_ZT; ++k) {
                if (z0 + k >= a.size[2]) break;
                const double wz = (double)(a.bmin[2] + z0 + k);
                const float sx = (float)fma(m02, wz, bx);
                const float sy = (float)fma(m12, wz, by);
                const float sz = (float)fma(m22, wz, bz);
                if (!(sx >= 0.f && sx <= dm1x && sy >= 0.f && sy <= dm1y && sz >= 0.f && sz <= dm1z)) continue;
                float w = 1.f;
                if (use_blend) {
                    if (!blend_axis(sx, dm1x, v.border[0], v.range[0], a.lut_n, s_lut, w)) continue;
                    if (!blend_axis(sy, dm1y, v.border[1], v.range[1], a.lut_n, s_lut, w)) continue;
                    if (!blend_axis(sz, dm1z, v.border[2], v.range[2], a.lut_n, s_lut, w)) continue;
                }
                const float val = sample_any<LINEAR>(data, dtype, dx, dy, dz, sx, sy, sz);
                if (KIND == 0) {
                    if (use_content) w *= sample<float, LINEAR>(v.content, dx, dy, dz, sx, sy, sz);
                    acc0[k] += w * val;
                    acc1[k] += w;
                } else {
                    if (!(w > 0.f)) continue;
                    if (ft == BS_FUSE_MAX_INTENSITY) {
                        if (acc1[k] == 0.f || val > acc0[k]) acc0[k] = val;
                        acc1[k] = 1.f;
                    } else if (ft == BS_FUSE_LOWEST_VIEWID_WINS) {
                        if (acc1[k] == 0.f) { acc0[k] = val; acc1[k] = 1.f; }
                    } else if (ft == BS_FUSE_HIGHEST_VIEWID_WINS) {
                        acc0[k] = val; acc1[k] = 1.f;
                    } else {  // CLOSEST_PIXEL_WINS: largest blending weight wins
                        if (w > acc1[k]) { acc0[k] = val; acc1[k] = w; }
                    }
                }
            }
        }
    }
    if (!valid) return;
#pragma unroll
    for (int k = 0; k < FUSE_ZT; ++k) {
        const int z = z0 + k;
        if (z >= a.size[2]) break;
        const size_t o = ((size_t)z * a.size[1] + y) * a.size[0] + x;
        if (ACCUM) {
            a.acc_wi[o] += acc0[k];
            a.acc_w[o] += acc1[k];
            continue;
        }
        float res;
        if (KIND == 0) res = acc1[k] > 0.f ? acc0[k] / acc1[k] : 0.f;
        else res = acc1[k] > 0.f ? acc0[k] : 0.f;
        if (OUT == BS_DTYPE_F32) {
            ((float*)a.out)[o] = res;
        } else {
            double c = floor(((double)res - a.cmin) * a.cscale + 0.5);
            c = fmin(fmax(c, 0.0), a.ctop);
            if (OUT == BS_DTYPE_U16) ((unsigned short*)a.out)[o] = (unsigned short)c;
            else ((unsigned char*)a.out)[o] = (unsigned char)c;
        }
    }
}

template <int OUT>
__global__ void fuse_finish_kernel(const float* __restrict__ swi, const float* __restrict__ sw, long long n,
                                   FuseArgs a) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float w = sw[i];
        float res = w > 0.f ? swi[i] / w : 0.f;
        if (OUT == BS_DTYPE_F32) {
            ((float*)a.out)[i] = res;
        } else {
            double c = floor(((double)res - a.cmin) * a.cscale + 0.5);
            c = fmin(fmax(c, 0.0), a.ctop);
            if (OUT == BS_DTYPE_U16) ((unsigned short*)a.out)[i] = (unsigned short)c;
            else ((unsigned char*)a.out)[i] = (unsigned char)c;
        }
    }
}

// ------------------------------------------------------------------------------------------
static bool invert34(const double* m, double* inv) {
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

static size_t out_elem_size(int dt) { return dt == BS_DTYPE_F32 ? 4 : dt == BS_DTYPE_U16 ? 2 : 1; }

struct FusePrepared {
    FuseArgs args;
    int nviews;
    int slot;
    const FuseViewDev* views_dev;
};

// validate + upload view descriptors (and the cosine table); fills args except out pointers
static int fuse_prepare(bs_ctx* ctx, const bs_view* views, int n_views, const long long block_min[3],
                        const long long block_size[3], const bs_fuse_params* p, FusePrepared* prep) {
    if (!views && n_views > 0) return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: views is NULL");
    if (!block_min || !block_size || !p) return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: NULL argument");
    if (n_views < 0) return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: n_views < 0");
    for (int d = 0; d < 3; ++d)
        if (block_size[d] <= 0 || block_size[d] > 0x7fffffffLL)
            return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: bad block_size[%d]=%lld", d, block_size[d]);
    if (p->fusion_type < BS_FUSE_AVG || p->fusion_type > BS_FUSE_CLOSEST_PIXEL_WINS)
        return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: unknown fusion_type %d", p->fusion_type);
    if (p->interpolation != 0 && p->interpolation != 1)
        return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: interpolation must be 0 or 1");
    if (p->out_dtype != BS_DTYPE_F32 && p->out_dtype != BS_DTYPE_U16 && p->out_dtype != BS_DTYPE_U8)
        return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: bad out_dtype %d", p->out_dtype);
    if (p->blend_lut_n < 0 || p->blend_lut_n > FUSE_MAX_LUT)
        return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: blend_lut_n out of range [0,%d]", FUSE_MAX_LUT);
    if (p->out_dtype != BS_DTYPE_F32 && !(p->max_intensity > p->min_intensity))
        return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: max_intensity must exceed min_intensity");
    const bool need_content = p->fusion_type == BS_FUSE_AVG_CONTENT || p->fusion_type == BS_FUSE_AVG_BLEND_CONTENT;

    std::vector<FuseViewDev> hv((size_t)n_views);
    for (int i = 0; i < n_views; ++i) {
        auto it = ctx->vols.find(views[i].vol_handle);
        if (it == ctx->vols.end())
            return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: view %d has unknown vol_handle %llu", i, views[i].vol_handle);
        const bs_volume& vol = it->second;
        FuseViewDev& d = hv[i];
        if (!invert34(views[i].src_to_world, d.inv))
            return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: view %d has a singular transform", i);
        d.data = vol.dev;
        d.dims[0] = (int)vol.dims[0]; d.dims[1] = (int)vol.dims[1]; d.dims[2] = (int)vol.dims[2];
        d.dtype = vol.dtype;
        d.content = nullptr;
        if (need_content) {
            auto ic = ctx->vols.find(views[i].content_handle);
            if (ic == ctx->vols.end() || ic->second.dtype != BS_DTYPE_F32 || ic->second.dims[0] != vol.dims[0] ||
                ic->second.dims[1] != vol.dims[1] || ic->second.dims[2] != vol.dims[2])
                return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: view %d needs a float32 content volume of equal dims", i);
            d.content = (const float*)ic->second.dev;
        }
        for (int k = 0; k < 3; ++k) {
            d.border[k] = views[i].blend_border[k];
            d.range[k] = views[i].blend_range[k];
        }
    }
    const size_t lut_bytes = (size_t)(FUSE_MAX_LUT + 2) * sizeof(float);
    const size_t need = lut_bytes + hv.size() * sizeof(FuseViewDev);
    if (need > ctx->fuse_slot_bytes) {
        size_t sb = 1 << 16;
        while (sb < need) sb <<= 1;
        BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->fuse_ring_dev) cudaFree(ctx->fuse_ring_dev);
        if (ctx->fuse_ring_host) cudaFreeHost(ctx->fuse_ring_host);
        ctx->fuse_ring_dev = ctx->fuse_ring_host = nullptr;
        ctx->fuse_slot_bytes = 0;
        BS_CUDA(ctx, cudaMalloc(&ctx->fuse_ring_dev, sb * bs_ctx::kFuseSlots));
        BS_CUDA(ctx, cudaHostAlloc(&ctx->fuse_ring_host, sb * bs_ctx::kFuseSlots, cudaHostAllocDefault));
        for (int i = 0; i < bs_ctx::kFuseSlots; ++i) {
            ctx->fuse_slot_used[i] = false;
            if (!ctx->fuse_slot_ev[i])
                BS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fuse_slot_ev[i], cudaEventDisableTiming));
        }
        ctx->fuse_slot_bytes = sb;
    }
    const int slot = ctx->fuse_next_slot;
    ctx->fuse_next_slot = (slot + 1) % bs_ctx::kFuseSlots;
    if (ctx->fuse_slot_used[slot]) BS_CUDA(ctx, cudaEventSynchronize(ctx->fuse_slot_ev[slot]));
    unsigned char* hslot = (unsigned char*)ctx->fuse_ring_host + (size_t)slot * ctx->fuse_slot_bytes;
    unsigned char* dslot = (unsigned char*)ctx->fuse_ring_dev + (size_t)slot * ctx->fuse_slot_bytes;
    float* lut = (float*)hslot;
    const int n = p->blend_lut_n;
    if (n > 0) {
        for (int i = 0; i <= n; ++i) lut[i] = (float)((std::cos((1.0 - (double)i / n) * M_PI) + 1.0) / 2.0);
        lut[n + 1] = lut[n];
    }
    if (!hv.empty()) memcpy(hslot + lut_bytes, hv.data(), hv.size() * sizeof(FuseViewDev));
    BS_CUDA(ctx, cudaMemcpyAsync(dslot, hslot, need, cudaMemcpyHostToDevice, ctx->stream));
    prep->slot = slot;
    prep->views_dev = (const FuseViewDev*)(dslot + lut_bytes);

    FuseArgs& a = prep->args;
    memset(&a, 0, sizeof(a));
    for (int d = 0; d < 3; ++d) { a.bmin[d] = block_min[d]; a.size[d] = (int)block_size[d]; }
    a.fusion_type = p->fusion_type;
    a.lut_n = n;
    a.lut = (const float*)dslot;
    a.ctop = p->out_dtype == BS_DTYPE_U8 ? 255.0 : 65535.0;
    a.cmin = p->min_intensity;
    a.cscale = p->out_dtype == BS_DTYPE_F32 ? 1.0 : a.ctop / (p->max_intensity - p->min_intensity);
    prep->nviews = n_views;
    return BS_OK;
}

template <int KIND, bool LINEAR, bool ACCUM>
static void launch_out(int out_dtype, dim3 grid, dim3 block, cudaStream_t s, const FuseViewDev* v, int n,
                       const FuseArgs& a) {
    if (ACCUM || out_dtype == BS_DTYPE_F32) fuse_kernel<KIND, LINEAR, BS_DTYPE_F32, ACCUM><<<grid, block, 0, s>>>(v, n, a);
    else if (out_dtype == BS_DTYPE_U16) fuse_kernel<KIND, LINEAR, BS_DTYPE_U16, false><<<grid, block, 0, s>>>(v, n, a);
    else fuse_kernel<KIND, LINEAR, BS_DTYPE_U8, false><<<grid, block, 0, s>>>(v, n, a);
}

static int fuse_launch(bs_ctx* ctx, const FusePrepared& prep, const bs_fuse_params* p, bool accum) {
    const FuseArgs& a = prep.args;
    dim3 block(FUSE_TX, FUSE_TY, 1);
    dim3 grid((a.size[0] + FUSE_TX - 1) / FUSE_TX, (a.size[1] + FUSE_TY - 1) / FUSE_TY,
              (a.size[2] + FUSE_ZT - 1) / FUSE_ZT);
    if (grid.y > 65535 || grid.z > 65535)
        return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse: block too large for one launch (y/z tiles > 65535)");
    const FuseViewDev* v = prep.views_dev;
    const bool winner = p->fusion_type >= BS_FUSE_MAX_INTENSITY;
    const bool lin = p->interpolation == 1;
    {
        bs_launch_scope scope(ctx, "fuse");
        if (accum) {
            if (lin) launch_out<0, true, true>(p->out_dtype, grid, block, ctx->stream, v, prep.nviews, a);
            else launch_out<0, false, true>(p->out_dtype, grid, block, ctx->stream, v, prep.nviews, a);
        } else if (!winner) {
            if (lin) launch_out<0, true, false>(p->out_dtype, grid, block, ctx->stream, v, prep.nviews, a);
            else launch_out<0, false, false>(p->out_dtype, grid, block, ctx->stream, v, prep.nviews, a);
        } else {
            if (lin) launch_out<1, true, false>(p->out_dtype, grid, block, ctx->stream, v, prep.nviews, a);
            else launch_out<1, false, false>(p->out_dtype, grid, block, ctx->stream, v, prep.nviews, a);
        }
    }
    BS_CUDA(ctx, cudaGetLastError());
    BS_CUDA(ctx, cudaEventRecord(ctx->fuse_slot_ev[prep.slot], ctx->stream));
    ctx->fuse_slot_used[prep.slot] = true;
    return BS_OK;
}

extern "C" {

void bs_fuse_default_params(bs_fuse_params* p) {
    if (!p) return;
    p->fusion_type = BS_FUSE_AVG_BLEND;
    p->interpolation = 1;
    p->out_dtype = BS_DTYPE_F32;
    p->blend_lut_n = 0;
    p->min_intensity = 0.0;
    p->max_intensity = 65535.0;
}

int bs_fuse_block(bs_ctx* ctx, const bs_view* views, int n_views, const long long block_min[3],
                  const long long block_size[3], const bs_fuse_params* params, void* out, int out_on_device) {
    if (!ctx) return BS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!out) return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse_block: out is NULL");
    BS_CUDA(ctx, cudaSetDevice(ctx->device));
    FusePrepared prep;
    int rc = fuse_prepare(ctx, views, n_views, block_min, block_size, params, &prep);
    if (rc) return rc;
    const size_t bytes = (size_t)block_size[0] * block_size[1] * block_size[2] * out_elem_size(params->out_dtype);
    if (out_on_device) {
        prep.args.out = out;
    } else {
        rc = bs_ensure_dev(ctx, &ctx->fuse_out, &ctx->fuse_out_cap, bytes);
        if (rc) return rc;
        prep.args.out = ctx->fuse_out;
    }
    rc = fuse_launch(ctx, prep, params, false);
    if (rc) return rc;
    if (!out_on_device) {
        BS_CUDA(ctx, cudaMemcpyAsync(out, ctx->fuse_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return BS_OK;
}

int bs_fuse_accumulate(bs_ctx* ctx, const bs_view* views, int n_views, const long long block_min[3],
                       const long long block_size[3], const bs_fuse_params* params, float* sum_wi_dev,
                       float* sum_w_dev) {
    if (!ctx) return BS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!sum_wi_dev || !sum_w_dev) return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse_accumulate: NULL accumulator");
    if (params && params->fusion_type >= BS_FUSE_MAX_INTENSITY)
        return bs_set_error(ctx, BS_ERR_UNSUPPORTED,
                            "bs_fuse_accumulate: only the weighted-average fusion types reduce by sum");
    BS_CUDA(ctx, cudaSetDevice(ctx->device));
    FusePrepared prep;
    int rc = fuse_prepare(ctx, views, n_views, block_min, block_size, params, &prep);
    if (rc) return rc;
    prep.args.acc_wi = sum_wi_dev;
    prep.args.acc_w = sum_w_dev;
    return fuse_launch(ctx, prep, params, true);
}

int bs_fuse_finish(bs_ctx* ctx, const float* sum_wi_dev, const float* sum_w_dev, long long n,
                   const bs_fuse_params* params, void* out, int out_on_device) {
    if (!ctx) return BS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!sum_wi_dev || !sum_w_dev || !out || !params || n <= 0)
        return bs_set_error(ctx, BS_ERR_ARG, "bs_fuse_finish: bad argument");
    BS_CUDA(ctx, cudaSetDevice(ctx->device));
    FuseArgs a;
    memset(&a, 0, sizeof(a));
    a.ctop = params->out_dtype == BS_DTYPE_U8 ? 255.0 : 65535.0;
    a.cmin = params->min_intensity;
    a.cscale = params->out_dtype == BS_DTYPE_F32 ? 1.0 : a.ctop / (params->max_intensity - params->min_intensity);
    const size_t bytes = (size_t)n * out_elem_size(params->out_dtype);
    if (out_on_device) {
        a.out = out;
    } else {
        int rc = bs_ensure_dev(ctx, &ctx->fuse_out, &ctx->fuse_out_cap, bytes);
        if (rc) return rc;
        a.out = ctx->fuse_out;
    }
    int threads = 256;
    int blocks = (int)std::min<long long>((n + threads - 1) / threads, (long long)ctx->sm_count * 16);
    {
        bs_launch_scope scope(ctx, "fuse_finish");
        if (params->out_dtype == BS_DTYPE_F32) fuse_finish_kernel<BS_DTYPE_F32><<<blocks, threads, 0, ctx->stream>>>(sum_wi_dev, sum_w_dev, n, a);
        else if (params->out_dtype == BS_DTYPE_U16) fuse_finish_kernel<BS_DTYPE_U16><<<blocks, threads, 0, ctx->stream>>>(sum_wi_dev, sum_w_dev, n, a);
        else fuse_finish_kernel<BS_DTYPE_U8><<<blocks, threads, 0, ctx->stream>>>(sum_wi_dev, sum_w_dev, n, a);
    }
    BS_CUDA(ctx, cudaGetLastError());
    if (!out_on_device) {
        BS_CUDA(ctx, cudaMemcpyAsync(out, ctx->fuse_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        BS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return BS_OK;
}

}  // extern "C"
