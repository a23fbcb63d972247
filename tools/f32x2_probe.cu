// Microbenchmark: issue throughput of packed FP32 (FADD2/FFMA2) against scalar FADD/FFMA on sm_100a,
// alone and mixed with shared-memory loads (the FFT stages' instruction mix).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/f32x2_probe tools/f32x2_probe.cu && /tmp/f32x2_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7};"
        " fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, int iters, float seed) {
    __shared__ float2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(seed * i, seed);
    __syncthreads();
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * seed + i, seed - i);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(seed, -seed);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0 || MODE == 2) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
                else a[i] = fma2(a[i], m, c);
            }
            if (MODE >= 2) {   // 4 LDS.64 per 16 (scalar) / 8 (packed) FP instructions
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float2 v = sm[(idx + 33 * i) & 2047]; a[i].x += v.x * 0.f; idx += (int)v.y & 1; }
            }
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name) {
    float2* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    k<MODE><<<148 * 8, 256>>>(out, 16, 0.f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 0.f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 148.0 * 8 * 256 * (double)iters * 4 * 16;
    printf("%-28s %8.3f ms  %7.2f TFMA/s (scalar-FMA equivalents)\n", name, ms, fma / ms * 1e-9);
    cudaFree(out);
}

int main() {
    run<0>("scalar FFMA");
    run<1>("packed FFMA2");
    run<2>("scalar FFMA + LDS.64");
    run<3>("packed FFMA2 + LDS.64");
    return 0;
}
