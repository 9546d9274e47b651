// Pipe-rate micro-benchmark for the selective-scan design: MUFU.EX2, FFMA, FFMA2 and a
// scan-like mix, in thread-operations per clock per SM.  Build: nvcc -arch=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed + threadIdx.x * 1e-6f + i * 1e-3f;
    float2 w[4] = {{v[0], v[1]}, {v[2], v[3]}, {v[4], v[5]}, {v[6], v[7]}};
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // 8 independent ex2 chains
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ex2f(v[i]);
        } else if (MODE == 1) {     // 8 independent FFMA chains
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], 0.999f, 1e-3f);
        } else if (MODE == 2) {     // 4 independent FFMA2 chains (8 fp32 fma)
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = __ffma2_rn(w[i], make_float2(0.999f, 0.999f), make_float2(1e-3f, 1e-3f));
        } else if (MODE == 3) {     // scan-like: per pair 1 FMUL2 + 2 ex2 + FMUL2 + 2 FFMA2
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 e = __fmul2_rn(make_float2(seed, seed), w[i]);
                float2 a = make_float2(ex2f(e.x), ex2f(e.y));
                float2 t = __fmul2_rn(make_float2(v[0], v[0]), a);
                w[i] = __ffma2_rn(a, w[i], t);
            }
        } else if (MODE == 4) {     // ex2 with the multiply feeding it (FMUL + MUFU)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ex2f(v[i] * seed);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    s += w[0].x + w[1].y + w[2].x + w[3].y;
    if (s == 123.456f) out[0] = s;
}
template <int MODE>
void run(const char* name, int ops_per_iter, int warps_per_sm) {
    int sms = 148, iters = 20000;
    float* out; cudaMalloc(&out, 4);
    dim3 grid(sms), block(32 * warps_per_sm);
    k<MODE><<<grid, block>>>(out, 10, -0.5f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, block>>>(out, iters, -0.5f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double ops = (double)sms * block.x * iters * ops_per_iter;
    printf("%-28s warps/SM %2d  %8.3f ms  %7.2f Gop/s/SM  %6.2f ops/clk/SM (at %d MHz nominal)\n", name, warps_per_sm, ms,
           ops / ms / 1e6 / sms, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
}
int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("MUFU.EX2 x8", 8, w);
        run<4>("FMUL+MUFU.EX2 x8", 8, w);
        run<1>("FFMA x8", 8, w);
        run<2>("FFMA2 x4 (8 fma)", 8, w);
        run<3>("scan mix (8 ex2 + 20 flop)", 8, w);
    }
    return 0;
}
