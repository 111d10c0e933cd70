// tools/fma_probe.cu — measures the fp32 FMA issue ceiling of this GPU with scalar FFMA and with packed
// fma.rn.f32x2 (sm_100+), the number that bounds the depthwise stencils (DESIGN.md "FP32 pipe bound").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_probe fma_probe.cu && ./fma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
    unsigned long long dd, aa, bb;
    dd = *reinterpret_cast<unsigned long long*>(&d);
    aa = *reinterpret_cast<const unsigned long long*>(&a);
    bb = *reinterpret_cast<const unsigned long long*>(&b);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

template <int NACC>
__global__ void k_ffma(float* out, int iters, float w0, float w1) {
    float acc[NACC];
    float x[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = threadIdx.x * 1e-3f + i; x[i] = 1.0f + i * 1e-3f + threadIdx.x * 1e-6f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(x[(i + r) % NACC], (r & 1) ? w0 : w1, acc[i]);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_ffma2(float* out, int iters, float w0, float w1) {
    float2 acc[NACC];
    float2 x[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i); x[i] = make_float2(1.0f + i * 1e-3f, 1.0f + threadIdx.x * 1e-6f); }
    const float2 wa = make_float2(w0, w0), wb = make_float2(w1, w1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) ffma2(acc[i], x[(i + r) % NACC], (r & 1) ? wa : wb);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 4096;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = 256, blocks = sms * warps * 32 / threads;
        constexpr int NACC = 8;
        double ms1 = time_ms([&] { k_ffma<NACC><<<blocks, threads>>>(out, iters, 1.0001f, 0.9999f); });
        double ms2 = time_ms([&] { k_ffma2<NACC><<<blocks, threads>>>(out, iters, 1.0001f, 0.9999f); });
        const double fma1 = (double)blocks * threads * iters * 8 * NACC, fma2 = fma1 * 2;
        printf("warps/SM=%2d  FFMA: %.3f ms  %.2f TFMA/s (%.1f FMA/clk/SM @1.965GHz)   FFMA2: %.3f ms  %.2f TFMA/s (%.1f FMA/clk/SM)\n",
               warps, ms1, fma1 / ms1 * 1e-9, fma1 / ms1 * 1e-9 * 1e12 / (sms * 1.965e9), ms2, fma2 / ms2 * 1e-9,
               fma2 / ms2 * 1e-9 * 1e12 / (sms * 1.965e9));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
