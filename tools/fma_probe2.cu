// tools/fma_probe2.cu — FFMA issue rate of the depthwise-stencil register pattern (weights and window in registers,
// 4 accumulators) versus a constant-bank operand; answers "what is the real FP32 ceiling of the conv inner loop".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_probe2 fma_probe2.cu && ./fma_probe2
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_stencil(const float* __restrict__ in, float* out, int iters) {
    float w[25], win[5][8 + (NACC - 4)], acc[NACC];
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 25; ++i) w[i] = in[(t + i) & 1023];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 8 + (NACC - 4); ++c) win[r][c] = in[(t * 3 + r * 12 + c) & 1023];
#pragma unroll
    for (int c = 0; c < NACC; ++c) acc[c] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int s = 0; s < 5; ++s)
#pragma unroll
                for (int c = 0; c < NACC; ++c) acc[c] = fmaf(w[r * 5 + s], win[r][c + s], acc[c]);
        win[it & 3][it & 7] += 1e-9f;  // keeps the loop body from being hoisted
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < NACC; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + t] = s;
}

// same FLOPs, weights from the constant bank (kernel parameters)
struct W25 { float w[25]; };
template <int NACC>
__global__ void k_stencil_const(const float* __restrict__ in, float* out, int iters, const __grid_constant__ W25 cw) {
    float win[5][8 + (NACC - 4)], acc[NACC];
    const int t = threadIdx.x;
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 8 + (NACC - 4); ++c) win[r][c] = in[(t * 3 + r * 12 + c) & 1023];
#pragma unroll
    for (int c = 0; c < NACC; ++c) acc[c] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int s = 0; s < 5; ++s)
#pragma unroll
                for (int c = 0; c < NACC; ++c) acc[c] = fmaf(cw.w[r * 5 + s], win[r][c + s], acc[c]);
        win[it & 3][it & 7] += 1e-9f;
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < NACC; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + t] = s;
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
    float *in, *out;
    cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    cudaMalloc(&out, sizeof(float) * sms * 64 * 256);
    W25 cw; for (int i = 0; i < 25; ++i) cw.w[i] = 1.0f + i * 1e-3f;
    const int iters = 2000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = 128, blocks = sms * warps * 32 / threads;
        auto rep = [&](const char* name, double ms, int nacc) {
            const double fma = (double)blocks * threads * iters * 25 * nacc;
            printf("warps/SM=%2d %-28s %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM @1.93GHz)\n", warps, name, ms, fma / ms * 1e-9,
                   fma / ms * 1e-9 * 1e12 / (sms * 1.93e9));
        };
        rep("stencil regs, 4 acc", time_ms([&] { k_stencil<4><<<blocks, threads>>>(in, out, iters); }), 4);
        rep("stencil regs, 8 acc", time_ms([&] { k_stencil<8><<<blocks, threads>>>(in, out, iters); }), 8);
        rep("stencil const-bank w, 4 acc", time_ms([&] { k_stencil_const<4><<<blocks, threads>>>(in, out, iters, cw); }), 4);
        rep("stencil const-bank w, 8 acc", time_ms([&] { k_stencil_const<8><<<blocks, threads>>>(in, out, iters, cw); }), 8);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
