// tools/mma_probe.cu — issue rates that decide where the depthwise stencils should run on B200:
//   (1) mma.sync.m16n8k16 bf16 (HMMA) with NACC independent accumulators per warp, 4..16 warps per SM
//   (2) the same with one ldmatrix.x4 per HMMA (the Toeplitz-stencil operand pattern)
//   (3) fma.rn.f32x2 (FFMA2) versus scalar FFMA, all-register operands
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <int NACC>
__global__ void k_hmma(const uint32_t* __restrict__ in, float* out, int iters) {
    uint32_t a[4], b[2];
    float d[NACC][4];
    const int t = threadIdx.x;
    for (int i = 0; i < 4; ++i) a[i] = in[(t + i) & 1023];
    for (int i = 0; i < 2; ++i) b[i] = in[(t * 3 + i) & 1023];
#pragma unroll
    for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) d[n][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < NACC; ++n) hmma(d[n], a, b);
    }
    float s = 0;
#pragma unroll
    for (int n = 0; n < NACC; ++n) s += d[n][0] + d[n][1] + d[n][2] + d[n][3];
    out[blockIdx.x * blockDim.x + t] = s;
}

// one ldmatrix.x4 (fresh A fragment, per-row addresses at a 144-byte pitch) per HMMA
template <int NACC>
__global__ void k_hmma_ldsm(const uint32_t* __restrict__ in, float* out, int iters) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint32_t b[2];
    float d[NACC][4];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < 144 * 80 / 4 * (blockDim.x / 32) / 1; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = in[i & 1023];
    __syncthreads();
    for (int i = 0; i < 2; ++i) b[i] = in[(t * 3 + i) & 1023];
#pragma unroll
    for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) d[n][i] = 0.f;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + warp * 144 * 80 + (lane & 15) * 144 + (lane >> 4) * 16;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < NACC; ++n) {
            uint32_t a[4];
            ldsm4(a, base + (n % 8) * 16 + ((it + n) & 31) * 144);
            hmma(d[n], a, b);
        }
    }
    float s = 0;
#pragma unroll
    for (int n = 0; n < NACC; ++n) s += d[n][0] + d[n][1] + d[n][2] + d[n][3];
    out[blockIdx.x * blockDim.x + t] = s;
}

template <int NACC>
__global__ void k_ffma(const float* __restrict__ in, float* out, int iters) {
    float w[8], x[8], acc[NACC];
    const int t = threadIdx.x;
    for (int i = 0; i < 8; ++i) { w[i] = in[(t + i) & 1023]; x[i] = in[(t * 5 + i) & 1023]; }
#pragma unroll
    for (int c = 0; c < NACC; ++c) acc[c] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c = 0; c < NACC; ++c) acc[c] = fmaf(w[j], x[(c + j) & 7], acc[c]);
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < NACC; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + t] = s;
}

template <int NACC>  // NACC packed accumulators = 2 * NACC FMAs per step
__global__ void k_ffma2(const float* __restrict__ in, float* out, int iters) {
    unsigned long long w[8], x[8], acc[NACC];
    const int t = threadIdx.x;
    for (int i = 0; i < 8; ++i) {
        float2 a = make_float2(in[(t + i) & 1023], in[(t + i + 9) & 1023]), b = make_float2(in[(t * 5 + i) & 1023], in[(t * 7 + i) & 1023]);
        w[i] = *reinterpret_cast<unsigned long long*>(&a);
        x[i] = *reinterpret_cast<unsigned long long*>(&b);
    }
#pragma unroll
    for (int c = 0; c < NACC; ++c) acc[c] = 0ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c = 0; c < NACC; ++c)
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;\n" : "+l"(acc[c]) : "l"(w[j]), "l"(x[(c + j) & 7]));
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < NACC; ++c) { float2 v = *reinterpret_cast<float2*>(&acc[c]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + t] = s;
}

template <class F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    printf("%s SMs=%d clock=%d MHz\n", p.name, sms, clk_khz / 1000);
    uint32_t* in; float* out;
    cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4); cudaMalloc(&out, (size_t)sms * 8 * 1024 * 4);
    const int iters = 20000;
#define HM(NACC, threads) { \
        float ms = time_ms([&] { k_hmma<NACC><<<sms, threads>>>(in, out, iters); }); \
        double mac = (double)sms * (threads / 32) * iters * NACC * 2048.0; \
        printf("hmma      nacc=%2d warps/SM=%2d: %7.3f ms  %7.1f TMAC/s  = %6.0f MAC/clk/SM @%d MHz (%.0f TFLOP/s)\n", NACC, threads / 32, ms, \
               mac / ms * 1e-9, mac / ms * 1e-3 / sms / (clk_khz * 1e-3) * 1e-3, clk_khz / 1000, 2 * mac / ms * 1e-9); }
    HM(1, 128) HM(2, 128) HM(4, 128) HM(8, 128) HM(4, 256) HM(8, 256) HM(4, 512) HM(8, 512) HM(2, 1024) HM(4, 1024)
#define HL(NACC, threads) { \
        size_t smb = (size_t)(threads / 32) * 144 * 80 + 144 * 40; \
        cudaFuncSetAttribute(k_hmma_ldsm<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb); \
        float ms = time_ms([&] { k_hmma_ldsm<NACC><<<sms, threads, smb>>>(in, out, iters); }); \
        double mac = (double)sms * (threads / 32) * iters * NACC * 2048.0; \
        printf("hmma+ldsm nacc=%2d warps/SM=%2d: %7.3f ms  %7.1f TMAC/s  = %6.0f MAC/clk/SM (%.1f ldsm.x4/clk/SM x1e-3)\n", NACC, threads / 32, ms, \
               mac / ms * 1e-9, mac / ms * 1e-3 / sms / (clk_khz * 1e-3) * 1e-3, mac / 2048 / ms * 1e-3 / sms / (clk_khz * 1e-3)); }
    HL(4, 128) HL(8, 128) HL(4, 256) HL(8, 256) HL(4, 512) HL(8, 512)
#define FM(KN, NACC, threads, per) { \
        float ms = time_ms([&] { KN<NACC><<<sms, threads>>>((const float*)in, out, iters); }); \
        double fma = (double)sms * threads * iters * NACC * 8.0 * per; \
        printf("%-9s nacc=%2d warps/SM=%2d: %7.3f ms  %7.2f TFMA/s = %6.1f FMA/clk/SM\n", #KN, NACC, threads / 32, ms, fma / ms * 1e-9, \
               fma / ms * 1e-3 / sms / (clk_khz * 1e-3) * 1e-3); }
    FM(k_ffma, 4, 512, 1) FM(k_ffma, 8, 512, 1) FM(k_ffma, 8, 1024, 1)
    FM(k_ffma2, 2, 512, 2) FM(k_ffma2, 4, 512, 2) FM(k_ffma2, 8, 512, 2) FM(k_ffma2, 4, 1024, 2)
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
