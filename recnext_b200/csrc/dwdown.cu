// dwdown.cu — the depthwise 7x7 stride-2 convolution of a RecNeXt `Downsample` block with channel multiplier 2 and the
// eval-mode BatchNorm that follows it folded into (w, b):   reference model/recnext.py:134-146
//     out[n, 2c + m, i, j] = b[2c + m] + sum_{r,s} w[2c + m, r, s] * x[n, c, 2i + r - 3, 2j + s - 3],   m = 0, 1
// (nn.Conv2d(C, 2C, 7, padding=3, stride=2, groups=C) followed by BatchNorm2d(2C)).  PyTorch dispatches this to its generic
// depthwise kernel: 0.62 ms per launch at [256, 64, 56, 56], 16 % of the RecNeXt-M3 inference step
// (profiles/r1_e_launches_bench_step.txt).  Here a CTA owns one input plane (several for small planes): a plane lands in shared memory as padded
// fp32 (zero border), each thread produces a 2 x 2 block of BOTH output channels from one 9 x 9 register window (81 shared
// loads for 392 FMAs), the filters come from shared memory as broadcasts.  Memory bound by design: x read once, out written once.
#include <cuda_runtime.h>
#include "devcfg.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace recnext {

template <typename T> __device__ __forceinline__ float dd_to_f(T v);
template <> __device__ __forceinline__ float dd_to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float dd_to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float dd_to_f<float>(float v) { return v; }
template <typename T> __device__ __forceinline__ T dd_from_f(float v);
template <> __device__ __forceinline__ __nv_bfloat16 dd_from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half dd_from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float dd_from_f<float>(float v) { return v; }

template <typename T>
__global__ void __launch_bounds__(256) recnext_dwdown_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                             T* __restrict__ out, int C, int H, int W, int Ho, int Wo, int pitch, int PP, long nplanes) {
    extern __shared__ __align__(16) float sm[];
    const int rows = H + 8;                    // two extra zero rows: the last 2x2 output block of an odd-sized plane reads past the padding
    const int pstride = 112 + rows * pitch;    // floats per plane slot: [2][49] filters (+ pad), then the padded plane, interior at (+3, +3)
    const int tid = threadIdx.x;
    const long plane0 = (long)blockIdx.x * PP; // the CTA owns PP consecutive (n, c) planes (small planes are batched to fill the threads)
    for (int i = tid; i < PP * pstride; i += 256) sm[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < PP * 98; i += 256) {
        const int p = i / 98, e = i - p * 98;
        if (plane0 + p < nplanes) sm[p * pstride + e] = w[(long)(2 * ((plane0 + p) % C)) * 49 + e];
    }
    const int HWp = H * W;
    for (int i = tid; i < PP * HWp; i += 256) {
        const int p = i / HWp, k = i - p * HWp, r = k / W, col = k - r * W;
        if (plane0 + p < nplanes) sm[p * pstride + 112 + (r + 3) * pitch + col + 3] = dd_to_f<T>(x[(plane0 + p) * HWp + k]);
    }
    __syncthreads();
    const int bw = (Wo + 1) >> 1, bh = (Ho + 1) >> 1, nblk = bw * bh;   // 2 x 2 output blocks per plane
    for (int it = tid; it < PP * nblk; it += 256) {
        const int p = it / nblk, blk = it - p * nblk;
        const long plane = plane0 + p;
        if (plane >= nplanes) break;
        const int c = (int)(plane % C);
        const float* ws = sm + p * pstride;
        const float b0 = b[2 * c], b1 = b[2 * c + 1];
        const int by = blk / bw, bx = blk - by * bw;
        const int oy = 2 * by, ox = 2 * bx;
        const float* base = ws + 112 + (2 * oy) * pitch + 2 * ox;   // window rows 2oy .. 2oy+8, cols 2ox .. 2ox+8 (padded coordinates)
        float a0[2][2] = {{b0, b0}, {b0, b0}}, a1[2][2] = {{b1, b1}, {b1, b1}};
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            float v[9];
#pragma unroll
            for (int s2 = 0; s2 < 9; ++s2) v[s2] = base[r * pitch + s2];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int fr = r - 2 * dy;   // filter row for output row oy + dy
                if (fr < 0 || fr > 6) continue;
#pragma unroll
                for (int s2 = 0; s2 < 7; ++s2) {
                    const float w0 = ws[fr * 7 + s2], w1 = ws[49 + fr * 7 + s2];
                    a0[dy][0] = fmaf(w0, v[s2], a0[dy][0]); a0[dy][1] = fmaf(w0, v[s2 + 2], a0[dy][1]);
                    a1[dy][0] = fmaf(w1, v[s2], a1[dy][0]); a1[dy][1] = fmaf(w1, v[s2 + 2], a1[dy][1]);
                }
            }
        }
        T* o0 = out + ((plane / C) * 2 * C + 2 * c) * (long)(Ho * Wo);
        T* o1 = o0 + (long)Ho * Wo;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx)
                if (oy + dy < Ho && ox + dx < Wo) {
                    o0[(oy + dy) * Wo + ox + dx] = dd_from_f<T>(a0[dy][dx]);
                    o1[(oy + dy) * Wo + ox + dx] = dd_from_f<T>(a1[dy][dx]);
                }
    }
}

// 0 ok, 1 unsupported (plane does not fit), 2 CUDA error in *err
int dwdown_launch(int B, int C, int H, int W, int dtype, const void* x, const float* w, const float* b, void* out, cudaStream_t stream, cudaError_t* err) {
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    int pitch = W + 8;
    if ((pitch & 1) == 0) ++pitch;                       // odd pitch: the stride-2 window rows of neighbouring threads spread over the banks
    const size_t pbytes = (112 + (size_t)(H + 8) * pitch) * sizeof(float);
    if (pbytes > 227 * 1024 || dtype < 0 || dtype > 2) return 1;
    const int nblk = ((Wo + 1) / 2) * ((Ho + 1) / 2);
    int PP = 256 / nblk;                                 // planes per CTA: enough 2x2 blocks for every thread
    if (PP < 1) PP = 1;
    if (PP > 16) PP = 16;
    while (PP > 1 && PP * pbytes > 72 * 1024) --PP;      // keep three CTAs per SM
    const long nplanes = (long)B * C;
    const size_t smem = PP * pbytes;
    static DeviceOnce configured = {};
    *err = rc_once_per_device(configured, [] {
        cudaError_t e = cudaFuncSetAttribute(recnext_dwdown_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(recnext_dwdown_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(recnext_dwdown_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        return e;
    });
    if (*err != cudaSuccess) return 2;
    const int grid = (int)((nplanes + PP - 1) / PP);
    if (dtype == 0) recnext_dwdown_kernel<float><<<grid, 256, smem, stream>>>((const float*)x, w, b, (float*)out, C, H, W, Ho, Wo, pitch, PP, nplanes);
    else if (dtype == 1) recnext_dwdown_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>((const __nv_bfloat16*)x, w, b, (__nv_bfloat16*)out, C, H, W, Ho, Wo, pitch, PP, nplanes);
    else recnext_dwdown_kernel<__half><<<grid, 256, smem, stream>>>((const __half*)x, w, b, (__half*)out, C, H, W, Ho, Wo, pitch, PP, nplanes);
    *err = cudaGetLastError();
    return *err == cudaSuccess ? 0 : 2;
}

}  // namespace recnext
