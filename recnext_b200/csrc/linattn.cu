// linattn.cu — the contraction core of the A-series linear attention (reference model/recattn.py:16-28 LinearAttention1 and
// :39-51 LinearAttention2, which are the same function: out_n = sum_m (q_n . k_m) v_m s^2 / (mean_m (q_n . k_m) + 1e-6), s = n^-1/2):
//
//     q, k = elu(qk_pre) + 1                          qk_pre: output of the grouped 1x1 ConvNorm `qk`, [B, 2, heads, d, n]
//     kv[i, j] = (1/n) sum_m k[i, m] v[j, m]            [d x d] per (image, head)
//     kbar[i]  = (1/n) sum_m k[i, m]
//     out[j, p] = sum_i q[i, p] kv[i, j] / (sum_i q[i, p] kbar[i] + 1e-6)  (+ pe[j, p])     [B, heads, d, n] = [B, dim, h, w]
//
// The reference spells this as elu, +1, view/unbind, two transposes, two scaled batched matmuls, a mean, a matmul, an add and a
// divide — a dozen launches over tensors of the size of the activation; here it is ONE kernel per call: a CTA owns one
// (image, head), streams k and v once through shared memory in 64-pixel chunks to build kv and kbar (fp32), then streams q
// once and writes out.  d = 20..40, n = 16..784: the contractions are tiny (K = d) — FP32 FMAs on data that is read once;
// the kernel is bound by HBM/L2 traffic (4 N e: q, k, v in, out out, + pe), not by math.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>

namespace recnext {

template <typename T> __device__ __forceinline__ float la_to_f(T v);
template <> __device__ __forceinline__ float la_to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float la_to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float la_to_f<float>(float v) { return v; }
template <typename T> __device__ __forceinline__ T la_from_f(float v);
template <> __device__ __forceinline__ __nv_bfloat16 la_from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half la_from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float la_from_f<float>(float v) { return v; }

__device__ __forceinline__ float la_elu1(float v) { return v > 0.f ? v + 1.f : expf(v); }   // elu(v) + 1 (expm1(v) + 1 = exp(v))

template <typename T, int D>
__global__ void __launch_bounds__(256, 3) recnext_linattn_kernel(const T* __restrict__ q_pre, const T* __restrict__ k_pre, const float* __restrict__ qbias,
                                                                 const float* __restrict__ kbias, const T* __restrict__ v, const T* __restrict__ pe,
                                                                 const float* __restrict__ pew, const float* __restrict__ peb, int pw,
                                                                 T* __restrict__ out, int heads, int n) {
    constexpr int TQ = (D + 3) / 4;            // 4 x 4 register tiles per kv dimension
    constexpr int DP = TQ * 4;                 // padded d
    constexpr int CH = 64;                     // pixels per chunk
    constexpr int NT = (TQ * TQ + 63) / 64;    // kv tiles per thread (64 tile threads x 4 pixel groups)
    __shared__ float ks[DP][CH + 1];           // chunk of k (phase 1) / q (phase 2): [i][pixel]; pitch 65: conflict-free stores and reads
    __shared__ float vs[DP][CH + 1];           // chunk of v: [j][pixel]
    __shared__ __align__(16) float kv[DP][DP]; // kv[i][j], scaled by 1/n
    __shared__ float kbar[DP];
    __shared__ float rden[CH];                 // phase 2: 1 / (q . kbar + 1e-6) per pixel of the chunk
    const int tid = threadIdx.x;
    const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;
    const int dim = heads * D;
    // q / k rows i of this head: base + i * n.  The two tensors are [B, dim, n] each (one [B, 2 dim, n] tensor when the caller's `qk`
    // ConvNorm wrote them together: then k_pre = q_pre + dim * n and the image stride is 2 dim n); optional per-channel biases.
    const long img = (k_pre == q_pre + (long)dim * n) ? 2l * dim * n : (long)dim * n;
    const T* qp = q_pre + (long)b * img + (long)h * D * n;
    const T* kp = k_pre + (long)b * img + (long)h * D * n;
    const float* qbp = qbias ? qbias + h * D : nullptr;
    const float* kbp = kbias ? kbias + h * D : nullptr;
    const T* vp = v + ((long)b * dim + h * D) * n;
    const float inv_n = 1.f / (float)n;
    for (int i = tid; i < (DP - D) * (CH + 1); i += 256) { ks[D + i / (CH + 1)][i % (CH + 1)] = 0.f; vs[D + i / (CH + 1)][i % (CH + 1)] = 0.f; }

    // ---- phase 1: kv = k v^T / n and kbar = mean(k).  Thread = (pixel group pg of 4, 4 x 4 tile of kv): 8 shared loads per 16 FMAs
    const int pg = tid >> 6, tt = tid & 63;
    float acc[NT][4][4], ksum[NT][4];
#pragma unroll
    for (int e = 0; e < NT; ++e)
#pragma unroll
        for (int a = 0; a < 4; ++a) { acc[e][a][0] = acc[e][a][1] = acc[e][a][2] = acc[e][a][3] = 0.f; ksum[e][a] = 0.f; }
    for (int c0 = 0; c0 < n; c0 += CH) {
        const int cn = (n - c0) < CH ? (n - c0) : CH;
        for (int i = tid; i < D * CH; i += 256) {
            const int row = i / CH, px = i - row * CH;     // consecutive threads read consecutive pixels of one channel row
            float kk = 0.f, vv = 0.f;
            if (px < cn) { kk = la_elu1(la_to_f<T>(kp[(long)row * n + c0 + px]) + (kbp ? kbp[row] : 0.f)); vv = la_to_f<T>(vp[(long)row * n + c0 + px]); }
            ks[row][px] = kk; vs[row][px] = vv;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < NT; ++e) {
            const int tile = tt + e * 64;
            if (tile < TQ * TQ) {
                const int i0 = (tile / TQ) * 4, j0 = (tile % TQ) * 4;
                for (int px = pg; px < cn; px += 4) {
                    float k4[4], v4[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) { k4[a] = ks[i0 + a][px]; v4[a] = vs[j0 + a][px]; }
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[e][a][c] = fmaf(k4[a], v4[c], acc[e][a][c]);
                        ksum[e][a] += k4[a];
                    }
                }
            }
        }
        __syncthreads();
    }
    // the four pixel groups add their partial sums in a fixed order (deterministic)
    for (int r = 0; r < 4; ++r) {
        if (pg == r) {
#pragma unroll
            for (int e = 0; e < NT; ++e) {
                const int tile = tt + e * 64;
                if (tile < TQ * TQ) {
                    const int i0 = (tile / TQ) * 4, j0 = (tile % TQ) * 4;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) kv[i0 + a][j0 + c] = (r ? kv[i0 + a][j0 + c] : 0.f) + acc[e][a][c] * inv_n;
                        if (j0 == 0) kbar[i0 + a] = (r ? kbar[i0 + a] : 0.f) + ksum[e][a] * inv_n;
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- phase 2: out[j, p] = q[:, p] . kv[:, j] / (q[:, p] . kbar + 1e-6) (+ pe); item = (pixel of the chunk, group of 4 columns j)
    const T* pep = pe ? pe + ((long)b * dim + h * D) * n : nullptr;
    T* op = out + ((long)b * dim + h * D) * n;
    for (int c0 = 0; c0 < n; c0 += CH) {
        const int cn = (n - c0) < CH ? (n - c0) : CH;
        for (int i = tid; i < D * CH; i += 256) {
            const int row = i / CH, px = i - row * CH;
            ks[row][px] = px < cn ? la_elu1(la_to_f<T>(qp[(long)row * n + c0 + px]) + (qbp ? qbp[row] : 0.f)) : 0.f;
        }
        __syncthreads();
        if (tid < CH) {
            float den = 1e-6f;
#pragma unroll
            for (int i = 0; i < D; ++i) den = fmaf(ks[i][tid], kbar[i], den);
            rden[tid] = 1.f / den;
        }
        __syncthreads();
        for (int item = tid; item < CH * TQ; item += 256) {
            const int px = item & (CH - 1), jq = item >> 6;
            if (px >= cn) continue;
            float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const float q = ks[i][px];
                const float4 w4 = *reinterpret_cast<const float4*>(&kv[i][4 * jq]);   // broadcast within a warp
                n0 = fmaf(q, w4.x, n0); n1 = fmaf(q, w4.y, n1); n2 = fmaf(q, w4.z, n2); n3 = fmaf(q, w4.w, n3);
            }
            const float r = rden[px];
            const float o[4] = {n0 * r, n1 * r, n2 * r, n3 * r};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = 4 * jq + c;
                if (j < D) {
                    float val = o[c];
                    if (pep) val += la_to_f<T>(pep[(long)j * n + c0 + px]);
                    if (pew) {   // + pe(v): depthwise 3x3 conv (+ bias) of v's plane j, evaluated at this pixel (v is L1 / L2 resident: it was just streamed)
                        const int pidx = c0 + px, yy = pidx / pw, xx = pidx - yy * pw, ph = n / pw;
                        const float* wj = pew + (long)(h * D + j) * 9;
                        const T* vj = vp + (long)j * n;
                        float acc9 = peb ? peb[h * D + j] : 0.f;
#pragma unroll
                        for (int dy = -1; dy <= 1; ++dy) {
                            const int y2 = yy + dy;
                            if (y2 < 0 || y2 >= ph) continue;
#pragma unroll
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int x2 = xx + dx;
                                if (x2 >= 0 && x2 < pw) acc9 = fmaf(wj[(dy + 1) * 3 + dx + 1], la_to_f<T>(vj[(long)y2 * pw + x2]), acc9);
                            }
                        }
                        val += acc9;
                    }
                    op[(long)j * n + c0 + px] = la_from_f<T>(val);
                }
            }
        }
        __syncthreads();
    }
}

template <typename T>
static cudaError_t la_launch_t(int B, int heads, int d, int n, const void* q, const void* k, const float* qb, const float* kb, const void* v, const void* pe,
                               const float* pew, const float* peb, int pw, void* out, cudaStream_t s) {
    const int grid = B * heads;
#define LA_CASE(DD) case DD: recnext_linattn_kernel<T, DD><<<grid, 256, 0, s>>>((const T*)q, (const T*)k, qb, kb, (const T*)v, (const T*)pe, pew, peb, pw, (T*)out, heads, n); break;
    switch (d) {
        LA_CASE(4) LA_CASE(8) LA_CASE(16) LA_CASE(20) LA_CASE(24) LA_CASE(28) LA_CASE(32) LA_CASE(40)
        default: return cudaErrorInvalidValue;
    }
#undef LA_CASE
    return cudaGetLastError();
}

// 0 ok, 1 unsupported head_dim / dtype, 2 CUDA error in *err
// tensor-core version for 16-bit activations (linattn_mma.cu)
cudaError_t linattn_mma_launch(int B, int heads, int d, int n, int dtype, const void* q, const void* k, const float* qb, const float* kb, const void* v,
                               const void* pe, const float* pew, const float* peb, int pw, void* out, cudaStream_t stream);

// q, k: [B, dim, n] each (k == q + dim * n elements: one [B, 2 dim, n] tensor); qb / kb: fp32 [dim] biases added before the elu, or null
// pew / peb (nullable): fp32 depthwise 3x3 filters [dim, 9] and biases [dim] of the `pe` ConvNorm, applied to v inside the kernel (planes pw columns wide)
int linattn_launch(int B, int dim, int heads, int n, int dtype, const void* q, const void* k, const float* qb, const float* kb, const void* v, const void* pe,
                   const float* pew, const float* peb, int pw, void* out, cudaStream_t stream, cudaError_t* err) {
    if (heads < 1 || dim % heads != 0 || dtype < 0 || dtype > 2) return 1;
    if (pew && (pw < 1 || n % pw != 0)) return 1;
    const int d = dim / heads;
    if (!(d == 4 || d == 8 || d == 16 || d == 20 || d == 24 || d == 28 || d == 32 || d == 40)) return 1;
    // 16-bit activations: the contractions run on the tensor cores (RECNEXT_LINATTN=fma keeps the FP32-FMA kernel: A/B measurements)
    static int use_fma = -1;
    if (use_fma < 0) { const char* e = getenv("RECNEXT_LINATTN"); use_fma = (e && e[0] == 'f') ? 1 : 0; }
    if (dtype != 0 && !use_fma && (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 15) == 0) {
        *err = linattn_mma_launch(B, heads, d, n, dtype, q, k, qb, kb, v, pe, pew, peb, pw, out, stream);
        return *err == cudaSuccess ? 0 : 2;
    }
    *err = dtype == 0 ? la_launch_t<float>(B, heads, d, n, q, k, qb, kb, v, pe, pew, peb, pw, out, stream)
         : dtype == 1 ? la_launch_t<__nv_bfloat16>(B, heads, d, n, q, k, qb, kb, v, pe, pew, peb, pw, out, stream)
                      : la_launch_t<__half>(B, heads, d, n, q, k, qb, kb, v, pe, pew, peb, pw, out, stream);
    return *err == cudaSuccess ? 0 : 2;
}

}  // namespace recnext
