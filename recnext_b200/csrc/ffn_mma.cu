// ffn_mma.cu — fused channel mixer of a RecNeXt block on NCHW tensors:
//
//     out[b, c, p] = x[b, c, p] + b2[c] + sum_h W2[c, h] * gelu(b1[h] + sum_k W1[h, k] * y[b, k, p])
//
// i.e. reference model/recnext.py:157-158  x + channel_mixer(norm(token_mixer(x)))  with y = token_mixer(x) (the RecConv
// output), the eval-mode BatchNorm `norm` folded into (W1, b1) by the host (per-channel affine: W1' = W1 diag(s),
// b1' = b1 + W1 t) and both ConvNorms of `mlp` (:125-131) already folded by ConvNorm.fuse (:75-97).
//
// Why a kernel: the reference evaluates this as conv1x1 -> GELU -> conv1x1 -> add with a BatchNorm in front; on B200
// cuDNN's bf16 1x1 path converts NCHW -> NHWC and back around EVERY conv, and those two conversion kernels alone are
// 47 % of the RecNeXt-M3 inference step (profiles/r1_d_launches_bench_step.txt).  In NCHW one image is a row-major
// [C x HW] matrix, so the two 1x1 convs are plain GEMMs  H = W1 Y,  O = W2 H  whose activation operand is "K x N with N
// contiguous": exactly what ldmatrix.trans feeds to mma.sync.  A CTA takes a tile of <= 64 pixels of one image: the
// Y and X tiles arrive with cp.async (whole 16/8-byte chunks of contiguous pixel rows), the hidden activation
// (bias + exact-erf GELU, rounded to the activation dtype like the reference's autocast graph) lives only in shared
// memory, the residual is added in the epilogue and the result leaves as coalesced row chunks.  HBM traffic: read y,
// read x, write out — 3 N e instead of ~14 N e.  Weights are read as MMA A-fragments straight from global memory
// (L1/L2 resident, each fragment reused for all n-tiles of a unit).
// Warp-level mma.sync (HMMA), not tcgen05: K = C is 64..256 and the pixel tile is tiny, the kernel is bound by HBM and
// by the GELU's FP32 work, not by tensor throughput; a tcgen05/TMEM version is the next step for the wide stages.
#include <cuda_runtime.h>
#include "devcfg.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>

namespace recnext {

struct FfnPlan {
    int B, C, HID, HW;
    int NTN;          // n-tiles (8 pixels) per CTA tile
    int tiles;        // CTA tiles per image
    int chunkB;       // bytes per cp.async chunk: 16 (HW % 8 == 0), 8 (HW % 4 == 0)
    int PB;           // row pitch (bytes) of the shared-memory tiles: odd number of 16-byte chunks
    int offX, offH;   // byte offsets of the X/O tile and the H tile
    int smem_bytes;
    int dtype;        // 1 bf16, 2 f16
    int NQ;           // n-tiles swept per unit: 8 or 4
    int staged, offW; // wide stages: weights stream through shared memory (kFfnStages chunk buffers at offW)
    int kc;           // staged: K-chunk (64 or 32)
};

template <typename T> struct FfnT;
template <> struct FfnT<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
    static __device__ __forceinline__ float2 unpack(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
    static __device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};
template <> struct FfnT<__half> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __half2 v = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
    static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
    static __device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};

__device__ __forceinline__ void f_ldsm4t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void f_cp_async(uint32_t dst, const void* src, int bytes) {
    if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");
}
// nn.GELU() (exact, erf form).  erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below the 16-bit output rounding):
// 5 FMA + one reciprocal + one exp2 instead of erff's ~35 instructions — the GELU is the largest FP32 item of the kernel.
__device__ __forceinline__ float f_gelu(float v) {
    const float x = v * 0.70710678118654752440f, ax = fabsf(x);
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * ax * -1.4426950408889634f));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(t, p, 1.421413741f);
    p = fmaf(t, p, -0.284496736f);
    p = fmaf(t, p, 0.254829592f);
    const float r = fmaf(-p * t, e, 1.f);                 // erf(|x|)
    const float hv = 0.5f * v;
    return fmaf(hv, copysignf(r, x), hv);
}

// One GEMM unit: 16 rows [m0, m0+16) of  D = W[rows x K] * S[K x pixels]  for the n-tiles [nq0, nq0 + NQ) of the tile.
// W row-major in global memory (leading dimension K), S in shared memory (row = k, pitch PB bytes, 8 pixels per chunk).
template <typename T, int NQ>
__device__ __forceinline__ void ffn_unit(float (&acc)[NQ][4], const T* __restrict__ W, int K, int m0, uint32_t S, int PB, int nq0, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const uint32_t* wa = reinterpret_cast<const uint32_t*>(W + (long)(m0 + g) * K + 2 * t);       // row g,     k = 2t
    const uint32_t* wb = reinterpret_cast<const uint32_t*>(W + (long)(m0 + g + 8) * K + 2 * t);   // row g + 8
    // ldmatrix.trans x4: lanes 0..15 address rows k0 + lane of chunk nq, lanes 16..31 the same rows of chunk nq + 1
    const uint32_t sb = S + (uint32_t)(lane & 15) * (uint32_t)PB + (uint32_t)(nq0 + (lane >> 4)) * 16u;
    // The weights stream from L2 (they do not fit in L1 next to the tiles for wide stages): a ring of PF k-steps of
    // A fragments is kept in flight so that the ~500-cycle L2 latency hides behind PF * (4 ldmatrix + NQ mma).
    constexpr int PF = 4;
    const int nk = K >> 4;
    uint32_t A[PF][4];
#pragma unroll
    for (int i = 0; i < PF; ++i)
        if (i < nk) { A[i][0] = __ldg(wa + 8 * i); A[i][1] = __ldg(wb + 8 * i); A[i][2] = __ldg(wa + 8 * i + 4); A[i][3] = __ldg(wb + 8 * i + 4); }
    for (int kb = 0; kb < nk; kb += PF) {
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            const int ks = kb + i;
            if (ks < nk) {
                const uint32_t a0 = A[i][0], a1 = A[i][1], a2 = A[i][2], a3 = A[i][3];
                if (ks + PF < nk) {
                    A[i][0] = __ldg(wa + 8 * (ks + PF)); A[i][1] = __ldg(wb + 8 * (ks + PF));
                    A[i][2] = __ldg(wa + 8 * (ks + PF) + 4); A[i][3] = __ldg(wb + 8 * (ks + PF) + 4);
                }
                const uint32_t srow = sb + (uint32_t)(ks * 16) * (uint32_t)PB;
#pragma unroll
                for (int q = 0; q < NQ; q += 2) {
                    uint32_t b0, b1, b2, b3;
                    f_ldsm4t(b0, b1, b2, b3, srow + (uint32_t)q * 16u);   // (k 0-7, nq), (k 8-15, nq), (k 0-7, nq+1), (k 8-15, nq+1)
                    FfnT<T>::mma(acc[q], a0, a1, a2, a3, b0, b1);
                    if (q + 1 < NQ) FfnT<T>::mma(acc[q + 1], a0, a1, a2, a3, b2, b3);
                }
            }
        }
    }
}

template <typename T, int NQ>
__global__ void __launch_bounds__(256) recnext_ffn_kernel(const __grid_constant__ FfnPlan pl, const T* __restrict__ y, const T* __restrict__ x,
                                                          const T* __restrict__ w1, const float* __restrict__ b1, const T* __restrict__ w2,
                                                          const float* __restrict__ b2, T* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int img = blockIdx.x / pl.tiles, tile = blockIdx.x - img * pl.tiles;
    const int C = pl.C, HID = pl.HID, HW = pl.HW, PB = pl.PB;
    const int p0 = tile * pl.NTN * 8;
    const int np = (HW - p0) < pl.NTN * 8 ? (HW - p0) : pl.NTN * 8;   // valid pixels of this tile
    const uint32_t Ys = (uint32_t)__cvta_generic_to_shared(smem), Xs = Ys + (uint32_t)pl.offX, Hs = Ys + (uint32_t)pl.offH;
    const long ibase = (long)img * C * HW + p0;

    // ---- Y and X tiles -> shared memory (rows = channels, whole chunks of contiguous pixels; chunks past HW are zero)
    {
        const int cpr = pl.NTN * 16 / pl.chunkB;          // chunks per row
        const int epc = pl.chunkB / 2;                    // elements per chunk
        for (int i = tid; i < C * cpr; i += 256) {
            const int row = i / cpr, ch = i - row * cpr;
            const uint32_t so = (uint32_t)row * (uint32_t)PB + (uint32_t)ch * (uint32_t)pl.chunkB;
            if (ch * epc < np) {
                f_cp_async(Ys + so, y + ibase + (long)row * HW + ch * epc, pl.chunkB);
                f_cp_async(Xs + so, x + ibase + (long)row * HW + ch * epc, pl.chunkB);
            } else {
                if (pl.chunkB == 16) { *reinterpret_cast<uint4*>(smem + so) = make_uint4(0, 0, 0, 0); *reinterpret_cast<uint4*>(smem + pl.offX + so) = make_uint4(0, 0, 0, 0); }
                else { *reinterpret_cast<uint2*>(smem + so) = make_uint2(0, 0); *reinterpret_cast<uint2*>(smem + pl.offX + so) = make_uint2(0, 0); }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
    }
    const int nq = pl.NTN;              // a unit = one 16-row m-tile x all n-tiles of the tile (A fragments read once)

    // ---- GEMM 1 + bias + GELU:  Hs[h, p] = gelu(b1[h] + sum_k W1[h, k] Ys[k, p])
    for (int mt = warp; mt < HID / 16; mt += 8) {
        const int nq0 = 0;
        float acc[NQ][4];
#pragma unroll
        for (int q = 0; q < NQ; ++q) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f; }
        ffn_unit<T, NQ>(acc, w1, C, mt * 16, Ys, PB, nq0, lane);
        const float ba = __ldg(b1 + mt * 16 + g), bb = __ldg(b1 + mt * 16 + g + 8);
        const uint32_t ha = Hs + (uint32_t)(mt * 16 + g) * (uint32_t)PB + (uint32_t)(nq0 * 16 + 4 * t), hb = ha + 8u * (uint32_t)PB;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (q < nq) {
                const uint32_t va = FfnT<T>::pack(f_gelu(acc[q][0] + ba), f_gelu(acc[q][1] + ba));
                const uint32_t vb = FfnT<T>::pack(f_gelu(acc[q][2] + bb), f_gelu(acc[q][3] + bb));
                asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(ha + 16u * q), "r"(va) : "memory");
                asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(hb + 16u * q), "r"(vb) : "memory");
            }
    }
    __syncthreads();

    // ---- GEMM 2 + bias + residual:  Xs[c, p] = Xs[c, p] + b2[c] + sum_h W2[c, h] Hs[h, p]
    for (int mt = warp; mt < C / 16; mt += 8) {
        const int nq0 = 0;
        float acc[NQ][4];
#pragma unroll
        for (int q = 0; q < NQ; ++q) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f; }
        ffn_unit<T, NQ>(acc, w2, HID, mt * 16, Hs, PB, nq0, lane);
        const float ba = __ldg(b2 + mt * 16 + g), bb = __ldg(b2 + mt * 16 + g + 8);
        const uint32_t xa = Xs + (uint32_t)(mt * 16 + g) * (uint32_t)PB + (uint32_t)(nq0 * 16 + 4 * t), xb = xa + 8u * (uint32_t)PB;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (q < nq) {
                uint32_t ra, rb;
                asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(ra) : "r"(xa + 16u * q));
                asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(rb) : "r"(xb + 16u * q));
                const float2 fa = FfnT<T>::unpack(ra), fb = FfnT<T>::unpack(rb);
                asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(xa + 16u * q), "r"(FfnT<T>::pack(fa.x + (acc[q][0] + ba), fa.y + (acc[q][1] + ba))) : "memory");
                asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(xb + 16u * q), "r"(FfnT<T>::pack(fb.x + (acc[q][2] + bb), fb.y + (acc[q][3] + bb))) : "memory");
            }
    }
    __syncthreads();

    // ---- result tile -> global memory (coalesced row chunks)
    {
        const int cpr = pl.NTN * 16 / pl.chunkB, epc = pl.chunkB / 2;
        for (int i = tid; i < C * cpr; i += 256) {
            const int row = i / cpr, ch = i - row * cpr;
            if (ch * epc >= np) continue;
            const unsigned char* s = smem + pl.offX + (long)row * PB + ch * pl.chunkB;
            T* d = out + ibase + (long)row * HW + ch * epc;
            if (pl.chunkB == 16) *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
            else *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(s);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Wide stages (C >= 192): the weights no longer sit in L1, and eight warps per SM cannot hide the L2 latency of
// per-fragment loads (ncu: long-scoreboard stalls).  Here the weights of one ROUND (8 m-tiles = 128 rows, one per warp)
// stream through shared memory in K-chunks with a 3-stage cp.async pipeline (full 128-byte row segments, issued
// two chunks ahead) and the A fragments come from ldmatrix.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void f_ldsm4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

constexpr int kFfnStages = 3;

// rows [r0, r0 + 128) of  D = W[rows x K] * S[K x pixels]; warp w computes rows r0 + 16 w .. + 15 (if < rows_total).
// Wb: shared address of kFfnStages chunk buffers of 128 rows x (KC * 2 + 16) bytes.  Ends with every warp past its last
// read of Wb NOT guaranteed: the caller synchronises before the buffers are reused.
template <typename T, int NQ, int KC, int NTHREADS>
__device__ __forceinline__ void ffn_round_staged(float (&acc)[NQ][4], const T* __restrict__ W, int K, int rows_total, int r0, uint32_t S, int PB,
                                                 uint32_t Wb, int tid) {
    constexpr int WP = KC * 2 + 16;                 // chunk row pitch (bytes): odd number of 16-byte pieces
    constexpr int PPR = KC * 2 / 16;                // 16-byte pieces per chunk row
    const int lane = tid & 31, warp = (tid >> 5) & 7, nhalf = tid >> 8;   // 16 warps: m-tile = warp % 8, n-tiles [NQ * nhalf, +NQ)
    const int nch = K / KC;
    // this thread's pieces of a chunk: piece i = tid + j * NTHREADS -> (row, 16-byte piece); loop invariant
    constexpr int NP = (128 * PPR + NTHREADS - 1) / NTHREADS;
    const T* gsrc[NP];
    uint32_t sdst[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const int i = tid + j * NTHREADS, row = i / PPR, pc = i - row * PPR;
        const bool ok = i < 128 * PPR && r0 + row < rows_total;
        gsrc[j] = ok ? W + (long)(r0 + row) * K + pc * 8 : nullptr;
        sdst[j] = Wb + (uint32_t)(row * WP + pc * 16);
    }
    auto issue = [&](int c) {
        const uint32_t boff = (uint32_t)(c % kFfnStages) * (uint32_t)(128 * WP);
#pragma unroll
        for (int j = 0; j < NP; ++j)
            if (gsrc[j]) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sdst[j] + boff), "l"(gsrc[j] + c * KC) : "memory");
    };
#pragma unroll
    for (int s = 0; s < kFfnStages - 1; ++s) {
        if (s < nch) issue(s);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    const bool active = r0 + warp * 16 < rows_total;
    const uint32_t arow = (uint32_t)((warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * WP) + (uint32_t)(lane >> 4) * 16u;
    const uint32_t sb = S + (uint32_t)(lane & 15) * (uint32_t)PB + (uint32_t)(lane >> 4) * 16u + (uint32_t)(nhalf * NQ) * 16u;
    for (int c = 0; c < nch; ++c) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kFfnStages - 2) : "memory");
        __syncthreads();
        if (c + kFfnStages - 1 < nch) issue(c + kFfnStages - 1);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        if (active) {
            // fragments of k-step ks + 1 are fetched before the MMAs of k-step ks are issued (the asm statements keep program
            // order, so the overlap of ldmatrix latency with the tensor pipe has to be written out)
            const uint32_t wbuf = Wb + (uint32_t)(c % kFfnStages) * (uint32_t)(128 * WP) + arow;
            const uint32_t srow0 = sb + (uint32_t)((c * KC) * PB);
            uint32_t A[2][4], B[2][NQ / 2][4];
            f_ldsm4(A[0][0], A[0][1], A[0][2], A[0][3], wbuf);
#pragma unroll
            for (int q = 0; q < NQ; q += 2) f_ldsm4t(B[0][q / 2][0], B[0][q / 2][1], B[0][q / 2][2], B[0][q / 2][3], srow0 + (uint32_t)q * 16u);
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks) {
                const int cur = ks & 1, nxt = cur ^ 1;
                if (ks + 1 < KC / 16) {
                    f_ldsm4(A[nxt][0], A[nxt][1], A[nxt][2], A[nxt][3], wbuf + (uint32_t)(ks + 1) * 32u);
                    const uint32_t srow = srow0 + (uint32_t)(((ks + 1) * 16) * PB);
#pragma unroll
                    for (int q = 0; q < NQ; q += 2) f_ldsm4t(B[nxt][q / 2][0], B[nxt][q / 2][1], B[nxt][q / 2][2], B[nxt][q / 2][3], srow + (uint32_t)q * 16u);
                }
#pragma unroll
                for (int q = 0; q < NQ; q += 2) {
                    FfnT<T>::mma(acc[q], A[cur][0], A[cur][1], A[cur][2], A[cur][3], B[cur][q / 2][0], B[cur][q / 2][1]);
                    if (q + 1 < NQ) FfnT<T>::mma(acc[q + 1], A[cur][0], A[cur][1], A[cur][2], A[cur][3], B[cur][q / 2][2], B[cur][q / 2][3]);
                }
            }
        }
    }
}

template <typename T, int KC1, int KC2>
__global__ void __launch_bounds__(512) recnext_ffn_staged_kernel(const __grid_constant__ FfnPlan pl, const T* __restrict__ y, const T* __restrict__ x,
                                                                 const T* __restrict__ w1, const float* __restrict__ b1, const T* __restrict__ w2,
                                                                 const float* __restrict__ b2, T* __restrict__ out) {
    constexpr int NQ = 4;          // n-tiles per warp: 16 warps = 8 m-tiles x 2 halves of the 8-wide pixel tile
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = (tid >> 5) & 7, nq0 = (tid >> 8) * NQ;
    const int g = lane >> 2, t = lane & 3;
    const int img = blockIdx.x / pl.tiles, tile = blockIdx.x - img * pl.tiles;
    const int C = pl.C, HID = pl.HID, HW = pl.HW, PB = pl.PB;
    const int p0 = tile * pl.NTN * 8;
    const int np = (HW - p0) < pl.NTN * 8 ? (HW - p0) : pl.NTN * 8;
    const uint32_t Ys = (uint32_t)__cvta_generic_to_shared(smem), Xs = Ys + (uint32_t)pl.offX, Hs = Ys + (uint32_t)pl.offH, Wb = Ys + (uint32_t)pl.offW;
    const long ibase = (long)img * C * HW + p0;
    {
        const int cpr = pl.NTN * 16 / pl.chunkB, epc = pl.chunkB / 2;
        for (int i = tid; i < C * cpr; i += 512) {
            const int row = i / cpr, ch = i - row * cpr;
            const uint32_t so = (uint32_t)row * (uint32_t)PB + (uint32_t)ch * (uint32_t)pl.chunkB;
            if (ch * epc < np) {
                f_cp_async(Ys + so, y + ibase + (long)row * HW + ch * epc, pl.chunkB);
                f_cp_async(Xs + so, x + ibase + (long)row * HW + ch * epc, pl.chunkB);
            } else {
                if (pl.chunkB == 16) { *reinterpret_cast<uint4*>(smem + so) = make_uint4(0, 0, 0, 0); *reinterpret_cast<uint4*>(smem + pl.offX + so) = make_uint4(0, 0, 0, 0); }
                else { *reinterpret_cast<uint2*>(smem + so) = make_uint2(0, 0); *reinterpret_cast<uint2*>(smem + pl.offX + so) = make_uint2(0, 0); }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
    }
    const int nq = pl.NTN;
    // ---- GEMM 1 + bias + GELU
    for (int r0 = 0; r0 < HID; r0 += 128) {
        float acc[NQ][4];
#pragma unroll
        for (int q = 0; q < NQ; ++q) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f; }
        ffn_round_staged<T, NQ, KC1, 512>(acc, w1, C, HID, r0, Ys, PB, Wb, tid);
        const int m0 = r0 + warp * 16;
        if (m0 < HID) {
            const float ba = __ldg(b1 + m0 + g), bb = __ldg(b1 + m0 + g + 8);
            const uint32_t ha = Hs + (uint32_t)(m0 + g) * (uint32_t)PB + (uint32_t)(16 * nq0 + 4 * t), hb = ha + 8u * (uint32_t)PB;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (nq0 + q < nq) {
                    const uint32_t va = FfnT<T>::pack(f_gelu(acc[q][0] + ba), f_gelu(acc[q][1] + ba));
                    const uint32_t vb = FfnT<T>::pack(f_gelu(acc[q][2] + bb), f_gelu(acc[q][3] + bb));
                    asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(ha + 16u * q), "r"(va) : "memory");
                    asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(hb + 16u * q), "r"(vb) : "memory");
                }
        }
        __syncthreads();   // weight buffers are reused by the next round; Hs complete before GEMM 2
    }
    // ---- GEMM 2 + bias + residual
    for (int r0 = 0; r0 < C; r0 += 128) {
        float acc[NQ][4];
#pragma unroll
        for (int q = 0; q < NQ; ++q) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f; }
        ffn_round_staged<T, NQ, KC2, 512>(acc, w2, HID, C, r0, Hs, PB, Wb, tid);
        const int m0 = r0 + warp * 16;
        if (m0 < C) {
            const float ba = __ldg(b2 + m0 + g), bb = __ldg(b2 + m0 + g + 8);
            const uint32_t xa = Xs + (uint32_t)(m0 + g) * (uint32_t)PB + (uint32_t)(16 * nq0 + 4 * t), xb = xa + 8u * (uint32_t)PB;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (nq0 + q < nq) {
                    uint32_t ra, rb;
                    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(ra) : "r"(xa + 16u * q));
                    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(rb) : "r"(xb + 16u * q));
                    const float2 fa = FfnT<T>::unpack(ra), fb = FfnT<T>::unpack(rb);
                    asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(xa + 16u * q), "r"(FfnT<T>::pack(fa.x + (acc[q][0] + ba), fa.y + (acc[q][1] + ba))) : "memory");
                    asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(xb + 16u * q), "r"(FfnT<T>::pack(fb.x + (acc[q][2] + bb), fb.y + (acc[q][3] + bb))) : "memory");
                }
        }
        __syncthreads();
    }
    {
        const int cpr = pl.NTN * 16 / pl.chunkB, epc = pl.chunkB / 2;
        for (int i = tid; i < C * cpr; i += 512) {
            const int row = i / cpr, ch = i - row * cpr;
            if (ch * epc >= np) continue;
            const unsigned char* s = smem + pl.offX + (long)row * PB + ch * pl.chunkB;
            T* d = out + ibase + (long)row * HW + ch * epc;
            if (pl.chunkB == 16) *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
            else *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(s);
        }
    }
}

// 0 ok; 1 unsupported shape (caller keeps the library path); fills pl
int ffn_make_plan(FfnPlan& pl, int B, int C, int HID, int HW, int dtype) {
    if (B < 1 || C < 16 || HID < 16 || HW < 1) return 1;
    if ((C % 16) != 0 || (HID % 16) != 0 || !(dtype == 1 || dtype == 2)) return 1;
    if ((HW % 4) != 0) return 1;                       // whole 8-byte chunks of a pixel row
    pl = FfnPlan();
    pl.B = B; pl.C = C; pl.HID = HID; pl.HW = HW; pl.dtype = dtype;
    pl.chunkB = (HW % 8) == 0 ? 16 : 8;
    const int nt = (HW + 7) / 8;                      // n-tiles per image
    // 8 n-tiles (64 pixels) per CTA tile; 4 when the wide-stage tiles would otherwise leave one CTA (8 warps) per SM
    // (a 4-wide variant exists — two CTAs per SM for the wide stages — but measured slower, 0.383 vs 0.341 ms at [256, 14x14]:
    // every CTA re-reads all weights from L2, which is what bounds those stages; RECNEXT_FFN_NQ=4 selects it for experiments)
    int nq = 8;
    { const char* e = getenv("RECNEXT_FFN_NQ"); if (e && atoi(e) == 4) nq = 4; }
    pl.NQ = nq;
    const int tiles = (nt + nq - 1) / nq;
    pl.tiles = tiles;
    pl.NTN = (nt + tiles - 1) / tiles;                // <= nq, balanced over the tiles
    pl.PB = 16 * (nq + 1);                            // units sweep nq n-tiles (chunks past NTN are ignored); odd pitch: conflict-free ldmatrix
    pl.offX = C * pl.PB;
    pl.offH = 2 * C * pl.PB;
    pl.smem_bytes = (2 * C + HID) * pl.PB + 64;       // + slack: a paired ldmatrix may over-read one chunk past the last row
    if (pl.smem_bytes > 227 * 1024) return 1;
    // staged weights: C >= 192 (weights beyond L1), 8-wide tiles, K multiples of 32; RECNEXT_FFN_STAGE=0|1 overrides
    pl.staged = (C >= 192 && nq == 8 && (C % 32) == 0 && (HID % 32) == 0) ? 1 : 0;
    { const char* e = getenv("RECNEXT_FFN_STAGE"); if (e) pl.staged = (atoi(e) != 0 && nq == 8 && (C % 32) == 0 && (HID % 32) == 0) ? 1 : 0; }
    if (pl.staged) {
        pl.offW = (2 * C + HID) * pl.PB + 64;
        int kc = ((C % 64) == 0 && (HID % 64) == 0) ? 64 : 32;
        if (pl.offW + 3 * 128 * (kc * 2 + 16) > 227 * 1024) kc = 32;
        const int bytes = pl.offW + 3 * 128 * (kc * 2 + 16);
        pl.kc = kc;
        if (bytes > 227 * 1024) pl.staged = 0; else pl.smem_bytes = bytes;
    }
    return 0;
}

cudaError_t ffn_launch(const FfnPlan& pl, const void* y, const void* x, const void* w1, const float* b1, const void* w2, const float* b2, void* out,
                       cudaStream_t stream) {
    static DeviceOnce configured = {};
    {
        const cudaError_t e0 = rc_once_per_device(configured, [] {
            cudaError_t e = cudaFuncSetAttribute(recnext_ffn_kernel<__nv_bfloat16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(recnext_ffn_kernel<__nv_bfloat16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(recnext_ffn_kernel<__half, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(recnext_ffn_kernel<__half, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            return e;
        });
        if (e0 != cudaSuccess) return e0;
    }
    const int grid = pl.B * pl.tiles;
    if (pl.staged) {
        static DeviceOnce configured_s = {};
        const cudaError_t e1 = rc_once_per_device(configured_s, [] {
            cudaError_t e = cudaFuncSetAttribute(recnext_ffn_staged_kernel<__nv_bfloat16, 64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(recnext_ffn_staged_kernel<__nv_bfloat16, 32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(recnext_ffn_staged_kernel<__half, 64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(recnext_ffn_staged_kernel<__half, 32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            return e;
        });
        if (e1 != cudaSuccess) return e1;
        const bool k64 = pl.kc == 64;
#define FFN_LAUNCH_S(TT, KC) recnext_ffn_staged_kernel<TT, KC, KC><<<grid, 512, pl.smem_bytes, stream>>>(pl, (const TT*)y, (const TT*)x, (const TT*)w1, b1, (const TT*)w2, b2, (TT*)out)
        if (pl.dtype == 1) { if (k64) FFN_LAUNCH_S(__nv_bfloat16, 64); else FFN_LAUNCH_S(__nv_bfloat16, 32); }
        else { if (k64) FFN_LAUNCH_S(__half, 64); else FFN_LAUNCH_S(__half, 32); }
#undef FFN_LAUNCH_S
        return cudaGetLastError();
    }
#define FFN_LAUNCH(TT, NQV) recnext_ffn_kernel<TT, NQV><<<grid, 256, pl.smem_bytes, stream>>>(pl, (const TT*)y, (const TT*)x, (const TT*)w1, b1, (const TT*)w2, b2, (TT*)out)
    if (pl.dtype == 1) { if (pl.NQ == 8) FFN_LAUNCH(__nv_bfloat16, 8); else FFN_LAUNCH(__nv_bfloat16, 4); }
    else { if (pl.NQ == 8) FFN_LAUNCH(__half, 8); else FFN_LAUNCH(__half, 4); }
#undef FFN_LAUNCH
    return cudaGetLastError();
}

}  // namespace recnext
