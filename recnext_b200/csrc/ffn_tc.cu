// ffn_tc.cu — fused channel mixer of a RecNeXt block on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulators,
// TMA bulk copies), NCHW tensors, 16-bit activations.  Replaces
//     x + mlp(norm(y))      mlp = 1x1 conv -> GELU -> 1x1 conv, ConvNorms / eval BatchNorm folded     model/recnext.py:125-131,153,157-158
// with ONE persistent kernel:   out[b,c,p] = x[b,c,p] + b2[c] + sum_h W2[c,h] gelu(b1[h] + sum_k W1[h,k] y[b,k,p]).
//
// In NCHW one image is a row-major [C x HW] matrix, so with the pixels of the whole batch flattened (P = B HW) both 1x1
// convs are GEMMs whose activation operand is "K x N with N (pixels) contiguous": the MN-major B operand of tcgen05.mma.
// A CTA owns a tile of NT pixels (NT = 128, or 64 for C > 256) and does, per 128-row chunk `hc` of the hidden layer,
//     GEMM1  D1[128 x NT]  = W1[hc] (128 x C)      . Y (C x NT)          accumulator in TMEM (double buffered)
//     EPI1   H[hc]         = gelu(D1 + b1)  -> 16-bit, shared memory, already in the MN-major layout GEMM2 reads
//     GEMM2  D2[C x NT]   += W2[:, hc] (C x 128)   . H[hc] (128 x NT)    accumulators in TMEM for the whole tile
// and finally EPI2 out = D2 + b2 + x.  The hidden activation never leaves the SM; HBM traffic is 3 N e.
//
// Warp roles (448 threads, one CTA per SM):
//   warps 0-7   epilogue: TMEM -> registers (tcgen05.ld) -> bias / GELU / residual -> shared memory or global memory
//   warp  8     weight producer: one elected lane streams the PRE-PACKED weight tiles (16 KB, exactly the shared-memory image
//               of a 128 x 64 K-major operand) through a 4-slot ring with TMA bulk copies (cp.async.bulk + mbarrier tx counts)
//   warp  9     MMA issuer: one elected lane issues every tcgen05.mma; tcgen05.commit releases ring slots / publishes accumulators
//   warps 10-13 activation loaders: Y tile global -> shared memory in the MN-major core-matrix layout (16-byte chunks)
// All hand-offs are mbarriers; the MMA stream is software pipelined (GEMM1 of chunk s + 1 is issued before GEMM2 of chunk s) so the
// tensor pipe works while the epilogue warps run the GELU of chunk s.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>
#include "tc05.cuh"
#include "devcfg.h"
#include "ffn_tc.h"

namespace recnext {

namespace {

constexpr int kRing = 8;             // weight ring slots at most (the plan takes as many as fit: 4 .. 8)
constexpr int kTileBytes = 16384;    // 128 rows x 64 K x 2 bytes
constexpr int kEpiWarps = 8, kLoadWarps = 4;
constexpr int kThreads = 32 * (kEpiWarps + 2 + kLoadWarps);
constexpr uint32_t kSboH = 128 * 16 + 16;   // H: one 8-pixel chunk of all 128 hidden rows + 16 bytes (conflict-free chunk stores)

template <typename T> struct Cvt;
template <> struct Cvt<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
    static __device__ __forceinline__ float2 unpack(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
    static constexpr int fmt = 1;
};
template <> struct Cvt<__half> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __half2 v = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
    static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
    static constexpr int fmt = 0;
};

// gelu(t) = t Phi(t) for TWO elements.  Phi through the hardware tanh (ONE MUFU per element: the exact-erf form needs two and the
// SFU pipe is scarce here): Phi(t) ~ 0.5 (1 + tanh(t (a + b t^2))) with (a, b) fitted to the erf form: max |gelu - gelu_erf| = 2.7e-4
// over the reals, below the 16-bit rounding of the hidden activation that follows (the reference rounds it to bf16 under autocast:
// relative 2^-9).  The arithmetic uses the packed fp32x2 instructions of sm_100 (FADD2 / FMUL2 / FFMA2): the epilogue warps are bound
// by instruction issue, and a packed instruction does two elements per issue slot.
__device__ __forceinline__ float2 gelu2(float2 v, float2 bias) {
    const float2 t = __fadd2_rn(v, bias);
    const float2 t2 = __fmul2_rn(t, t);
    const float2 q = __ffma2_rn(t2, make_float2(0.03470089f, 0.03470089f), make_float2(0.80015708f, 0.80015708f));
    const float2 u = __fmul2_rn(t, q);
    float2 th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(u.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(u.y));
    const float2 h = __fmul2_rn(t, make_float2(0.5f, 0.5f));
    return __ffma2_rn(h, th, h);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
    return e != 0;
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Ampere-style asynchronous copies (LDGSTS): no register staging, so a loader thread keeps ALL its loads of a tile in flight.
// src_size 0 zero-fills the destination (pixels past the batch, padded channels).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Position of a pixel of the flattened batch: image b, pixel i inside the image.  Tiles walk forward through the pixels, so the
// (slow) integer division happens once per tile and thread; everything else is increments.
struct Pix {
    int b, i;
};
__device__ __forceinline__ Pix pix_make(uint32_t p, uint32_t HW) {
    Pix r;
    r.b = (int)(p / HW);
    r.i = (int)(p - (uint32_t)r.b * HW);
    return r;
}
__device__ __forceinline__ Pix pix_add(Pix a, int n, int HW) {
    a.i += n;
    while (a.i >= HW) { a.i -= HW; ++a.b; }
    return a;
}

// 8 consecutive pixels starting at `px` of channel row `ch`; pixels of images >= B read as zero / are not written.
// VEC = 8: HW % 8 == 0, one 16-byte access; VEC = 4: HW % 4 == 0, two 8-byte accesses; VEC = 1: element-wise.
template <typename T, int VEC>
__device__ __forceinline__ uint4 load8(const T* __restrict__ base, Pix px, int B, int C, int HW, int ch) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (VEC == 8) {
        if (px.b < B) v = *reinterpret_cast<const uint4*>(base + (((long)px.b * C + ch) * (long)HW + px.i));
    } else if (VEC == 4) {
        uint2 h[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (px.b < B) h[u] = *reinterpret_cast<const uint2*>(base + (((long)px.b * C + ch) * (long)HW + px.i));
            px = pix_add(px, 4, HW);
        }
        v = make_uint4(h[0].x, h[0].y, h[1].x, h[1].y);
    } else {
        unsigned short e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            e[u] = 0;
            if (px.b < B) e[u] = *reinterpret_cast<const unsigned short*>(base + (((long)px.b * C + ch) * (long)HW + px.i));
            px = pix_add(px, 1, HW);
        }
        v = make_uint4(e[0] | ((uint32_t)e[1] << 16), e[2] | ((uint32_t)e[3] << 16), e[4] | ((uint32_t)e[5] << 16), e[6] | ((uint32_t)e[7] << 16));
    }
    return v;
}
template <typename T, int VEC>
__device__ __forceinline__ void store8(T* __restrict__ base, Pix px, int B, int C, int HW, int ch, uint4 v) {
    if (VEC == 8) {
        if (px.b < B) *reinterpret_cast<uint4*>(base + (((long)px.b * C + ch) * (long)HW + px.i)) = v;
    } else if (VEC == 4) {
        const uint2 h[2] = {make_uint2(v.x, v.y), make_uint2(v.z, v.w)};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (px.b < B) *reinterpret_cast<uint2*>(base + (((long)px.b * C + ch) * (long)HW + px.i)) = h[u];
            px = pix_add(px, 4, HW);
        }
    } else {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (px.b < B)
                *reinterpret_cast<unsigned short*>(base + (((long)px.b * C + ch) * (long)HW + px.i)) = (unsigned short)((w[u >> 1] >> (16 * (u & 1))) & 0xffffu);
            px = pix_add(px, 1, HW);
        }
    }
}

constexpr int kMaxY = 4;   // activation tile buffers at most
enum Bar { W_FULL = 0, W_EMPTY = kRing, Y_FULL = 2 * kRing, Y_EMPTY = Y_FULL + kMaxY, D1_FULL = Y_EMPTY + kMaxY, H_FULL = D1_FULL + 2,
           H_EMPTY = H_FULL + 2, D2_FULL = H_EMPTY + 2, D2_EMPTY = D2_FULL + 2, NUM_BARS = D2_EMPTY + 2 };

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA bulk copy global -> shared memory of EVERY CTA in `mask` (same offset in each), completion on the mbarrier at the same offset in each
__device__ __forceinline__ void bulk_g2s_mcast(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst_saddr), "l"(src),
                 "r"(bytes), "r"(bar), "h"(mask)
                 : "memory");
}
// all tcgen05.mma issued so far by this thread arrive, when they complete, on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_mcast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// arrive on an mbarrier once every cp.async issued so far by this thread has landed (the arrival is part of the barrier's count)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }

template <typename T, int NT, int VEC>
__global__ void __launch_bounds__(kThreads, 1) recnext_ffn_tc_kernel(const __grid_constant__ FfnTcPlan p, const T* __restrict__ gy, const T* __restrict__ gx,
                                                                     const uint8_t* __restrict__ wpk, const float* __restrict__ b1,
                                                                     const float* __restrict__ b2, T* __restrict__ gout, long long* __restrict__ prof) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const uint32_t sb = tc::smem_u32(smem);
    const uint32_t bars = sb + p.offBar;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.offBar + 8 * NUM_BARS);
    float* sb1 = reinterpret_cast<float*>(smem + p.offBias);   // b1 padded to HIDP, then b2 padded to nCT * 128
    float* sb2 = sb1 + p.HIDP;
    const int CS = p.cs;                       // CTAs per cluster sharing one weight stream (TMA multicast)
    const uint32_t crank = CS > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << CS) - 1u);

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.RS; ++i) { tc::mbar_init(bar(W_FULL + i), 1); tc::mbar_init(bar(W_EMPTY + i), (uint32_t)CS); }
        for (int i = 0; i < kMaxY; ++i) { tc::mbar_init(bar(Y_FULL + i), kLoadWarps * 32); tc::mbar_init(bar(Y_EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(bar(D1_FULL + i), 1); tc::mbar_init(bar(H_FULL + i), 4); tc::mbar_init(bar(H_EMPTY + i), 1);
            tc::mbar_init(bar(D2_FULL + i), 1); tc::mbar_init(bar(D2_EMPTY + i), 4);
        }
        tc::mbar_init_fence();
    }
    for (int i = threadIdx.x; i < p.HIDP; i += kThreads) sb1[i] = i < p.HID ? b1[i] : 0.f;
    for (int i = threadIdx.x; i < p.nCT * 128; i += kThreads) sb2[i] = i < p.C ? b2[i] : 0.f;
    if (warp == 9) tc::tmem_alloc(tc::smem_u32(tmem_slot), 512);
    tc::fence_before_sync();
    __syncthreads();
    if (CS > 1) cluster_sync_all();   // every CTA's barriers exist before a peer's multicast copy or commit can reach them
    tc::fence_after_sync();
    const uint32_t tbase = *tmem_slot;

    const int nH = p.nH, nK1 = p.nK1, nCT = p.nCT, C = p.C, HW = p.HW;
    // RECNEXT_FFN_PROF: CTA 0 records clock64() at the hand-off points of the first chunks (timing experiments only)
    auto stamp = [&](int role, int idx) {
        if (prof != nullptr && blockIdx.x == 0 && lane == 0 && idx < 500) prof[role * 512 + idx] = clock64();
    };
    // Tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ... — the SAME count for every CTA (tiles past the end read zeros and
    // write nothing), because the CTAs of a cluster consume one multicast weight stream in lock step.  The hidden chunks of all
    // tiles form ONE stream of G chunks that the MMA and epilogue warps walk, software pipelined across tile borders.
    const int my_tiles = p.tiles_per_cta;
    const int G = my_tiles * nH;
    const int RS = p.RS, nY = p.nY, nD2 = p.nD2;
    const uint32_t w1_step = (uint32_t)(128 * 2) * (uint32_t)p.CP, w2_step = (uint32_t)(2 * nCT) * kTileBytes;

    if (warp == 8) {
        // ================= weight producer =================
        // stream order = the MMA warp's: W1 tiles of chunk g, then the W2 tiles of chunk g - 1.  In a cluster every CTA fetches
        // 1/CS of each tile and multicasts it to all of them: L2 is read once per cluster.
        uint32_t use = 0;
        const bool leader = elect_one();
        const uint8_t* w2base = wpk + (size_t)nH * w1_step;
        int s = 0;
        for (int g = 0; g <= G; ++g) {
            const int s_prev = s == 0 ? nH - 1 : s - 1;
            const int n1 = g < G ? nK1 : 0, n2 = g >= 1 ? 2 * nCT : 0;
            const uint8_t* src1 = wpk + (size_t)s * w1_step;
            const uint8_t* src2 = w2base + (size_t)s_prev * w2_step;
            for (int t = 0; t < n1 + n2; ++t) {
                const uint32_t bytes = t < n1 ? (uint32_t)(128 * 2 * (t == nK1 - 1 ? p.kwLast : 64)) : (uint32_t)kTileBytes;
                const uint8_t* src = t < n1 ? src1 + (size_t)t * kTileBytes : src2 + (size_t)(t - n1) * kTileBytes;
                const uint32_t slot = use % RS;
                tc::mbar_wait(bar(W_EMPTY + slot), ((use / RS) & 1u) ^ 1u);
                if (leader) {
                    if (p.dbg & 1) tc::mbar_arrive(bar(W_FULL + slot));   // timing experiment: no weight traffic (results are garbage)
                    else {
                        tc::mbar_expect_tx(bar(W_FULL + slot), bytes);
                        if (CS == 1) tc::bulk_g2s(sb + p.offW + slot * kTileBytes, src, bytes, bar(W_FULL + slot));
                        else {
                            const uint32_t part = bytes / (uint32_t)CS;
                            bulk_g2s_mcast(sb + p.offW + slot * kTileBytes + crank * part, src + crank * part, part, bar(W_FULL + slot), cmask);
                        }
                    }
                }
                ++use;
            }
            if (++s == nH) s = 0;
        }
    } else if (warp == 9) {
        // ================= MMA issuer =================
        const bool leader = elect_one();
        const uint32_t idesc = tc::make_idesc(128, NT, Cvt<T>::fmt, 0, 1);
        auto release_slot = [&](uint32_t slot) {
            if (CS == 1) tc::mma_commit(bar(W_EMPTY + slot));
            else mma_commit_mcast(bar(W_EMPTY + slot), cmask);
        };
        uint32_t wuse = 0;
        int s = 0, t = 0;            // chunk g = (tile t of this CTA, hidden chunk s)
        int s2 = 0, t2 = 0;          // chunk g - 1
        for (int g = 0; g <= G; ++g) {
            if (g < G) {   // GEMM1 of chunk g:  D1[g & 1] = W1[s] . Y[t]
                const uint32_t ybuf = (uint32_t)t % nY;
                if (s == 0) {
                    tc::mbar_wait(bar(Y_FULL + ybuf), ((uint32_t)t / nY) & 1u);
                    tc::fence_proxy_async();   // the loaders' cp.async / st.shared writes -> the tensor core's async-proxy reads
                    tc::fence_after_sync();
                }
                const uint32_t ybase = sb + p.offY + ybuf * p.yBytes;
                const uint32_t buf = (uint32_t)g & 1u;
                const uint32_t d1 = tbase + buf * NT;
                stamp(0, 4 * g + 0);
                for (int kc = 0; kc < nK1; ++kc) {
                    const uint32_t slot = wuse % RS;
                    tc::mbar_wait(bar(W_FULL + slot), (wuse / RS) & 1u);
                    tc::fence_after_sync();
                    const int ksteps = (kc == nK1 - 1 ? p.kwLast : 64) / 16;
                    const uint64_t ad = tc::make_sdesc(sb + p.offW + slot * kTileBytes, 2048, 128);
                    const uint64_t bd = tc::make_sdesc(ybase + (uint32_t)(kc * 64) * 16u, 128, p.sboY);
                    for (int j = 0; j < ksteps; ++j)
                        if (leader) tc::mma_ss(d1, tc::sdesc_advance(ad, (uint32_t)j * 4096u), tc::sdesc_advance(bd, (uint32_t)j * 256u), idesc, (kc | j) != 0);
                    if (leader) release_slot(slot);
                    ++wuse;
                }
                if (leader) {
                    tc::mma_commit(bar(D1_FULL + buf));
                    if (s == nH - 1) tc::mma_commit(bar(Y_EMPTY + ybuf));   // the loaders may fetch this buffer's next tile
                }
                stamp(0, 4 * g + 1);
                if (++s == nH) { s = 0; ++t; }
            }
            if (g >= 1) {   // GEMM2 of chunk g - 1:  D2[t2] += W2[:, s2] . H[(g - 1) & 1]
                const uint32_t gg = (uint32_t)(g - 1), buf = gg & 1u;
                const uint32_t db = (uint32_t)t2 % nD2;
                tc::mbar_wait(bar(H_FULL + buf), (gg >> 1) & 1u);
                if (s2 == 0) tc::mbar_wait(bar(D2_EMPTY + db), (((uint32_t)t2 / nD2) & 1u) ^ 1u);   // that accumulator's previous tile has left TMEM
                tc::fence_after_sync();
                stamp(0, 4 * (g - 1) + 2);
                const uint32_t hbase = sb + p.offH + buf * p.hBytes;
                for (int ct = 0; ct < nCT; ++ct) {
                    const uint32_t d2 = tbase + 2 * NT + (db * nCT + ct) * NT;
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t slot = wuse % RS;
                        tc::mbar_wait(bar(W_FULL + slot), (wuse / RS) & 1u);
                        tc::fence_after_sync();
                        const uint64_t ad = tc::make_sdesc(sb + p.offW + slot * kTileBytes, 2048, 128);
                        const uint64_t bd = tc::make_sdesc(hbase + (uint32_t)(half * 64) * 16u, 128, kSboH);
                        for (int j = 0; j < 4; ++j)
                            if (leader) tc::mma_ss(d2, tc::sdesc_advance(ad, (uint32_t)j * 4096u), tc::sdesc_advance(bd, (uint32_t)j * 256u), idesc, !(s2 == 0 && half == 0 && j == 0));
                        if (leader) release_slot(slot);
                        ++wuse;
                    }
                }
                if (leader) {
                    tc::mma_commit(bar(H_EMPTY + buf));
                    if (s2 == nH - 1) tc::mma_commit(bar(D2_FULL + db));
                }
                stamp(0, 4 * (g - 1) + 3);
                if (++s2 == nH) { s2 = 0; ++t2; }
            }
        }
    } else if (warp >= 10) {
        // ================= activation loaders =================
        // Fully asynchronous: a thread issues the cp.async copies of its part of a tile and attaches an mbarrier arrival to them
        // (no wait), then moves on to the next tile as soon as that buffer is free: up to nY tiles are in flight.
        constexpr int CH = NT / 8;                       // 8-pixel chunks per channel row of a tile
        constexpr int KPP = (kLoadWarps * 32) / CH;      // channels per pass
        const int lt = threadIdx.x - 320;
        const int n = lt % CH, k0 = lt / CH;
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const uint32_t ybuf = (uint32_t)t % nY;
            tc::mbar_wait(bar(Y_EMPTY + ybuf), (((uint32_t)t / nY) & 1u) ^ 1u);
            if (warp == 10) stamp(3, 2 * t);
            const uint32_t dst = sb + p.offY + ybuf * p.yBytes + (uint32_t)n * p.sboY;
            const Pix px = pix_make((uint32_t)tile * NT + 8u * n, (uint32_t)HW);
            if (VEC >= 4) {
                const Pix px4 = pix_add(px, 4, HW);
                const T* src0 = gy + (((long)px.b * C) * (long)HW + px.i);
                const T* src1 = gy + (((long)px4.b * C) * (long)HW + px4.i);
                const bool in0 = px.b < p.B && !(p.dbg & 8), in1 = px4.b < p.B && !(p.dbg & 8);
                for (int k = k0; k < p.CP; k += KPP) {
                    const bool kin = k < C;
                    if (VEC == 8) cp_async16(dst + (uint32_t)k * 16u, in0 && kin ? (const void*)(src0 + (long)k * HW) : (const void*)gy, in0 && kin);
                    else {
                        cp_async8(dst + (uint32_t)k * 16u, in0 && kin ? (const void*)(src0 + (long)k * HW) : (const void*)gy, in0 && kin);
                        cp_async8(dst + (uint32_t)k * 16u + 8u, in1 && kin ? (const void*)(src1 + (long)k * HW) : (const void*)gy, in1 && kin);
                    }
                }
                cp_async_arrive_noinc(bar(Y_FULL + ybuf));
            } else {
                for (int kb = k0; kb < p.CP; kb += 4 * KPP) {
                    uint4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = kb + u * KPP;
                        v[u] = (k < C) ? load8<T, VEC>(gy, px, p.B, C, HW, k) : make_uint4(0u, 0u, 0u, 0u);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = kb + u * KPP;
                        if (k < p.CP) sts128(dst + (uint32_t)k * 16u, v[u].x, v[u].y, v[u].z, v[u].w);
                    }
                }
                tc::mbar_arrive(bar(Y_FULL + ybuf));
            }
            if (warp == 10) stamp(3, 2 * t + 1);
        }
        if (VEC >= 4) cp_async_wait_all();
    } else {
        // ================= epilogue warps: two groups of four, chunk g belongs to group g & 1 =================
        // A group owns D1[grp] / H[grp]: while one group runs the GELU of chunk g, the other waits for (or already works on) chunk g + 1,
        // so barrier and memory latencies of the two chains overlap.  A thread owns one accumulator row (TMEM lane) and all NT columns.
        const int q = warp & 3, grp = warp >> 2;
        const int row = 32 * q + lane;
        const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
        constexpr int NCH = NT / 8;
        const uint32_t hdst = sb + p.offH + (uint32_t)grp * p.hBytes + (uint32_t)row * 16u;
        int s = grp % nH, t = grp / nH;   // chunk g = grp, grp + 2, ...
        for (int g = grp; g < G; g += 2) {
            if (q == 0) stamp(1 + grp, 6 * (g >> 1) + 0);
            // ---- H[grp] = gelu(D1[grp] + b1[s])
            const uint32_t use = (uint32_t)g >> 1;
            const float bias = sb1[s * 128 + row];
            const float2 bias2v = make_float2(bias, bias);
            tc::mbar_wait(bar(H_EMPTY + grp), (use & 1u) ^ 1u);   // GEMM2 of chunk g - 2 has read this H buffer
            tc::mbar_wait(bar(D1_FULL + grp), use & 1u);
            tc::fence_after_sync();
            if (q == 0) stamp(1 + grp, 6 * (g >> 1) + 1);
#pragma unroll
            for (int c0 = 0; c0 < NT; c0 += 32) {
                uint32_t v[32];
                tc::tmem_ld32(trow + grp * NT + c0, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 r = gelu2(make_float2(__uint_as_float(v[8 * ch + 2 * e]), __uint_as_float(v[8 * ch + 2 * e + 1])), bias2v);
                        w[e] = Cvt<T>::pack(r.x, r.y);
                    }
                    sts128(hdst + (uint32_t)(c0 / 8 + ch) * kSboH, w[0], w[1], w[2], w[3]);
                }
            }
            tc::fence_proxy_async();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bar(H_FULL + grp));
            if (q == 0) stamp(1 + grp, 6 * (g >> 1) + 2);
            // ---- last chunk of a tile: out = D2 + b2 + x once its GEMM2 has finished (the other group works meanwhile).
            // The accumulator rows (one channel per lane) are staged in shared memory (this group's H buffer is free by then) and
            // leave through a second pass whose lanes run along the pixels: global loads / stores of x and out are coalesced
            // (a row-per-lane access costs 32 L1 wavefronts per instruction and was the bottleneck of the HBM-bound stages).
            if (s == nH - 1) {
                constexpr int CH = NT / 8, RPI = 32 / CH;      // 8-pixel chunks per row; rows per warp pass of the coalesced phase
                constexpr uint32_t SP = NT * 2 + 16;           // staging row pitch (bytes): conflict-free for both access patterns
                const int tile = (int)blockIdx.x + t * (int)gridDim.x;
                const uint32_t db = (uint32_t)t % nD2;
                const bool traffic = !(p.dbg & 4);
                const int n = lane % CH, rs = lane / CH;
                const Pix pxn = pix_make((uint32_t)tile * NT + 8u * (uint32_t)n, (uint32_t)HW);
                const uint32_t stage = sb + p.offH + (uint32_t)grp * p.hBytes;
                auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); };
                // the residual of the first channel tile is fetched before the wait: its latency hides under GEMM2
                uint4 xr[NCH];
#pragma unroll
                for (int it = 0; it < NCH; ++it) {
                    const int c = 32 * q + it * RPI + rs;
                    xr[it] = (c < C && traffic) ? load8<T, VEC>(gx, pxn, p.B, C, HW, c) : make_uint4(0u, 0u, 0u, 0u);
                }
                tc::mbar_wait(bar(D2_FULL + db), ((uint32_t)t / nD2) & 1u);
                tc::fence_after_sync();
                group_sync();   // (the mbarrier chain already orders this group's H stores before the staging stores below; the barrier
                                //  states it in a form compute-sanitizer's racecheck can follow)
                if (q == 0) stamp(1 + grp, 6 * (g >> 1) + 3);
                for (int ct = 0; ct < nCT; ++ct) {
                    if (ct > 0) {
                        group_sync();   // the previous channel tile has left the staging buffer
#pragma unroll
                        for (int it = 0; it < NCH; ++it) {
                            const int c = ct * 128 + 32 * q + it * RPI + rs;
                            xr[it] = (c < C && traffic) ? load8<T, VEC>(gx, pxn, p.B, C, HW, c) : make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                    // pass A: accumulator row -> + b2 -> 16-bit -> staging (the reference's conv output is a 16-bit tensor too)
                    const float bias2 = sb2[ct * 128 + row];
#pragma unroll
                    for (int c0 = 0; c0 < NT; c0 += 32) {
                        uint32_t v[32];
                        tc::tmem_ld32(trow + 2 * NT + (db * nCT + ct) * NT + c0, v);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) {
                            uint32_t w[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 r = __fadd2_rn(make_float2(__uint_as_float(v[8 * ch + 2 * e]), __uint_as_float(v[8 * ch + 2 * e + 1])), make_float2(bias2, bias2));
                                w[e] = Cvt<T>::pack(r.x, r.y);
                            }
                            sts128(stage + (uint32_t)row * SP + (uint32_t)(c0 + 8 * ch) * 2u, w[0], w[1], w[2], w[3]);
                        }
                    }
                    group_sync();
                    // pass B: lanes along the pixels: out = x + staged
#pragma unroll
                    for (int it = 0; it < NCH; ++it) {
                        const int r = 32 * q + it * RPI + rs, c = ct * 128 + r;
                        uint32_t sw[4];
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(sw[0]), "=r"(sw[1]), "=r"(sw[2]), "=r"(sw[3]) : "r"(stage + (uint32_t)r * SP + (uint32_t)n * 16u));
                        const uint32_t xw[4] = {xr[it].x, xr[it].y, xr[it].z, xr[it].w};
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 r = __fadd2_rn(Cvt<T>::unpack(sw[e]), Cvt<T>::unpack(xw[e]));
                            w[e] = Cvt<T>::pack(r.x, r.y);
                        }
                        if (c < C && traffic) store8<T, VEC>(gout, pxn, p.B, C, HW, c, make_uint4(w[0], w[1], w[2], w[3]));
                    }
                }
                tc::fence_before_sync();
                group_sync();   // staging reads are done before this group's next GELU overwrites the buffer
                if (lane == 0) tc::mbar_arrive(bar(D2_EMPTY + db));
                if (q == 0) stamp(1 + grp, 6 * (g >> 1) + 4);
            }
            s += 2;
            while (s >= nH) { s -= nH; ++t; }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (CS > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into its shared memory or arrive on its barriers
    if (warp == 9) tc::tmem_dealloc(tbase, 512);
}

// one CTA per packed 128-row tile: writes the shared-memory image [k-chunk j][row m][8 elements] of the tile.
// Stream layout: W1 chunk s (s = 0 .. nH-1: K tiles kc = 0 .. nK1-1 of rows 128 s ..), then W2 chunk s (tiles (ct, half): rows 128 ct ..,
// columns 128 s + 64 half ..)
template <typename T>
__global__ void recnext_ffn_pack_kernel(const T* __restrict__ w1, const T* __restrict__ w2, uint8_t* __restrict__ out, int C, int HID, int CP, int nH, int nK1, int nCT,
                                        int kwLast) {
    const int t = blockIdx.x;
    const long w1_step = (long)128 * 2 * CP, w2_step = (long)2 * nCT * kTileBytes;
    int row0, k0, kw, ld, nrows, ncols;
    long off;
    const T* src;
    if (t < nH * nK1) {
        const int s = t / nK1, kc = t - s * nK1;
        row0 = s * 128; k0 = kc * 64; kw = (kc == nK1 - 1) ? kwLast : 64;
        off = (long)s * w1_step + (long)kc * kTileBytes;
        src = w1; ld = C; nrows = HID; ncols = C;
    } else {
        const int u = t - nH * nK1;
        const int s = u / (2 * nCT), v = u - s * 2 * nCT;
        const int ct = v >> 1, half = v & 1;
        row0 = ct * 128; k0 = s * 128 + half * 64; kw = 64;
        off = (long)nH * w1_step + (long)s * w2_step + (long)v * kTileBytes;
        src = w2; ld = HID; nrows = C; ncols = HID;
    }
    unsigned short* dst = reinterpret_cast<unsigned short*>(out + off);
    const unsigned short* s16 = reinterpret_cast<const unsigned short*>(src);
    for (int i = threadIdx.x; i < 128 * kw; i += blockDim.x) {
        const int e = i & 7, m = (i >> 3) & 127, j = i >> 10;
        const int r = row0 + m, k = k0 + 8 * j + e;
        dst[i] = (r < nrows && k < ncols) ? s16[(long)r * ld + k] : (unsigned short)0;
    }
}

}  // namespace

int ffn_tc_make_plan(FfnTcPlan& p, int B, int C, int HID, int HW, int dtype, int num_sms) {
    if (!(dtype == 1 || dtype == 2) || B < 1 || C < 8 || HID < 1 || HW < 1 || (C % 8) != 0) return 1;
    p = FfnTcPlan{};
    p.B = B; p.C = C; p.HID = HID; p.HW = HW; p.dtype = dtype;
    p.P = (long)B * HW;
    if (p.P + 65536 >= (1l << 31)) return 1;   // pixel indices are 32-bit
    p.CP = (C + 15) / 16 * 16;
    p.HIDP = (HID + 127) / 128 * 128;
    p.nH = p.HIDP / 128;
    p.nK1 = (p.CP + 63) / 64;
    p.kwLast = p.CP - 64 * (p.nK1 - 1);
    p.nCT = (C + 127) / 128;
    p.NT = C > 256 ? 64 : 128;
    if (2 * p.NT + p.nCT * p.NT > 512) return 1;          // TMEM columns: D1 x 2 + D2 x nCT
    p.nD2 = (2 * p.NT + 2 * p.nCT * p.NT <= 512) ? 2 : 1;  // a second output accumulator when TMEM has room (C <= 128)
    p.vec = (HW % 8 == 0) ? 8 : ((HW % 4 == 0) ? 4 : 1);
    p.sboY = (uint32_t)p.CP * 16u + 16u;
    p.yBytes = ((uint32_t)(p.NT / 8) * p.sboY + 127u) / 128u * 128u;
    p.hBytes = (uint32_t)(p.NT / 8) * kSboH;
    if (p.hBytes < 128u * (uint32_t)(p.NT * 2 + 16)) p.hBytes = 128u * (uint32_t)(p.NT * 2 + 16);   // also the output staging buffer of a group
    p.hBytes = (p.hBytes + 127u) / 128u * 128u;
    p.offW = 0;
    const uint32_t limit = 227u * 1024u, misc = 8 * NUM_BARS + 64 + 4 * (uint32_t)(p.HIDP + p.nCT * 128);
    // two activation buffers when they fit next to a 4-slot weight ring; then ring slots up to 6, then more activation buffers
    // (small C: the kernel is HBM bound and wants several tiles of loads in flight), then ring slots up to 8
    if (4 * kTileBytes + 2 * p.hBytes + misc + p.yBytes > limit) return 1;
    p.nY = (4 * kTileBytes + 2 * p.hBytes + misc + 2 * p.yBytes <= limit) ? 2 : 1;
    p.RS = 4;
    auto used = [&]() { return (uint32_t)p.RS * kTileBytes + 2 * p.hBytes + misc + (uint32_t)p.nY * p.yBytes; };
    while (p.RS < 6 && used() + kTileBytes <= limit) ++p.RS;
    while (p.nY >= 2 && p.nY < kMaxY && used() + p.yBytes <= limit) ++p.nY;
    while (p.RS < kRing && used() + kTileBytes <= limit) ++p.RS;
    p.offY = (uint32_t)p.RS * kTileBytes;
    p.offH = p.offY + p.nY * p.yBytes;
    p.offBias = p.offH + 2 * p.hBytes;
    p.offBar = p.offBias + 4 * (uint32_t)(p.HIDP + p.nCT * 128);
    p.offBar = (p.offBar + 15u) / 16u * 16u;
    p.smem_bytes = p.offBar + 8 * NUM_BARS + 64;
    p.ntiles = (int)((p.P + p.NT - 1) / p.NT);
    // Weight multicast (RECNEXT_FFN_CS=2|4): the CTAs of a cluster share one weight stream, L2 is read once per cluster.  Measured on
    // B200 it does not pay (C = 256: 0.096 ms with pairs vs 0.087 ms alone): a ring slot is released only when every CTA of the
    // cluster is done with it, which lengthens the refill latency the ring has to cover.  Default: every CTA streams for itself.
    p.cs = 1;
    if (const char* e = getenv("RECNEXT_FFN_CS")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) p.cs = v; }
    int grid = p.ntiles < num_sms ? p.ntiles : num_sms;
    grid = grid / p.cs * p.cs;
    if (grid < p.cs) { p.cs = 1; grid = p.ntiles < num_sms ? p.ntiles : num_sms; }
    p.grid = grid;
    p.tiles_per_cta = (p.ntiles + grid - 1) / grid;
    if (const char* e = getenv("RECNEXT_FFN_DBG")) p.dbg = atoi(e);
    p.packed_bytes = (size_t)p.nH * ((size_t)128 * 2 * p.CP + (size_t)2 * p.nCT * kTileBytes);
    return 0;
}

static long long* g_ffn_prof = nullptr;
template <typename T, int NT, int VEC>
static cudaError_t launch_one(const FfnTcPlan& p, const void* y, const void* x, const void* wpk, const float* b1, const float* b2, void* out, cudaStream_t st) {
    static DeviceOnce configured = {};
    const cudaError_t e = rc_once_per_device(configured, [] {
        return cudaFuncSetAttribute(recnext_ffn_tc_kernel<T, NT, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)p.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, recnext_ffn_tc_kernel<T, NT, VEC>, p, reinterpret_cast<const T*>(y), reinterpret_cast<const T*>(x),
                              reinterpret_cast<const uint8_t*>(wpk), b1, b2, reinterpret_cast<T*>(out), g_ffn_prof);
}
template <typename T>
static cudaError_t launch_t(const FfnTcPlan& p, const void* y, const void* x, const void* wpk, const float* b1, const float* b2, void* out, cudaStream_t st) {
    if (p.NT == 128) {
        if (p.vec == 8) return launch_one<T, 128, 8>(p, y, x, wpk, b1, b2, out, st);
        if (p.vec == 4) return launch_one<T, 128, 4>(p, y, x, wpk, b1, b2, out, st);
        return launch_one<T, 128, 1>(p, y, x, wpk, b1, b2, out, st);
    }
    if (p.vec == 8) return launch_one<T, 64, 8>(p, y, x, wpk, b1, b2, out, st);
    if (p.vec == 4) return launch_one<T, 64, 4>(p, y, x, wpk, b1, b2, out, st);
    return launch_one<T, 64, 1>(p, y, x, wpk, b1, b2, out, st);
}

cudaError_t ffn_tc_launch(const FfnTcPlan& p, const void* y, const void* x, const void* wpk, const float* b1, const float* b2, void* out, cudaStream_t st) {
    if (p.dtype == 1) return launch_t<__nv_bfloat16>(p, y, x, wpk, b1, b2, out, st);
    return launch_t<__half>(p, y, x, wpk, b1, b2, out, st);
}

void ffn_tc_set_prof(long long* buf) { g_ffn_prof = buf; }

cudaError_t ffn_tc_pack(const FfnTcPlan& p, const void* w1, const void* w2, void* packed, cudaStream_t st) {
    const int ntile = p.nH * (p.nK1 + 2 * p.nCT);
    recnext_ffn_pack_kernel<unsigned short><<<ntile, 256, 0, st>>>(reinterpret_cast<const unsigned short*>(w1), reinterpret_cast<const unsigned short*>(w2),
                                                                 reinterpret_cast<uint8_t*>(packed), p.C, p.HID, p.CP, p.nH, p.nK1, p.nCT, p.kwLast);
    return cudaGetLastError();
}

}  // namespace recnext
