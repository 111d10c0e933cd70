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
// Warp roles (512 threads, one CTA per SM):
//   warps 0-7   epilogue, two ping-pong groups of four: TMEM -> registers (tcgen05.ld) -> bias / GELU -> H in shared memory; at the end of a
//               tile D2 + b2 -> 16-bit rows in the group's (then free) H buffer, one 128-channel tile at a time: the accumulator leaves
//               TMEM without waiting for global memory
//   warp  8     weight producer: one elected lane streams the PRE-PACKED weight tiles (16 KB, exactly the shared-memory image
//               of a 128 x 64 K-major operand) through a 4..8-slot ring with TMA bulk copies (cp.async.bulk + mbarrier tx counts)
//   warp  9     MMA issuer: one elected lane issues every tcgen05.mma; tcgen05.commit releases ring slots / publishes accumulators
//   warps 10-11 activation loaders: Y tile global -> shared memory in the MN-major core-matrix layout (cp.async, 16-byte chunks); with one
//               tile buffer (C >= 256) the tile is handed over in 64-channel slabs, each with its own barriers
//   warps 12-15 output writers: staged rows + residual x -> out, lanes along the pixels (coalesced), row cursors, residual prefetched
// Planes whose size is not a multiple of 4 pixels (7 x 7) are padded to a multiple of 8 columns in TILE space (FfnTcPlan::HWp): no 8-pixel
// chunk straddles two images and the padding columns are never stored; loaders and writers of that path take one pixel per lane
// (2-byte accesses, lanes along the pixels: coalesced) and walk down the channels.
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
constexpr int kEpiWarps = 8, kLoadWarps = 2, kWriteWarps = 4;
constexpr int kThreads = 32 * (kEpiWarps + 2 + kLoadWarps + kWriteWarps);   // 512
constexpr int kLoad0 = 32 * (kEpiWarps + 2);                                // first loader thread
constexpr uint32_t kSboH = 128 * 16 + 16;   // H: one 8-pixel chunk of all 128 hidden rows + 16 bytes (conflict-free chunk stores)

template <typename T> struct Cvt;
template <> struct Cvt<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
    static __device__ __forceinline__ float2 unpack(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {   // one correctly rounded 16-bit add per element (HADD2.BF16)
        __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
    static constexpr int fmt = 1;
};
template <> struct Cvt<__half> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __half2 v = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
    static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {
        __half2 r = __hadd2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
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

// Named barriers that restate, in a form compute-sanitizer's racecheck can follow, orderings the mbarrier chains already guarantee
// (staging-buffer hand-offs between an epilogue group and the writer warps).  They never block: every arrival has happened by the
// time the matching mbarrier wait has returned.
__device__ __forceinline__ void nb_arrive(uint32_t id, uint32_t n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_sync(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void sts16(uint32_t addr, unsigned short v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory"); }
__device__ __forceinline__ unsigned short lds16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
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
// size-parametrised forms: `bytes` = 0 zero-fills (src must still be a valid address)
__device__ __forceinline__ void cp_async16z(uint32_t dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8z(uint32_t dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
// predicated global accesses without a branch (a helper warp runs alone on its scheduler: every instruction costs its full latency)
__device__ __forceinline__ uint2 ldg64p(const void* p, bool v) {
    uint2 r;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\t@q ld.global.nc.v2.b32 {%0, %1}, [%2];\n\t}"
                 : "=r"(r.x), "=r"(r.y) : "l"(p), "r"((int)v));
    return r;
}
__device__ __forceinline__ uint4 ldg128p(const void* p, bool v) {
    uint4 r;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t@q ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "r"((int)v));
    return r;
}
__device__ __forceinline__ void stg64p(void* p, uint32_t a, uint32_t b, bool v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q st.global.v2.b32 [%0], {%1, %2};\n\t}" ::"l"(p), "r"(a), "r"(b), "r"((int)v) : "memory");
}
__device__ __forceinline__ void stg128p(void* p, uint4 w, bool v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q st.global.v4.b32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w), "r"((int)v) : "memory");
}

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

// Where the 8-pixel chunk at tile-space position `px` lives: element offsets (for channel 0 of its image) of its aligned pieces.
//   VEC = 8: one 16-byte piece.   VEC = 4: two 8-byte pieces (HW % 8 == 4: the second may lie in the next image).
//   VEC = 1: images are padded to a multiple of 8 columns in TILE SPACE (FfnTcPlan::HWp), so a chunk is `nv` <= 8 consecutive
//            16-bit elements of one row (nv < 8 only for the last chunk of an image); the padding columns compute garbage that is
//            never stored (columns of a GEMM do not mix).
template <int VEC>
struct ChunkAddr {
    long o0, o1;
    bool v0, v1;
    int nv;
    __device__ __forceinline__ ChunkAddr(Pix px, int B, int C, int HW) {
        o0 = ((long)px.b * C) * (long)HW + px.i; v0 = px.b < B;
        if (VEC == 4) {
            const Pix q = pix_add(px, 4, HW);
            o1 = ((long)q.b * C) * (long)HW + q.i; v1 = q.b < B;
        } else { o1 = o0; v1 = false; }
        nv = (VEC == 1 && v0) ? max(0, min(8, HW - px.i)) : 0;
    }
};
// the element-wise forms (VEC = 1): `row` points at the chunk's first element of one channel row
__device__ __forceinline__ uint4 load8e(const unsigned short* __restrict__ row, int nv) {
    uint32_t e[8];
    if (nv == 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) e[u] = __ldg(row + u);
    } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) e[u] = u < nv ? (uint32_t)__ldg(row + u) : 0u;
    }
    return make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
}
__device__ __forceinline__ void store8e(unsigned short* __restrict__ row, int nv, uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (nv == 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) row[u] = (unsigned short)((w[u >> 1] >> (16 * (u & 1))) & 0xffffu);
    } else {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (u < nv) row[u] = (unsigned short)((w[u >> 1] >> (16 * (u & 1))) & 0xffffu);
    }
}

constexpr int kMaxY = 12;  // activation tile buffers (at most 4) -- or, with ONE buffer, its 64-channel slabs (C <= 768), each with its own barriers
enum Bar { W_FULL = 0, W_EMPTY = kRing, Y_FULL = 2 * kRing, Y_EMPTY = Y_FULL + kMaxY, D1_FULL = Y_EMPTY + kMaxY, H_FULL = D1_FULL + 2,
           H_EMPTY = H_FULL + 2, D2_FULL = H_EMPTY + 2, D2_EMPTY = D2_FULL + 2, OUT_FULL = D2_EMPTY + 2, OUT_EMPTY = OUT_FULL + 2,
           NUM_BARS = OUT_EMPTY + 2 };

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
            tc::mbar_init(bar(OUT_FULL + i), 4); tc::mbar_init(bar(OUT_EMPTY + i), kWriteWarps);
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

    // Order of the weight stream (producer and MMA issuer walk it identically): per chunk g the W1 tiles of chunk g, then the W2 tiles
    // of chunk g - 1 -- except at a tile border when there is ONE activation buffer: then GEMM2 of the previous tile's last chunk goes
    // first, so that tile's output leaves (and its accumulator frees) while the next activation tile is still loading.
    const bool border_swap = (p.nY == 1);
    if (warp == 8) {
        // ================= weight producer =================
        // In a cluster every CTA fetches 1/CS of each tile and multicasts it to all of them: L2 is read once per cluster.
        const bool leader = elect_one();
        const uint8_t* w2base = wpk + (size_t)nH * w1_step;
        const uint32_t wfull0 = bar(W_FULL), wempty0 = bar(W_EMPTY), ring0 = sb + p.offW;
        const uint32_t last_bytes = (uint32_t)(128 * 2) * (uint32_t)p.kwLast;
        uint32_t slot = 0, phase = 1;    // (phase of the EMPTY barrier a fresh slot passes immediately)
        auto put = [&](const uint8_t* src, uint32_t bytes) {
            tc::mbar_wait(wempty0 + 8u * slot, phase);
            if (leader) {
                if (p.dbg & 1) tc::mbar_arrive(wfull0 + 8u * slot);   // timing experiment: no weight traffic (results are garbage)
                else {
                    tc::mbar_expect_tx(wfull0 + 8u * slot, bytes);
                    if (CS == 1) tc::bulk_g2s(ring0 + slot * kTileBytes, src, bytes, wfull0 + 8u * slot);
                    else {
                        const uint32_t part = bytes / (uint32_t)CS;
                        bulk_g2s_mcast(ring0 + slot * kTileBytes + crank * part, src + crank * part, part, wfull0 + 8u * slot, cmask);
                    }
                }
            }
            if (++slot == (uint32_t)RS) { slot = 0; phase ^= 1u; }
        };
        int s = 0;
        for (int g = 0; g <= G; ++g) {
            const int s_prev = s == 0 ? nH - 1 : s - 1;
            const uint8_t* src1 = wpk + (size_t)s * w1_step;
            const uint8_t* src2 = w2base + (size_t)s_prev * w2_step;
            const bool second_first = border_swap && s == 0;
            for (int pass = 0; pass < 2; ++pass) {
                if ((pass == 0) != second_first) {
                    if (g < G) {
                        for (int t = 0; t < nK1 - 1; ++t) put(src1 + (size_t)t * kTileBytes, kTileBytes);
                        put(src1 + (size_t)(nK1 - 1) * kTileBytes, last_bytes);
                    }
                } else if (g >= 1) {
                    for (int t = 0; t < 2 * nCT; ++t) put(src2 + (size_t)t * kTileBytes, kTileBytes);
                }
            }
            if (++s == nH) s = 0;
        }
    } else if (warp == 9) {
        // ================= MMA issuer =================
        // One warp, warp-uniform control flow, one elected lane issues.  The loop is the critical resource of the kernel for C >= 128
        // (a hand-off costs the tensor pipe as much as the instructions between two tcgen05.mma take to issue), so ring positions,
        // barrier phases and descriptors are carried incrementally: no divisions, nothing recomputed per slot.
        const bool leader = elect_one();
        const uint32_t idesc = tc::make_idesc(128, NT, Cvt<T>::fmt, 0, 1);
        const uint32_t wfull0 = bar(W_FULL), wempty0 = bar(W_EMPTY);
        const uint64_t a_ring = tc::make_sdesc(sb + p.offW, 2048, 128);
        const uint64_t y_desc0 = tc::make_sdesc(sb + p.offY, 128, p.sboY);
        const uint64_t h_desc0 = tc::make_sdesc(sb + p.offH, 128, kSboH);
        const int last_steps = p.kwLast / 16;
        uint32_t wslot = 0, wphase = 0;
        // four (or `steps`) K steps of one ring slot: D (+)= A[slot] . B
        long long acc_wait = 0, acc_issue = 0, acc_slots = 0;   // (per-slot clock reads were removed from the hot loop: `acc_slots` only)
        const bool profiling = prof != nullptr && blockIdx.x == 0;
        // The barrier of the NEXT slot is probed (non-blocking) before this slot's MMAs are issued: the probe's ~100-clock latency
        // passes while the tensor pipe takes the MMAs, instead of in front of the next slot.
        bool have = false;                          // W_FULL of the current slot has already been observed
        const bool solo = CS == 1;
        const uint32_t rs_last = (uint32_t)RS - 1u;
        uint32_t wfull = wfull0, wempty = wempty0;  // barrier addresses and A descriptor of the current slot, advanced by increments
        uint64_t a_desc = a_ring;
        auto slot_mmas = [&](uint32_t d, uint64_t bd, bool fresh, int steps) {
            if (!have) tc::mbar_wait(wfull, wphase);
            tc::fence_after_sync();
            const bool wrap = wslot == rs_last;
            const uint32_t nfull = wrap ? wfull0 : wfull + 8u, nphase = wrap ? wphase ^ 1u : wphase;
            const bool next_ready = tc::mbar_test_wait(nfull, nphase);
            const uint64_t ad = a_desc;
            if (steps == 4) {
                if (leader) {
                    tc::mma_ss(d, ad, bd, idesc, fresh ? 0u : 1u);
                    tc::mma_ss(d, ad + 256, bd + 16, idesc, 1u);
                    tc::mma_ss(d, ad + 512, bd + 32, idesc, 1u);
                    tc::mma_ss(d, ad + 768, bd + 48, idesc, 1u);
                }
            } else {
                for (int j = 0; j < steps; ++j)
                    if (leader) tc::mma_ss(d, ad + (uint64_t)(256 * j), bd + (uint64_t)(16 * j), idesc, (fresh && j == 0) ? 0u : 1u);
            }
            if (leader) {
                if (solo) tc::mma_commit(wempty);
                else mma_commit_mcast(wempty, cmask);
            }
            have = next_ready;
            wfull = nfull; wphase = nphase;
            wempty = wrap ? wempty0 : wempty + 8u;
            a_desc = wrap ? a_ring : a_desc + (uint64_t)(kTileBytes / 16);
            wslot = wrap ? 0u : wslot + 1u;
            ++acc_slots;
        };
        int s = 0;                                  // chunk g = (tile of this CTA, hidden chunk s)
        uint32_t ybuf = 0, yphase = 0;              // activation buffer of chunk g's tile
        int s2 = 0;                                 // chunk g - 1
        uint32_t db = 0, dphase = 1;                // output accumulator of chunk g - 1's tile (EMPTY barrier: a fresh one passes)
        const bool slabs = (nY == 1);                // one activation buffer, handed over in 64-channel slabs (barrier per slab)
        auto y_acquire = [&](uint32_t bi) {
            tc::mbar_wait(bar(Y_FULL) + 8u * bi, yphase);
            tc::fence_proxy_async();   // the loaders' cp.async / st.shared writes -> the tensor core's async-proxy reads
            tc::fence_after_sync();
        };
        auto gemm1 = [&](int g) {                   // D1[g & 1] = W1[s] . Y
            if (s == 0 && !slabs) y_acquire(ybuf);
            const uint32_t buf = (uint32_t)g & 1u;
            const uint32_t d1 = tbase + buf * NT;
            const uint64_t bd0 = y_desc0 + (uint64_t)(ybuf * (p.yBytes / 16));
            stamp(0, 4 * g + 0);
            for (int kc = 0; kc < nK1; ++kc) {
                if (slabs && s == 0) y_acquire((uint32_t)kc);
                slot_mmas(d1, bd0 + (uint64_t)(kc * 64), kc == 0, kc == nK1 - 1 ? last_steps : 4);
                if (slabs && s == nH - 1 && leader) tc::mma_commit(bar(Y_EMPTY) + 8u * (uint32_t)kc);   // the next tile's slab may land
            }
            if (leader) {
                tc::mma_commit(bar(D1_FULL) + 8u * buf);
                if (s == nH - 1 && !slabs) tc::mma_commit(bar(Y_EMPTY) + 8u * ybuf);   // the loaders may fetch this buffer's next tile
            }
            stamp(0, 4 * g + 1);
            if (++s == nH) {
                s = 0;
                if (slabs) yphase ^= 1u;
                else if (++ybuf == (uint32_t)nY) { ybuf = 0; yphase ^= 1u; }
            }
        };
        auto gemm2 = [&](int gg) {                  // D2 += W2[:, s2] . H[gg & 1]
            const uint32_t buf = (uint32_t)gg & 1u;
            tc::mbar_wait(bar(H_FULL) + 8u * buf, ((uint32_t)gg >> 1) & 1u);
            if (s2 == 0) tc::mbar_wait(bar(D2_EMPTY) + 8u * db, dphase);   // that accumulator's previous tile has left TMEM
            tc::fence_after_sync();
            stamp(0, 4 * gg + 2);
            const uint64_t hd = h_desc0 + (uint64_t)(buf * (p.hBytes / 16));
            for (int ct = 0; ct < nCT; ++ct) {
                const uint32_t d2 = tbase + 2 * NT + (db * nCT + ct) * NT;
                slot_mmas(d2, hd, s2 == 0, 4);
                slot_mmas(d2, hd + 64, false, 4);
            }
            if (leader) {
                tc::mma_commit(bar(H_EMPTY) + 8u * buf);
                if (s2 == nH - 1) tc::mma_commit(bar(D2_FULL) + 8u * db);
            }
            stamp(0, 4 * gg + 3);
            if (++s2 == nH) {
                s2 = 0;
                if (++db == (uint32_t)nD2) { db = 0; dphase ^= 1u; }
            }
        };
        const long long role0 = profiling ? clock64() : 0;
        for (int g = 0; g <= G; ++g) {
            const bool second_first = border_swap && s == 0;
            if (second_first && g >= 1) gemm2(g - 1);
            if (g < G) gemm1(g);
            if (!second_first && g >= 1) gemm2(g - 1);
        }
        if (profiling && lane == 0) { prof[500] = acc_wait; prof[501] = acc_issue; prof[502] = acc_slots; prof[503] = clock64() - role0; }
    } else if (warp >= 10 && warp < 10 + kLoadWarps) {
        // ================= activation loaders =================
        // Fully asynchronous: a thread issues the cp.async copies of its part of a tile and attaches an mbarrier arrival to them (no
        // wait): up to nY tiles are in flight.  With ONE buffer (C >= 256) the tile is handed over in 64-channel slabs, each with its
        // own barriers: the slabs of the next tile arrive while the last GEMM1 of this tile still reads the later ones.
        constexpr int CH = NT / 8;                       // 8-pixel chunks per channel row of a tile
        constexpr int KPP = (kLoadWarps * 32) / CH;      // channels per pass
        const int lt = threadIdx.x - kLoad0;
        const int n = lt % CH, k0 = lt / CH;
        const bool slabs = (nY == 1);
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const uint32_t ybuf = slabs ? 0u : (uint32_t)t % nY;
            const uint32_t yph = (slabs ? (uint32_t)t : (uint32_t)t / nY) & 1u;
            const uint32_t dst = sb + p.offY + ybuf * p.yBytes + (uint32_t)n * p.sboY;
            const Pix px = pix_make((uint32_t)tile * NT + 8u * n, (uint32_t)p.HWp);
            const ChunkAddr<VEC> ca(px, (p.dbg & 8) ? 0 : p.B, C, HW);
            const int nslab = slabs ? nK1 : 1;
            for (int sl = 0; sl < nslab; ++sl) {
                const uint32_t bi = slabs ? (uint32_t)sl : ybuf;
                const int kbeg = slabs ? 64 * sl : 0, kend = slabs ? min(64 * sl + 64, p.CP) : p.CP;
                tc::mbar_wait(bar(Y_EMPTY) + 8u * bi, yph ^ 1u);
                if (warp == 10 && sl == 0) stamp(3, 2 * t);
                if (VEC >= 4) {
                    // pointer increments only: an invalid piece (pixels past the batch) reads gy[0] with size 0 = zero fill
                    const long rowb = (long)HW * (long)sizeof(T);
                    const char* s0 = ca.v0 ? reinterpret_cast<const char*>(gy + ca.o0) + (long)(kbeg + k0) * rowb : reinterpret_cast<const char*>(gy);
                    const char* s1 = ca.v1 ? reinterpret_cast<const char*>(gy + ca.o1) + (long)(kbeg + k0) * rowb : reinterpret_cast<const char*>(gy);
                    const long st0 = ca.v0 ? (long)KPP * rowb : 0, st1 = ca.v1 ? (long)KPP * rowb : 0;
                    const uint32_t z0 = ca.v0 ? (VEC == 8 ? 16u : 8u) : 0u, z1 = ca.v1 ? 8u : 0u;
                    uint32_t d = dst + (uint32_t)(kbeg + k0) * 16u;
                    const int kreal = min(kend, C);
                    int k = kbeg + k0;
                    for (; k < kreal; k += KPP) {
                        if (VEC == 8) cp_async16z(d, s0, z0);
                        else { cp_async8z(d, s0, z0); cp_async8z(d + 8u, s1, z1); }
                        s0 += st0; s1 += st1; d += (uint32_t)KPP * 16u;
                    }
                    for (; k < kend; k += KPP) {   // channels padded up to a multiple of 16: zeros
                        if (VEC == 8) cp_async16z(d, gy, 0u);
                        else { cp_async8z(d, gy, 0u); cp_async8z(d + 8u, gy, 0u); }
                        d += (uint32_t)KPP * 16u;
                    }
                    cp_async_arrive_noinc(bar(Y_FULL) + 8u * bi);
                } else {
                    // Odd planes (7 x 7): no aligned vector covers an 8-pixel chunk, so the lanes run ALONG the tile's pixels (a warp's
                    // 2-byte loads of one channel are consecutive addresses: 2-3 sectors per request, where a lane-per-chunk mapping
                    // costs 32) and every lane walks down the channels of its pixel: pointer + HW in global, + 16 bytes in the tile.
                    constexpr int NPOS = NT / (kLoadWarps * 32);
                    constexpr int U = 64 / NPOS;                 // 64 two-byte loads in flight per lane
                    const unsigned short* g16 = reinterpret_cast<const unsigned short*>(gy);
                    const unsigned short* src[NPOS];
                    uint32_t d[NPOS];
                    bool ok[NPOS];
#pragma unroll
                    for (int q = 0; q < NPOS; ++q) {
                        const int m = lt + q * (kLoadWarps * 32);
                        const Pix pm = pix_make((uint32_t)tile * NT + (uint32_t)m, (uint32_t)p.HWp);
                        ok[q] = pm.b < ((p.dbg & 8) ? 0 : p.B) && pm.i < HW;
                        src[q] = g16 + (ok[q] ? ((long)pm.b * C) * (long)HW + pm.i : 0) + (long)kbeg * HW;
                        d[q] = sb + p.offY + ybuf * p.yBytes + (uint32_t)(m >> 3) * p.sboY + (uint32_t)(m & 7) * 2u + (uint32_t)kbeg * 16u;
                    }
                    for (int kb = kbeg; kb < kend; kb += U) {      // (kend - kbeg is a multiple of 16; U of 32 or 64 may overshoot: guarded)
                        unsigned short e[NPOS][U];
#pragma unroll
                        for (int q = 0; q < NPOS; ++q)
#pragma unroll
                            for (int u = 0; u < U; ++u) e[q][u] = (ok[q] && kb + u < C) ? __ldg(src[q] + (long)u * HW) : (unsigned short)0;
#pragma unroll
                        for (int q = 0; q < NPOS; ++q) {
#pragma unroll
                            for (int u = 0; u < U; ++u)
                                if (kb + u < kend) sts16(d[q] + (uint32_t)u * 16u, e[q][u]);
                            src[q] += (long)U * HW; d[q] += (uint32_t)U * 16u;
                        }
                    }
                    tc::mbar_arrive(bar(Y_FULL) + 8u * bi);
                }
            }
            if (warp == 10) stamp(3, 2 * t + 1);
        }
        if (VEC >= 4) cp_async_wait_all();
    } else if (warp >= 10 + kLoadWarps) {
        // ================= output writers =================
        // An epilogue group leaves the finished rows of a channel tile (D2 + b2, 16-bit) in its staging buffer; these warps add the
        // residual with their lanes running ALONG the pixels (coalesced x loads and out stores) and store.  The accumulator leaves TMEM
        // without waiting for HBM, and the epilogue warps (the scarce resource for small C) do not spend issue slots on the stores.
        // The residual rows are fetched 8 iterations ahead, across channel-tile and tile borders of the wait for the staged rows.
        constexpr int CH = NT / 8, RPI = 32 / CH;        // 8-pixel chunks per row; rows per warp pass
        constexpr uint32_t SP = NT * 2 + 16;             // staging row pitch (bytes): conflict-free for both access patterns
        constexpr int PF = 8;                            // prefetch distance (iterations); CH is a multiple of it
        const int q = warp - (10 + kLoadWarps), no = lane % CH, rs = lane / CH;
        const bool traffic = !(p.dbg & 4);
        uint32_t ocnt0 = 0, ocnt1 = 0;   // staging hand-offs received from epilogue group 0 / 1
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const uint32_t grp = (uint32_t)(t * nH + nH - 1) & 1u;   // the group that owns the tile's last hidden chunk
            const uint32_t stage = sb + p.offH + grp * p.hBytes + (uint32_t)no * 16u;
            const Pix pxn = pix_make((uint32_t)tile * NT + 8u * (uint32_t)no, (uint32_t)p.HWp);
            const ChunkAddr<VEC> ca(pxn, (traffic && !(p.dbg & 16)) ? p.B : 0, C, HW);    // residual loads
            const ChunkAddr<VEC> cs(pxn, (traffic && !(p.dbg & 32)) ? p.B : 0, C, HW);    // stores
            // Pass j = ct * CH + it: the four warps take the rows [4 RPI it, 4 RPI (it + 1)) of channel tile ct, RPI rows each, so narrow
            // tensors (C = 64: half a channel tile) keep every warp busy; the last channel tile stops after its last valid row
            // (rounded up to the prefetch block).
            const int last_rows = C - (nCT - 1) * 128;
            const int total = (nCT - 1) * CH + min(CH, ((last_rows + 4 * RPI - 1) / (4 * RPI) + PF - 1) / PF * PF);
            auto rowof = [&](int j) { return (j % CH) * (4 * RPI) + q * RPI + rs; };
            auto chan = [&](int j) { return (j / CH) * 128 + rowof(j); };
            if (VEC >= 4) {
                // Row cursors: the loop body is a shared-memory load, one or two global loads (PF iterations ahead), packed 16-bit
                // adds, one or two stores and pointer increments.
                const long rowb = (long)HW * (long)sizeof(T);
                const long step = (long)(4 * RPI) * rowb;             // (a channel-tile border is just the next pass: passes tile the rows)
                const long first = (long)(q * RPI + rs) * rowb;
                const char* xp0 = reinterpret_cast<const char*>(gx + ca.o0) + first;   // prefetch cursors (iteration j + PF)
                const char* xp1 = reinterpret_cast<const char*>(gx + ca.o1) + first;
                char* wp0 = reinterpret_cast<char*>(gout + cs.o0) + first;             // store cursors (iteration j)
                char* wp1 = reinterpret_cast<char*>(gout + cs.o1) + first;
                uint4 xr[PF];
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const bool in = chan(u) < C;
                    if (VEC == 8) xr[u] = ldg128p(xp0, in && ca.v0);
                    else { const uint2 lo = ldg64p(xp0, in && ca.v0), hi = ldg64p(xp1, in && ca.v1); xr[u] = make_uint4(lo.x, lo.y, hi.x, hi.y); }
                    xp0 += step; xp1 += step;
                }
                for (int jb = 0; jb < total; jb += PF) {
                    if (jb % CH == 0) {
                        tc::mbar_wait(bar(OUT_FULL) + 8u * grp, (grp ? ocnt1 : ocnt0) & 1u);
                        nb_sync(3u + grp, 256u);
                        if (warp == 10 + kLoadWarps) stamp(3, 256 + 4 * t + 2 * (jb / CH));
                    }
                    const bool ct_end = (jb + PF) % CH == 0 || jb + PF == total;   // this channel tile's staged rows are done after this block
#pragma unroll
                    for (int u = 0; u < PF; ++u) {
                        const int j = jb + u, it = j % CH;
                        uint32_t sw[4];
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(sw[0]), "=r"(sw[1]), "=r"(sw[2]), "=r"(sw[3]) : "r"(stage + (uint32_t)rowof(j) * SP));
                        const uint4 xw = xr[u];
                        const bool in_next = (j + PF < total) && chan(j + PF) < C;
                        if (VEC == 8) xr[u] = ldg128p(xp0, in_next && ca.v0);
                        else { const uint2 lo = ldg64p(xp0, in_next && ca.v0), hi = ldg64p(xp1, in_next && ca.v1); xr[u] = make_uint4(lo.x, lo.y, hi.x, hi.y); }
                        const uint4 w = make_uint4(Cvt<T>::add2(sw[0], xw.x), Cvt<T>::add2(sw[1], xw.y), Cvt<T>::add2(sw[2], xw.z), Cvt<T>::add2(sw[3], xw.w));
                        const bool in = chan(j) < C;
                        if (VEC == 8) stg128p(wp0, w, in && cs.v0);
                        else { stg64p(wp0, w.x, w.y, in && cs.v0); stg64p(wp1, w.z, w.w, in && cs.v1); }
                        xp0 += step; xp1 += step; wp0 += step; wp1 += step;
                    }
                    if (ct_end) {
                        __syncwarp();
                        nb_arrive(5u + grp, 256u);
                        if (lane == 0) tc::mbar_arrive(bar(OUT_EMPTY) + 8u * grp);   // the staging buffer may be overwritten
                        if (grp) ++ocnt1; else ++ocnt0;
                        if (warp == 10 + kLoadWarps) stamp(3, 256 + 4 * t + 2 * (jb / CH) + 1);
                    }
                }
            } else {
                // Odd planes: one tile pixel per lane (lanes along the pixels: coalesced 2-byte residual loads and stores), the lane
                // walks down the staged channel rows; the residuals of a block of rows are in flight while the previous block is stored.
                constexpr int NPL = (kWriteWarps * 32) / NT;     // row phases (NT = 64: two lanes per pixel, alternate rows)
                constexpr int U = 16;
                const int wl = (int)threadIdx.x - (kLoad0 + kLoadWarps * 32);
                const int m = wl % NT, rph = wl / NT;
                const Pix pm = pix_make((uint32_t)tile * NT + (uint32_t)m, (uint32_t)p.HWp);
                const bool inimg = pm.b < p.B && pm.i < HW;
                const bool okl = inimg && traffic && !(p.dbg & 16), oks = inimg && traffic && !(p.dbg & 32);
                const long o = inimg ? ((long)pm.b * C) * (long)HW + pm.i : 0;
                const unsigned short* xs = reinterpret_cast<const unsigned short*>(gx) + o;
                unsigned short* os = reinterpret_cast<unsigned short*>(gout) + o;
                const uint32_t st = sb + p.offH + grp * p.hBytes + (uint32_t)m * 2u;
                for (int ct = 0; ct < nCT; ++ct) {
                    const int c0 = ct * 128, rows = min(128, C - c0);
                    unsigned short e[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) { const int r = rph + NPL * u; e[u] = (okl && r < rows) ? __ldg(xs + (long)(c0 + r) * HW) : (unsigned short)0; }
                    tc::mbar_wait(bar(OUT_FULL) + 8u * grp, (grp ? ocnt1 : ocnt0) & 1u);
                    nb_sync(3u + grp, 256u);
                    if (warp == 10 + kLoadWarps) stamp(3, 256 + 4 * t + 2 * ct);
                    for (int rb = 0; rb < rows; rb += U * NPL) {
                        unsigned short cur[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) cur[u] = e[u];
                        if (rb + U * NPL < rows) {
#pragma unroll
                            for (int u = 0; u < U; ++u) { const int r = rb + U * NPL + rph + NPL * u; e[u] = (okl && r < rows) ? __ldg(xs + (long)(c0 + r) * HW) : (unsigned short)0; }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int r = rb + rph + NPL * u;
                            if (r < rows) {
                                const uint32_t w = Cvt<T>::add2((uint32_t)lds16(st + (uint32_t)r * SP), (uint32_t)cur[u]);
                                if (oks) os[(long)(c0 + r) * HW] = (unsigned short)(w & 0xffffu);
                            }
                        }
                    }
                    __syncwarp();
                    nb_arrive(5u + grp, 256u);
                    if (lane == 0) tc::mbar_arrive(bar(OUT_EMPTY) + 8u * grp);   // the staging buffer may be overwritten
                    if (grp) ++ocnt1; else ++ocnt0;
                    if (warp == 10 + kLoadWarps) stamp(3, 256 + 4 * t + 2 * ct + 1);
                }
            }
        }
    } else {
        // ================= epilogue warps: two groups of four, chunk g belongs to group g & 1 =================
        // A group owns D1[grp] / H[grp]: while one group runs the GELU of chunk g, the other waits for (or already works on) chunk g + 1,
        // so barrier and memory latencies of the two chains overlap.  A thread owns one accumulator row (TMEM lane) and all NT columns.
        const int q = warp & 3, grp = warp >> 2;
        const int row = 32 * q + lane;
        const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
        const uint32_t hdst = sb + p.offH + (uint32_t)grp * p.hBytes + (uint32_t)row * 16u;
        const uint32_t out_full = bar(OUT_FULL) + 8u * (uint32_t)grp, out_empty = bar(OUT_EMPTY) + 8u * (uint32_t)grp;
        uint32_t out_uses = 0;       // staging hand-offs of this group so far: H[grp] doubles as its staging buffer
        uint32_t out_synced = 0;     // ... of which the writers' release has been consumed on the named barrier (racecheck's view)
        auto wait_drained = [&]() {  // the writers have taken the last staged rows out of H[grp]
            tc::mbar_wait(out_empty, (out_uses & 1u) ^ 1u);
            while (out_synced < out_uses) { nb_sync(5u + (uint32_t)grp, 256u); ++out_synced; }
        };
        int s = grp % nH, t = grp / nH;   // chunk g = grp, grp + 2, ...
        for (int g = grp; g < G; g += 2) {
            if (q == 0) stamp(1 + grp, 6 * (g >> 1) + 0);
            // ---- H[grp] = gelu(D1[grp] + b1[s])
            const uint32_t use = (uint32_t)g >> 1;
            const float bias = sb1[s * 128 + row];
            const float2 bias2v = make_float2(bias, bias);
            tc::mbar_wait(bar(H_EMPTY + grp), (use & 1u) ^ 1u);   // GEMM2 of chunk g - 2 has read this H buffer
            wait_drained();                                       // ... and the writers have taken the last staged rows out of it
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
            // ---- last chunk of a tile: once its GEMM2 has finished, the accumulator rows (one channel per lane) + b2 go, as 16-bit
            // values (the reference's conv output is a 16-bit tensor too), into the staging buffer, one channel tile at a time; the
            // writer warps add the residual and store.  The accumulator is free again after the last tcgen05.ld: GEMM2 of the next tile
            // does not wait for global memory.
            if (s == nH - 1) {
                constexpr uint32_t SP = NT * 2 + 16;
                const uint32_t stage = sb + p.offH + (uint32_t)grp * p.hBytes;
                const uint32_t db = nD2 == 2 ? ((uint32_t)t & 1u) : 0u, dphase = (nD2 == 2 ? ((uint32_t)t >> 1) : (uint32_t)t) & 1u;
                tc::mbar_wait(bar(D2_FULL) + 8u * db, dphase);
                tc::fence_after_sync();
                nb_sync(1u + (uint32_t)grp, 128u);   // (this group's H stores of the last chunk precede the staging stores below: the mbarrier chain
                                                     //  H_FULL -> GEMM2 -> D2_FULL orders them; stated for racecheck)
                if (q == 0) stamp(1 + grp, 6 * (g >> 1) + 3);
                for (int ct = 0; ct < nCT; ++ct) {
                    wait_drained();                                  // the previous channel tile has left the staging buffer
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
                    __syncwarp();
                    nb_arrive(3u + (uint32_t)grp, 256u);
                    if (lane == 0) tc::mbar_arrive(out_full);
                    ++out_uses;
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(bar(D2_EMPTY) + 8u * db);
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
    p.HWp = p.vec == 1 ? (HW + 7) / 8 * 8 : HW;
    p.P = (long)B * p.HWp;
    if (p.P + 65536 >= (1l << 31)) return 1;   // pixel indices are 32-bit
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
