// stem.cu — the RecNeXt stem as ONE kernel (inference, 16-bit activations):
//     ConvNorm(3 -> C/2, 3x3, stride 2, pad 1) -> GELU -> ConvNorm(C/2 -> C, 3x3, stride 2, pad 1)          model/recnext.py:139-146
// with both BatchNorms folded into the convs.  The library path is five launches per image batch (two convs, the layout
// transposes cuDNN wants around them, bias adds, the GELU) that move the 112 x 112 intermediate through HBM four times; here the
// intermediate lives in shared memory and HBM sees the image once and the output once.
//
// A CTA owns an 8 x 8 tile of output pixels of one image.  Per tile:
//   1. the 35 x 35 x Cin input patch -> shared memory ([ci][y][x], zero outside the image = conv1's padding)
//   2. conv1 as an implicit GEMM on the tensor cores (mma.sync m16n8k16; M = the 17 x 17 intermediate pixels conv2 needs, N = C/2,
//      K = 9 Cin padded to a multiple of 16; the A fragments are gathered element-wise from the patch, the weights sit in registers)
//      -> + bias -> 16-bit -> GELU -> 16-bit -> shared memory, pixel-major with the columns split by parity (so conv2's stride-2
//      windows are unit-stride ldmatrix rows); positions outside the intermediate map are ZERO (conv2's padding)
//   3. conv2 as an implicit GEMM (M = 64 output pixels, N = C, K = 9 taps x C/2): A by ldmatrix from the intermediate tile, B by
//      ldmatrix from the weight image in shared memory -> + bias -> 16-bit -> NCHW stores.
// Roundings follow the reference's autocast graph (conv outputs and the GELU output are 16-bit tensors).  GELU uses the tanh fit
// of ffn_tc.cu (max |error| 2.7e-4 against the erf form, below the 16-bit rounding that follows).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "devcfg.h"
#include "stem.h"

namespace recnext {

namespace {

constexpr int kTile = 8;                 // output pixels per tile edge
constexpr int kMid = 2 * kTile + 1;      // 17: intermediate pixels per tile edge
constexpr int kIn = 2 * kMid + 1;        // 35: input pixels per tile edge
constexpr int kInPitch = 40;             // elements per input row in shared memory: the patch row sits at columns 1 .. 35 (the window that starts one
                                         // column to its left is 8-byte aligned in global memory: 10 cp.async pieces of 4 pixels per row)
constexpr int kHalf = (kMid + 1) / 2;    // 9: columns of one parity plane of the intermediate tile

template <typename T> struct H16;
template <> struct H16<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
    static __device__ __forceinline__ float2 round2(float2 v) { return __bfloat1622float2(__float22bfloat162_rn(v)); }
    static __device__ __forceinline__ unsigned short bits(float v) { __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<unsigned short*>(&h); }
};
template <> struct H16<__half> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __half2 v = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
    static __device__ __forceinline__ float2 round2(float2 v) { return __half22float2(__float22half2_rn(v)); }
    static __device__ __forceinline__ unsigned short bits(float v) { __half h = __float2half_rn(v); return *reinterpret_cast<unsigned short*>(&h); }
};

template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// gelu for two elements with the packed fp32x2 instructions of sm_100 (the tanh fit of ffn_tc.cu)
__device__ __forceinline__ float2 gelu_fit2(float2 t) {
    const float2 t2 = __fmul2_rn(t, t);
    const float2 q = __ffma2_rn(t2, make_float2(0.03470089f, 0.03470089f), make_float2(0.80015708f, 0.80015708f));
    const float2 u = __fmul2_rn(t, q);
    float2 th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(u.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(u.y));
    const float2 h = __fmul2_rn(t, make_float2(0.5f, 0.5f));
    return __ffma2_rn(h, th, h);
}

// NP1 = n-tiles (8 channels) of conv1 = K16 steps x 2 of conv2 per tap: C/2 padded to 16 -> KS = C1P / 16 in {2, 3}
template <typename T, int KS>
__global__ void __launch_bounds__(KS == 2 ? 256 : 320, KS == 2 ? 3 : 2) recnext_stem_kernel(const __grid_constant__ StemPlan p, const T* __restrict__ x, const T* __restrict__ w1p,
                                                            const float* __restrict__ b1p, const T* __restrict__ w2p, const float* __restrict__ b2p,
                                                            T* __restrict__ out) {
    constexpr int C1P = 16 * KS, NT1 = C1P / 8;
    constexpr int MIDP = C1P * 2 + 16;                  // bytes per intermediate pixel (odd multiple of 16: conflict-free ldmatrix rows)
    constexpr int K2 = 9 * C1P;                         // K of conv2
    constexpr int W2P = K2 * 2 + 16;                    // bytes per output-channel row of the conv2 weight image
    extern __shared__ __align__(128) uint8_t smem[];
    unsigned short* s_in = reinterpret_cast<unsigned short*>(smem);                 // [Cin = 3][35][36]
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t s_mid = sb + p.offMid;               // [17 rows][2 parities][9 columns][MIDP bytes]
    const uint32_t s_w2 = sb + p.offW2;                 // [C2P rows][W2P bytes]
    float* s_b2 = reinterpret_cast<float*>(smem + p.offB2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int nwarps = blockDim.x >> 5;
    const int H = p.H, W = p.W, H1 = p.H1, W1 = p.W1, H2 = p.H2, W2 = p.W2, C2 = p.C2;

    // ---- per-CTA constants: conv2 weight image and bias -> shared memory; conv1 weights -> fragment registers
    for (int i = tid; i < p.C2P * (K2 / 8); i += blockDim.x) {
        const int n = i / (K2 / 8), c = i - n * (K2 / 8);
        *reinterpret_cast<uint4*>(smem + p.offW2 + n * W2P + c * 16) = *reinterpret_cast<const uint4*>(w2p + (size_t)n * K2 + c * 8);
    }
    for (int i = tid; i < p.C2P; i += blockDim.x) s_b2[i] = b2p[i];
    uint32_t bw1[NT1][2][2];       // B fragments of conv1: [n-tile][k-step][2]; w1p is [C1P][32] (k = ci * 9 + ky * 3 + kx, zero padded)
    float bias1[NT1][2];
#pragma unroll
    for (int nt = 0; nt < NT1; ++nt) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const T* row = w1p + (size_t)(nt * 8 + g) * 32 + ks * 16 + 2 * t4;
            bw1[nt][ks][0] = *reinterpret_cast<const uint32_t*>(row);
            bw1[nt][ks][1] = *reinterpret_cast<const uint32_t*>(row + 8);
        }
        bias1[nt][0] = b1p[nt * 8 + 2 * t4];
        bias1[nt][1] = b1p[nt * 8 + 2 * t4 + 1];
    }
    // shared-memory offsets (elements) of this thread's 8 K positions of conv1's A fragment: k -> (ci, ky, kx)
    int koff[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = ks * 16 + 2 * t4 + (j & 1) + 8 * (j >> 1);
            const int ci = k / 9, r = k - 9 * ci;
            koff[ks][j] = k < 27 ? ci * (kIn * kInPitch) + (r / 3) * kInPitch + (r % 3) + 1 : 1;
        }

    const int tiles_x = (W2 + kTile - 1) / kTile, tiles_y = (H2 + kTile - 1) / kTile;
    const long ntiles = (long)p.B * tiles_y * tiles_x;
    long long ph[5] = {0, 0, 0, 0, 0};
    const bool prof = p.dbg && blockIdx.x == 0;
    long long c_prev = prof ? clock64() : 0;
    auto mark = [&](int i) { if (prof) { const long long c = clock64(); ph[i] += c - c_prev; c_prev = c; } };
    // The patch of a tile: rows [4 oy0 - 3, + 35) x columns [4 ox0 - 3, + 35) of the three input planes, zero outside the image.  Aligned images
    // (W % 4 == 0, 8-byte aligned base) take it by 8-byte cp.async -- issued for the NEXT tile as soon as conv1 has read this one, so the copies
    // land under conv2; anything else is loaded element-wise.
    const bool aligned = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0;
    const uint32_t s_in32 = sb;
    auto load_patch = [&](long tile) {
        const int b = (int)(tile / (tiles_y * tiles_x));
        const int tr = (int)(tile - (long)b * tiles_y * tiles_x);
        const int yi_0 = 4 * (tr / tiles_x) * kTile - 3, xi_0 = 4 * (tr % tiles_x) * kTile - 3;
        const unsigned short* xb = reinterpret_cast<const unsigned short*>(x) + (size_t)b * 3 * H * W;
        if (aligned) {
            const int xw0 = xi_0 - 1;                         // multiple of 4
            for (int i = tid; i < 3 * kIn * 10; i += blockDim.x) {
                const int r = i / 10, pc = i - r * 10;
                const int ci = r / kIn, iy = r - ci * kIn, yy = yi_0 + iy, xx = xw0 + 4 * pc;
                const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;      // (a piece is entirely inside or outside: the edges are multiples of 4)
                const unsigned short* src = ok ? xb + ((size_t)ci * H + yy) * W + xx : xb;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s_in32 + (uint32_t)((ci * (kIn * kInPitch) + iy * kInPitch + 4 * pc) * 2)), "l"(src),
                             "r"(ok ? 8 : 0) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            const int xa = xi_0 + lane, xc = xa + 32;
            const bool oka = xa >= 0 && xa < W, okc = lane < kIn - 32 && xc >= 0 && xc < W;
            for (int r0 = warp; r0 < 3 * kIn; r0 += 4 * nwarps) {   // a warp per patch row, lanes along x, four rows in flight per warp
                unsigned short va[4], vc[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = r0 + j * nwarps;
                    const int ci = r / kIn, yy = yi_0 + (r - ci * kIn);
                    const bool rowok = r < 3 * kIn && yy >= 0 && yy < H;
                    const unsigned short* src = xb + ((size_t)ci * H + (rowok ? yy : 0)) * W + xi_0;
                    va[j] = (rowok && oka) ? __ldg(src + lane) : (unsigned short)0;
                    vc[j] = (rowok && okc) ? __ldg(src + lane + 32) : (unsigned short)0;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = r0 + j * nwarps;
                    if (r < 3 * kIn) {
                        const int ci = r / kIn, iy = r - ci * kIn;
                        unsigned short* dst = s_in + ci * (kIn * kInPitch) + iy * kInPitch + 1;
                        dst[lane] = va[j];
                        if (lane < kIn - 32) dst[lane + 32] = vc[j];
                    }
                }
            }
        }
    };
    __syncthreads();   // the weight image is written
    if ((long)blockIdx.x < ntiles) load_patch(blockIdx.x);
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = (int)(tile / (tiles_y * tiles_x));
        const int tr = (int)(tile - (long)b * tiles_y * tiles_x);
        const int oy0 = (tr / tiles_x) * kTile, ox0 = (tr % tiles_x) * kTile;
        const int y1_0 = 2 * oy0 - 1, x1_0 = 2 * ox0 - 1;           // intermediate coordinates of the tile's first row / column
        mark(0);
        // ---- 1. this tile's patch has been requested (before the loop, or under the previous tile's conv2)
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();   // ... and is visible to every warp; the previous tile's conv2 has read the intermediate tile
        mark(1);
        mark(2);
        // ---- 2. conv1 + GELU -> intermediate tile
        for (int mt = warp; mt < (kMid * kMid + 15) / 16; mt += nwarps) {   // (half-tile work items balance the warps better but pay the A gather twice: slower)
            int base[2], py[2], px[2];
            bool live[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int s = mt * 16 + g + 8 * h;
                live[h] = s < kMid * kMid;
                const int sc = live[h] ? s : kMid * kMid - 1;
                py[h] = sc / kMid; px[h] = sc - py[h] * kMid;
                base[h] = 2 * py[h] * kInPitch + 2 * px[h];
            }
            float acc[NT1][4];
#pragma unroll
            for (int nt = 0; nt < NT1; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t a[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {   // a0: row g, k 2t..; a1: row g + 8; a2: row g, k + 8; a3: row g + 8, k + 8
                    const int h = j & 1, kk = (j >> 1) * 2;
                    const uint32_t lo = s_in[base[h] + koff[ks][kk]], hi = s_in[base[h] + koff[ks][kk + 1]];
                    a[j] = lo | (hi << 16);
                }
#pragma unroll
                for (int nt = 0; nt < NT1; ++nt) mma16816<T>(acc[nt], a, bw1[nt][ks][0], bw1[nt][ks][1]);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!live[h]) continue;
                const int y1 = y1_0 + py[h], x1 = x1_0 + px[h];
                const bool inside = y1 >= 0 && y1 < H1 && x1 >= 0 && x1 < W1;
                const uint32_t dst = s_mid + (uint32_t)(((py[h] * 2 + (px[h] & 1)) * kHalf + (px[h] >> 1)) * MIDP) + (uint32_t)(4 * t4);
#pragma unroll
                for (int nt = 0; nt < NT1; ++nt) {
                    uint32_t w = 0u;
                    if (inside) {
                        const float2 v = gelu_fit2(H16<T>::round2(__fadd2_rn(make_float2(acc[nt][2 * h], acc[nt][2 * h + 1]), make_float2(bias1[nt][0], bias1[nt][1]))));
                        w = H16<T>::pack(v.x, v.y);
                    }
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst + (uint32_t)(nt * 16)), "r"(w) : "memory");
                }
            }
        }
        __syncthreads();   // conv1 is done with the patch and the intermediate tile is complete
        mark(3);
        if (tile + gridDim.x < ntiles) load_patch(tile + gridDim.x);   // the next tile's patch lands under conv2
        // ---- 3. conv2: warp = (pair of n-tiles, half of the m-tiles)
        {
            const int np = warp >> 1, mh = warp & 1;       // n-tiles 2 np, 2 np + 1; m-tiles 2 mh, 2 mh + 1 (m-tile = two output rows)
            float acc[2][2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int n = 0; n < 2; ++n) { acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f; }
            // ldmatrix row addresses: A: lane -> (matrix = lane / 8: rows 0-7 | 8-15, k 0-7 | 8-15), row = lane % 8 = output column
            const int arow = lane & 7, amat = lane >> 3;
            const int a_oy = (amat & 1), a_khalf = (amat >> 1);   // rows 8-15 of an m-tile are its second output row
            // B: [n][k] rows: matrices (n 0-7, k 0-7), (n 0-7, k 8-15), (n 8-15, k 0-7), (n 8-15, k 8-15)
            const uint32_t b_addr0 = s_w2 + (uint32_t)((np * 16 + (amat >> 1) * 8 + arow) * W2P) + (uint32_t)((amat & 1) * 16);
            // intermediate-tile position of (output row oyl, column arow) under tap (ky, kx): row 2 oyl + ky, column 2 arow + kx, i.e.
            // parity plane kx & 1, plane column arow + (kx >> 1): the tap offset does not depend on the lane
            uint32_t a_addr0[2];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int oyl = 2 * (2 * mh + m) + a_oy;
                a_addr0[m] = s_mid + (uint32_t)(((2 * oyl) * 2 * kHalf + arow) * MIDP) + (uint32_t)(a_khalf * 16);
            }
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap - 3 * ky;
                const uint32_t tap_off = (uint32_t)(((ky * 2 + (kx & 1)) * kHalf + (kx >> 1)) * MIDP);
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    uint32_t bf[4];
                    ldsm4(bf, b_addr0 + (uint32_t)((tap * C1P + kk * 16) * 2));
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        uint32_t af[4];
                        ldsm4(af, a_addr0[m] + tap_off + (uint32_t)(kk * 32));
                        mma16816<T>(acc[m][0], af, bf[0], bf[1]);
                        mma16816<T>(acc[m][1], af, bf[2], bf[3]);
                    }
                }
            }
            // epilogue: + bias -> 16-bit -> NCHW
            const size_t plane = (size_t)H2 * W2;
            const int ox = ox0 + g;
            unsigned short* ob = reinterpret_cast<unsigned short*>(out) + (size_t)b * C2 * plane + ox;
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int n = 0; n < 2; ++n)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int oy = oy0 + 2 * (2 * mh + m) + (e >> 1);
                        const int co = (2 * np + n) * 8 + 2 * t4 + (e & 1);
                        if (co < C2 && oy < H2 && ox < W2) ob[(size_t)co * plane + (size_t)oy * W2] = H16<T>::bits(acc[m][n][e] + s_b2[co]);
                    }
        }
        mark(4);
    }
    if (prof && tid == 0)
        printf("stem CTA 0 warp 0 clocks: wait-for-others %lld, patch -> smem %lld, next patch loads issued %lld, conv1 + GELU %lld, conv2 + stores %lld\n",
               ph[0], ph[1], ph[2], ph[3], ph[4]);
}

}  // namespace

int stem_make_plan(StemPlan& p, int B, int H, int W, int C1, int C2, int dtype, int num_sms) {
    if (!(dtype == 1 || dtype == 2) || B < 1 || H < 1 || W < 1 || C1 < 1 || C1 > 48 || C2 < 1 || C2 > 128) return 1;
    p = StemPlan{};
    p.B = B; p.H = H; p.W = W; p.C1 = C1; p.C2 = C2; p.dtype = dtype;
    p.H1 = (H - 1) / 2 + 1; p.W1 = (W - 1) / 2 + 1;
    p.H2 = (p.H1 - 1) / 2 + 1; p.W2 = (p.W1 - 1) / 2 + 1;
    p.C1P = C1 <= 32 ? 32 : 48;
    p.C2P = (C2 + 15) / 16 * 16;
    p.threads = 64 * (p.C2P / 16);       // a warp per (pair of n-tiles, half of the m-tiles)
    const int midp = p.C1P * 2 + 16, w2p = 9 * p.C1P * 2 + 16;
    p.offMid = (3 * kIn * kInPitch * 2 + 127) / 128 * 128;
    p.offW2 = (p.offMid + kMid * 2 * kHalf * midp + 127) / 128 * 128;
    p.offB2 = p.offW2 + p.C2P * w2p;
    p.smem_bytes = p.offB2 + p.C2P * 4;
    if (p.smem_bytes > 227 * 1024 || p.threads > (p.C1P == 32 ? 256 : 320)) return 1;   // (the kernel variants' launch bounds)
    if (const char* e = getenv("RECNEXT_STEM_DBG")) p.dbg = atoi(e);
    p.grid = num_sms;        // x resident CTAs per SM, set at launch from the occupancy of the kernel variant
    return 0;
}

template <typename T, int KS>
static cudaError_t launch_one(const StemPlan& p, const void* x, const void* w1p, const float* b1p, const void* w2p, const float* b2p, void* out, cudaStream_t st) {
    static DeviceOnce configured = {};
    const cudaError_t e = rc_once_per_device(configured, [] {
        return cudaFuncSetAttribute(recnext_stem_kernel<T, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (e != cudaSuccess) return e;
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, recnext_stem_kernel<T, KS>, p.threads, (size_t)p.smem_bytes) != cudaSuccess || occ < 1) occ = 1;
    const long ntiles = (long)p.B * ((p.H2 + kTile - 1) / kTile) * ((p.W2 + kTile - 1) / kTile);
    long grid = (long)p.grid * occ;
    if (grid > ntiles) grid = ntiles;
    recnext_stem_kernel<T, KS><<<(unsigned)grid, p.threads, p.smem_bytes, st>>>(p, reinterpret_cast<const T*>(x), reinterpret_cast<const T*>(w1p), b1p,
                                                                        reinterpret_cast<const T*>(w2p), b2p, reinterpret_cast<T*>(out));
    return cudaGetLastError();
}

cudaError_t stem_launch(const StemPlan& p, const void* x, const void* w1p, const float* b1p, const void* w2p, const float* b2p, void* out, cudaStream_t st) {
    if (p.dtype == 1) return p.C1P == 32 ? launch_one<__nv_bfloat16, 2>(p, x, w1p, b1p, w2p, b2p, out, st) : launch_one<__nv_bfloat16, 3>(p, x, w1p, b1p, w2p, b2p, out, st);
    return p.C1P == 32 ? launch_one<__half, 2>(p, x, w1p, b1p, w2p, b2p, out, st) : launch_one<__half, 3>(p, x, w1p, b1p, w2p, b2p, out, st);
}

}  // namespace recnext
