// mfwd.cuh — TENSOR-CORE fused RecConv forward for 16-bit activations, k = 5 (see mplan.h for the formulation).
// Reference semantics: model/recnext.py:24-34 under autocast — every conv output, every `f + x` and every
// interpolate result is rounded to the activation dtype, accumulation is fp32.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "recconv_device.cuh"  // mbarrier / cp.async.bulk wrappers, KernelArgs, IdxLam tables
#include "mplan.h"

namespace recnext {

template <typename T> struct MmaT;
template <> struct MmaT<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
        __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    static __device__ __forceinline__ float2 unpack(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
    static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
    static __device__ __forceinline__ void mma16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    static __device__ __forceinline__ void mma8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
    }
};
template <> struct MmaT<__half> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
        __half2 v = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
    static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
    static __device__ __forceinline__ void mma16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    static __device__ __forceinline__ void mma8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
    }
};

__device__ __forceinline__ void m_ldsm4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void m_ldsm2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t m_lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void m_sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void m_sts16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;\n" ::"r"(addr), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ uint32_t m_lds16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(v) : "r"(addr));
    return v;
}

// one padded level buffer of one plane (shared-memory byte addresses)
struct MBuf {
    uint32_t base, parDelta;
    int pitchB, H, W;
    __device__ __forceinline__ uint32_t row(int r) const { return base + (uint32_t)(r & 1) * parDelta + (uint32_t)(r >> 1) * (uint32_t)pitchB; }
};

// ---------------------------------------------------------------------------------------------------------
// 16 output rows [i0, i0+16) x NTC n-tiles [q0, q0+NTC) of a stride-1 5x5 depthwise conv of buffer `in`.
// bfr: this lane's column of the channel's Toeplitz fragment table for this conv (register (r, j) at bfr[(2r+j)*32]).
// epi(q, acc): acc[0..1] = (row i0 + 2(lane/4), columns 8q + 2(lane%4) + {0,1}), acc[2..3] = same columns of the next row.
// ---------------------------------------------------------------------------------------------------------
template <typename T, int NTC, class Epi>
__device__ __forceinline__ void m_conv_s1_tile(const MBuf& in, int i0, int q0, int nt, const uint32_t* __restrict__ bfr, float bias,
                                               int lane, Epi epi) {
    float acc[NTC][4];
#pragma unroll
    for (int q = 0; q < NTC; ++q) { acc[q][0] = bias; acc[q][1] = bias; acc[q][2] = bias; acc[q][3] = bias; }
    // fragment rows 0..7 are image rows i0 + 0, 2, .., 14 and fragment rows 8..15 are i0 + 1, 3, .., 15: the 8 rows of
    // every 8x8 ldmatrix phase then have ONE parity for every filter row r, i.e. they are consecutive rows of one
    // parity array (odd chunk pitch): conflict free.  A thread ends up with the 2x2 block (rows i0 + 2g + {0,1}).
    const int lrow = 2 * (lane & 7) + ((lane >> 3) & 1), lkb = lane >> 4;
    const int rowMax = in.H + 3;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        int rr = i0 + lrow + r;
        rr = rr < rowMax ? rr : rowMax;
        const uint32_t arow = in.row(rr) + (uint32_t)q0 * 16u;
        uint32_t A[NTC + 1][2];
#pragma unroll
        for (int kb = 0; kb + 1 <= NTC; kb += 2)
            if (q0 + kb <= nt) m_ldsm4(A[kb][0], A[kb][1], A[kb + 1][0], A[kb + 1][1], arow + (uint32_t)(kb + lkb) * 16u);
        if (((NTC + 1) & 1) != 0) {
            if (q0 + NTC <= nt) m_ldsm2(A[NTC][0], A[NTC][1], arow + (uint32_t)NTC * 16u);
        }
        const uint32_t b0 = bfr[(2 * r) * 32], b1 = bfr[(2 * r + 1) * 32];
#pragma unroll
        for (int q = 0; q < NTC; ++q)
            if (q0 + q < nt) MmaT<T>::mma16(acc[q], A[q][0], A[q][1], A[q + 1][0], A[q + 1][1], b0, b1);
    }
#pragma unroll
    for (int q = 0; q < NTC; ++q)
        if (q0 + q < nt) epi(q0 + q, acc[q]);
}

// 16 output rows x NTC n-tiles of the stride-2 5x5 depthwise conv (`down`): input rows 2i + r, columns 16q + k, k < 24.
// bfr: register (r, j) at bfr[(4r+j)*32], j = 0,1 -> k 0..15 (m16n8k16), j = 2 -> k 16..23 (m16n8k8).
template <typename T, int NTC, class Epi>
__device__ __forceinline__ void m_conv_s2_tile(const MBuf& in, int i0, int q0, int nt, const uint32_t* __restrict__ bfr, float bias,
                                               int lane, Epi epi) {
    float acc[NTC][4];
#pragma unroll
    for (int q = 0; q < NTC; ++q) { acc[q][0] = bias; acc[q][1] = bias; acc[q][2] = bias; acc[q][3] = bias; }
    const int lrow = lane & 15, lkb = lane >> 4;
    const int rowMax = in.H + 3;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        int rr = 2 * (i0 + lrow) + r;
        rr = rr < rowMax ? rr : rowMax;
        const uint32_t arow = in.row(rr) + (uint32_t)q0 * 32u;
        uint32_t A[2 * NTC + 1][2];
#pragma unroll
        for (int p = 0; p < NTC; ++p)
            if (q0 + p <= nt) m_ldsm4(A[2 * p][0], A[2 * p][1], A[2 * p + 1][0], A[2 * p + 1][1], arow + (uint32_t)(2 * p + lkb) * 16u);
        if (q0 + NTC <= nt) m_ldsm2(A[2 * NTC][0], A[2 * NTC][1], arow + (uint32_t)(2 * NTC) * 16u);
        const uint32_t b0 = bfr[(4 * r) * 32], b1 = bfr[(4 * r + 1) * 32], b2 = bfr[(4 * r + 2) * 32];
#pragma unroll
        for (int q = 0; q < NTC; ++q)
            if (q0 + q < nt) {
                MmaT<T>::mma16(acc[q], A[2 * q][0], A[2 * q][1], A[2 * q + 1][0], A[2 * q + 1][1], b0, b1);
                MmaT<T>::mma8(acc[q], A[2 * q + 2][0], A[2 * q + 2][1], b2);
            }
    }
#pragma unroll
    for (int q = 0; q < NTC; ++q)
        if (q0 + q < nt) epi(q0 + q, acc[q]);
}

template <typename T, bool S2, class Epi>
__device__ __forceinline__ void m_conv_rows(const MBuf& in, int ntc, int i0, int nt, const uint32_t* __restrict__ bfr, float bias, int lane,
                                            Epi epi) {
    if (S2) {
        switch (ntc) {
            case 1: m_conv_s2_tile<T, 1>(in, i0, 0, nt, bfr, bias, lane, epi); break;
            case 2: m_conv_s2_tile<T, 2>(in, i0, 0, nt, bfr, bias, lane, epi); break;
            case 4: m_conv_s2_tile<T, 4>(in, i0, 0, nt, bfr, bias, lane, epi); break;
            default: for (int q0 = 0; q0 < nt; q0 += 7) m_conv_s2_tile<T, 7>(in, i0, q0, nt, bfr, bias, lane, epi); break;
        }
    } else {
        switch (ntc) {
            case 1: m_conv_s1_tile<T, 1>(in, i0, 0, nt, bfr, bias, lane, epi); break;
            case 2: m_conv_s1_tile<T, 2>(in, i0, 0, nt, bfr, bias, lane, epi); break;
            case 4: m_conv_s1_tile<T, 4>(in, i0, 0, nt, bfr, bias, lane, epi); break;
            default: for (int q0 = 0; q0 < nt; q0 += 7) m_conv_s1_tile<T, 7>(in, i0, q0, nt, bfr, bias, lane, epi); break;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct MTeam {
    const MPlan& pl;
    unsigned char* smem;
    unsigned char* tsm;     // team slice
    int team, wt, lane, tl;
    __device__ __forceinline__ void sync() const {
        if (pl.TW == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(pl.team_lanes) : "memory");
    }
    __device__ __forceinline__ MBuf buf(int g, int l) const {
        MBuf b;
        const MLevel& lv = pl.lv[l];
        b.base = rc_smem_u32(tsm + (long)g * pl.plane_bytes + lv.off);
        b.parDelta = (uint32_t)lv.parDelta;
        b.pitchB = lv.pitchB; b.H = lv.H; b.W = lv.W;
        return b;
    }
    __device__ __forceinline__ uint32_t tbuf(int g) const { return rc_smem_u32(tsm + (long)g * pl.plane_bytes + pl.offT); }
    __device__ __forceinline__ const uint32_t* frag(int g, int reg0) const {
        return reinterpret_cast<const uint32_t*>(smem + pl.smFrag) + ((long)g * pl.nregs + reg0) * 32 + lane;
    }
    __device__ __forceinline__ float bias(int g, int slot) const {
        return pl.has_bias ? reinterpret_cast<const float*>(smem + pl.smBias)[g * (pl.L + 2) + slot] : 0.f;
    }
};

// Toeplitz fragments + biases of channel group cg -> shared table (all threads of the CTA)
template <typename T>
__device__ __forceinline__ void m_build_frags(const MPlan& pl, const KernelArgs& a, unsigned char* smem, int cg, int tid, int nthreads) {
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + pl.smFrag);
    const int per = pl.nregs * 32;
    for (int idx = tid; idx < pl.G * per; idx += nthreads) {
        const int p = idx / per, rem = idx - p * per;
        const int reg = rem >> 5, ln = rem & 31;
        const int gq = ln >> 2, t4 = ln & 3;
        const long ch = (long)cg * pl.G + p;
        float v[2] = {0.f, 0.f};
        if (reg < 20) {
            if (pl.L > 0 && (reg & 3) != 3) {
                const int r = reg >> 2, kbase = 2 * t4 + 8 * (reg & 3);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int s = kbase + e - 2 * gq;
                    if (s >= 0 && s <= 4) v[e] = rc_load_param(a.w[0], pl.wdtype, ch * 25 + r * 5 + s);
                }
            }
        } else {
            const int rg = reg - 20, j = rg / 10, r = (rg % 10) >> 1, kbase = 2 * t4 + 8 * (rg & 1);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int s = kbase + e - gq;
                if (s >= 0 && s <= 4) v[e] = rc_load_param(a.w[1 + j], pl.wdtype, ch * 25 + r * 5 + s);
            }
        }
        tab[idx] = MmaT<T>::pack(v[0], v[1]);
    }
    if (pl.has_bias) {
        float* bt = reinterpret_cast<float*>(smem + pl.smBias);
        for (int idx = tid; idx < pl.G * (pl.L + 2); idx += nthreads) {
            const int p = idx / (pl.L + 2), slot = idx - p * (pl.L + 2);
            float v = 0.f;
            if (a.b[slot] && !(slot == 0 && pl.L == 0)) v = rc_load_param(a.b[slot], pl.wdtype, (long)cg * pl.G + p);
            bt[idx] = v;
        }
    }
}

// raw planes (dense, element type T, generic address: shared after a bulk copy, else global) -> padded level 0
template <typename T>
__device__ __forceinline__ void m_repack(const MTeam<T>& tm, const T* __restrict__ src) {
    const MPlan& pl = tm.pl;
    const int H = pl.H, W = pl.W;
    const int LW = 1 << pl.rp_shift, RG = pl.team_lanes >> pl.rp_shift;
    const int j = tm.tl & (LW - 1), rg = tm.tl >> pl.rp_shift;
    if ((W & 1) == 0) {
        const int CP = W >> 1;
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
        for (int g = 0; g < pl.G; ++g) {
            const MBuf b = tm.buf(g, 0);
            for (int i = rg; i < H; i += RG) {
                const uint32_t drow = b.row(i + 2) + 4u;
                const uint32_t* srow = s32 + ((long)g * H + i) * CP;
                for (int jj = j; jj < CP; jj += LW) m_sts32(drow + 4u * jj, srow[jj]);
            }
        }
    } else {
        const unsigned short* s16 = reinterpret_cast<const unsigned short*>(src);
        for (int g = 0; g < pl.G; ++g) {
            const MBuf b = tm.buf(g, 0);
            for (int i = rg; i < H; i += RG) {
                const uint32_t drow = b.row(i + 2) + 4u;
                const unsigned short* srow = s16 + ((long)g * H + i) * W;
                for (int jj = j; jj < W; jj += LW) m_sts16(drow + 2u * jj, srow[jj]);
            }
        }
    }
}

// s_{l-1} = round(x_{l-1} + round(interpolate(t_l)))   (model/recnext.py:33 and the next `f + x`), table driven
template <typename T>
__device__ __forceinline__ void m_up_add(const MTeam<T>& tm, int l) {
    const MPlan& pl = tm.pl;
    const MLevel& ls = pl.lv[l];
    const MLevel& ld = pl.lv[l - 1];
    const IdxLam* ytab = reinterpret_cast<const IdxLam*>(tm.smem + pl.smTab + ls.tabY);
    const IdxLam* xtab = reinterpret_cast<const IdxLam*>(tm.smem + pl.smTab + ls.tabX);
    const int LW = 1 << ls.up_shift, RG = pl.team_lanes >> ls.up_shift;
    const int j = tm.tl & (LW - 1), rg = tm.tl >> ls.up_shift;
    const int CP = (ld.W + 1) >> 1;
    const bool nearest = pl.mode == 1;
    for (int jj = j; jj < CP; jj += LW) {
        const int c0 = 2 * jj, c1 = c0 + 1;
        const bool v1 = c1 < ld.W;
        const IdxLam tx0 = xtab[c0], tx1 = xtab[v1 ? c1 : c0];
        // byte offsets inside a T row (interior at element 2)
        const uint32_t xa0 = 2u * (tx0.i0 + 2), xb0 = 2u * (tx0.i0 + ((!nearest && tx0.i0 < ls.W - 1) ? 1 : 0) + 2);
        const uint32_t xa1 = 2u * (tx1.i0 + 2), xb1 = 2u * (tx1.i0 + ((!nearest && tx1.i0 < ls.W - 1) ? 1 : 0) + 2);
        const float lx0 = tx0.lam, lx1 = tx1.lam, hx0 = 1.f - lx0, hx1 = 1.f - lx1;
        for (int g = 0; g < pl.G; ++g) {
            const MBuf b = tm.buf(g, l - 1);
            const uint32_t Tb = tm.tbuf(g);
            for (int i = rg; i < ld.H; i += RG) {
                const IdxLam ty = ytab[i];
                const uint32_t t0 = Tb + (uint32_t)ty.i0 * ls.tpB;
                float u0, u1;
                if (nearest) {
                    const uint32_t w0 = m_lds16(t0 + xa0), w1 = m_lds16(t0 + xa1);
                    u0 = MmaT<T>::unpack(w0).x; u1 = MmaT<T>::unpack(w1).x;
                } else {
                    const uint32_t t1 = t0 + ((ty.i0 < ls.H - 1) ? (uint32_t)ls.tpB : 0u);
                    const float ly = ty.lam, hy = 1.f - ly;
                    const float a00 = MmaT<T>::unpack(m_lds16(t0 + xa0)).x, a01 = MmaT<T>::unpack(m_lds16(t0 + xb0)).x;
                    const float a10 = MmaT<T>::unpack(m_lds16(t1 + xa0)).x, a11 = MmaT<T>::unpack(m_lds16(t1 + xb0)).x;
                    const float b00 = MmaT<T>::unpack(m_lds16(t0 + xa1)).x, b01 = MmaT<T>::unpack(m_lds16(t0 + xb1)).x;
                    const float b10 = MmaT<T>::unpack(m_lds16(t1 + xa1)).x, b11 = MmaT<T>::unpack(m_lds16(t1 + xb1)).x;
                    u0 = MmaT<T>::rnd(hy * (hx0 * a00 + lx0 * a01) + ly * (hx0 * a10 + lx0 * a11));
                    u1 = MmaT<T>::rnd(hy * (hx1 * b00 + lx1 * b01) + ly * (hx1 * b10 + lx1 * b11));
                }
                const uint32_t daddr = b.row(i + 2) + 4u + 4u * jj;
                const float2 s = MmaT<T>::unpack(m_lds32(daddr));
                m_sts32(daddr, MmaT<T>::pack(s.x + u0, v1 ? s.y + u1 : 0.f));
            }
        }
    }
}

// exact-2x bilinear (align_corners=False): fixed 0.75 / 0.25 stencil with the source index clamped at the borders
// (ATen: even destination 2a -> sources a-1 (.25), a (.75); odd 2a+1 -> a (.75), a+1 (.25)).  A lane owns one column
// pair (2a, 2a+1) of level l-1 = source columns a-1, a, a+1 (replicate border kept in T) and walks down source rows.
template <typename T>
__device__ __forceinline__ void m_up2x_add(const MTeam<T>& tm, int l) {
    const MPlan& pl = tm.pl;
    const MLevel& ls = pl.lv[l];
    const int LW = 1 << ls.up_shift, RG = pl.team_lanes >> ls.up_shift;
    const int j = tm.tl & (LW - 1), rg = tm.tl >> ls.up_shift;
    const int Hs = ls.H, Ws = ls.W;
    const int rpg = (Hs + RG - 1) / RG;          // source rows per row group
    const int m0 = rg * rpg, m1 = (m0 + rpg) < Hs ? (m0 + rpg) : Hs;
    if (m0 >= m1) return;
    for (int a = j; a < Ws; a += LW) {
        for (int g = 0; g < pl.G; ++g) {
            const MBuf b = tm.buf(g, l - 1);
            const uint32_t Tc = tm.tbuf(g) + 2u * (a + 1);   // element a - 1 (interior at 2)
            auto hrow = [&](int m, float& h0, float& h1) {
                const uint32_t p = Tc + (uint32_t)m * ls.tpB;
                const float ta = MmaT<T>::unpack(m_lds16(p)).x, tb = MmaT<T>::unpack(m_lds16(p + 2)).x, tc = MmaT<T>::unpack(m_lds16(p + 4)).x;
                h0 = 0.25f * ta + 0.75f * tb;
                h1 = 0.75f * tb + 0.25f * tc;
            };
            float p0, p1, c0, c1, n0, n1;
            hrow(m0 > 0 ? m0 - 1 : 0, p0, p1);
            hrow(m0, c0, c1);
            for (int m = m0; m < m1; ++m) {
                hrow(m + 1 < Hs ? m + 1 : Hs - 1, n0, n1);
                const uint32_t d0 = b.row(2 * m + 2) + 4u + 4u * a, d1 = b.row(2 * m + 3) + 4u + 4u * a;
                const float2 s0 = MmaT<T>::unpack(m_lds32(d0)), s1 = MmaT<T>::unpack(m_lds32(d1));
                const float e0 = MmaT<T>::rnd(0.25f * p0 + 0.75f * c0), e1 = MmaT<T>::rnd(0.25f * p1 + 0.75f * c1);
                const float f0 = MmaT<T>::rnd(0.75f * c0 + 0.25f * n0), f1 = MmaT<T>::rnd(0.75f * c1 + 0.25f * n1);
                m_sts32(d0, MmaT<T>::pack(s0.x + e0, s0.y + e1));
                m_sts32(d1, MmaT<T>::pack(s1.x + f0, s1.y + f1));
                p0 = c0; p1 = c1; c0 = n0; c1 = n1;
            }
        }
    }
}

// MAXT = 320: up to 10 warps with ~200 registers (wide planes: 7 n-tiles per pass); MAXT = 512: up to 16 warps at 128
template <typename T, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) recconv_mfwd_kernel(const __grid_constant__ MPlan pl, const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int team = warp / pl.TW, wt = warp - team * pl.TW;
    MTeam<T> tm{pl, smem, smem + pl.smTeams + (long)team * pl.team_bytes, team, wt, lane, wt * 32 + lane};
    const int L = pl.L, G = pl.G;

    // ---- CTA init: zero the team slices (the borders of the padded buffers stay zero), interpolation tables, mbarriers
    {
        uint4* z = reinterpret_cast<uint4*>(smem + pl.smTeams);
        const int n16 = pl.NTEAM * pl.team_bytes / 16;
        const uint4 zero = {0u, 0u, 0u, 0u};
        for (int i = tid; i < n16; i += blockDim.x) z[i] = zero;
        for (int l = 1; l <= L; ++l) {
            rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + pl.smTab + pl.lv[l].tabY), pl.lv[l].H, pl.lv[l - 1].H, pl.mode, tid, blockDim.x);
            rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + pl.smTab + pl.lv[l].tabX), pl.lv[l].W, pl.lv[l - 1].W, pl.mode, tid, blockDim.x);
        }
    }
    const uint32_t bar = rc_smem_u32(smem + pl.smBar + 8 * team);
    uint32_t phase = 0;
    if (pl.use_tma && tm.tl == 0) rc_mbar_init(bar, 1);
    rc_fence_proxy_async();  // the zero fill above also touched the bulk-copy buffers
    __syncthreads();

    const long plane_elems = (long)pl.H * pl.W;
    const T* gx = reinterpret_cast<const T*>(a.x);
    T* gy = reinterpret_cast<T*>(a.out);
    const long total = (long)pl.n_cg * pl.B;
    const long start = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1) / gridDim.x;
    if (start >= end) return;
    long my = start + team;
    unsigned char* raw = tm.tsm + pl.off_raw;
    auto plane0_of = [&](long idx) { const long cg = idx / pl.B, n = idx - cg * pl.B; return (n * pl.C + cg * G) * plane_elems; };
    if (pl.use_tma && my < end && tm.tl == 0) {
        rc_mbar_expect_tx(bar, (uint32_t)pl.raw_bytes);
        rc_bulk_g2s(raw, gx + plane0_of(my), (uint32_t)pl.raw_bytes, bar);
    }
    const int cg_first = (int)(start / pl.B), cg_last = (int)((end - 1) / pl.B);
    for (int cg = cg_first; cg <= cg_last; ++cg) {
        __syncthreads();  // every team is done with the previous channel group's fragments
        m_build_frags<T>(pl, a, smem, cg, tid, blockDim.x);
        __syncthreads();
        const long cg_end = ((long)(cg + 1) * pl.B) < end ? ((long)(cg + 1) * pl.B) : end;
        for (; my < cg_end; my += pl.NTEAM) {
            const long p0 = plane0_of(my);
            // ---- x -> padded level 0
            if (pl.use_tma) {
                while (!rc_mbar_try_wait(bar, phase)) {}
                phase ^= 1u;
                m_repack<T>(tm, reinterpret_cast<const T*>(raw));
                rc_fence_proxy_async();  // order these generic reads of `raw` before the next bulk copy's writes
                tm.sync();
                const long nxt = my + pl.NTEAM;
                if (nxt < end && tm.tl == 0) {
                    rc_mbar_expect_tx(bar, (uint32_t)pl.raw_bytes);
                    rc_bulk_g2s(raw, gx + plane0_of(nxt), (uint32_t)pl.raw_bytes, bar);
                }
            } else {
                m_repack<T>(tm, gx + p0);
                tm.sync();
            }
            // ---- down chain: x_l = down(x_{l-1})   (model/recnext.py:27-29)
            for (int l = 1; l <= L; ++l) {
                const MLevel& lo = pl.lv[l];
                for (int it = wt; it < G * lo.MT; it += pl.TW) {
                    const int g = it / lo.MT, mt = it - g * lo.MT;
                    const MBuf in = tm.buf(g, l - 1);
                    const MBuf out = tm.buf(g, l);
                    const int i0 = mt * 16, Ho = lo.H, Wo = lo.W;
                    m_conv_rows<T, true>(in, lo.ntc, i0, lo.NT, tm.frag(g, 0), tm.bias(g, 0), lane, [&](int q, const float (&acc)[4]) {
                        const int c = 8 * q + 2 * (lane & 3);
                        if (c < Wo) {
                            const bool pair = c + 1 < Wo;
                            const int ia = i0 + (lane >> 2), ib = ia + 8;
                            if (ia < Ho) m_sts32(out.row(ia + 2) + 4u + 2u * c, MmaT<T>::pack(acc[0], pair ? acc[1] : 0.f));
                            if (ib < Ho) m_sts32(out.row(ib + 2) + 4u + 2u * c, MmaT<T>::pack(acc[2], pair ? acc[3] : 0.f));
                        }
                    });
                }
                tm.sync();
            }
            // ---- up pass: t_l = convs[L-l](s_l); s_{l-1} = x_{l-1} + interpolate(t_l)   (model/recnext.py:31-33)
            for (int l = L; l >= 1; --l) {
                const MLevel& lv = pl.lv[l];
                for (int it = wt; it < G * lv.MT; it += pl.TW) {
                    const int g = it / lv.MT, mt = it - g * lv.MT;
                    const MBuf in = tm.buf(g, l);
                    const uint32_t Tb = tm.tbuf(g);
                    const int i0 = mt * 16, Hl = lv.H, Wl = lv.W, tpB = lv.tpB;
                    m_conv_rows<T, false>(in, lv.ntc, i0, lv.NT, tm.frag(g, 20 + 10 * (L - l)), tm.bias(g, 1 + (L - l)), lane,
                                          [&](int q, const float (&acc)[4]) {
                        const int c = 8 * q + 2 * (lane & 3);
                        if (c < Wl) {
                            const int ia = i0 + 2 * (lane >> 2);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int i = ia + h;
                                if (i < Hl) {
                                    const float lo = acc[2 * h], hi = (c + 1 < Wl) ? acc[2 * h + 1] : acc[2 * h];
                                    const uint32_t w = MmaT<T>::pack(lo, hi);
                                    const uint32_t ad = Tb + (uint32_t)i * tpB + 2u * (c + 2);
                                    m_sts32(ad, w);                                          // (c == W-1: the pair's high half is the replicate border)
                                    if (c == 0) m_sts16(ad - 2u, w & 0xffffu);                // left replicate border
                                    if (c + 1 == Wl - 1) m_sts16(ad + 4u, w >> 16);           // right replicate border
                                }
                            }
                        }
                    });
                }
                tm.sync();
                if (lv.exact2x && pl.mode == 0 && !(pl.dbg & 1)) m_up2x_add<T>(tm, l);
                else m_up_add<T>(tm, l);
                tm.sync();
            }
            // ---- y = convs[L](s_0) -> global   (model/recnext.py:34)
            {
                const MLevel& lv = pl.lv[0];
                const int H = pl.H, W = pl.W;
                for (int it = wt; it < G * lv.MT; it += pl.TW) {
                    const int g = it / lv.MT, mt = it - g * lv.MT;
                    const MBuf in = tm.buf(g, 0);
                    T* dst = gy + p0 + (long)g * plane_elems;
                    const int i0 = mt * 16;
                    m_conv_rows<T, false>(in, lv.ntc, i0, lv.NT, tm.frag(g, 20 + 10 * L), tm.bias(g, 1 + L), lane,
                                          [&](int q, const float (&acc)[4]) {
                        const int c = 8 * q + 2 * (lane & 3);
                        if (c < W) {
                            const int ia = i0 + 2 * (lane >> 2);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int i = ia + h;
                                if (i < H) {
                                    const uint32_t w = MmaT<T>::pack(acc[2 * h], acc[2 * h + 1]);
                                    T* d = dst + (long)i * W + c;
                                    if ((W & 1) == 0) *reinterpret_cast<uint32_t*>(d) = w;
                                    else {
                                        *reinterpret_cast<unsigned short*>(d) = (unsigned short)(w & 0xffffu);
                                        if (c + 1 < W) *reinterpret_cast<unsigned short*>(d + 1) = (unsigned short)(w >> 16);
                                    }
                                }
                            }
                        }
                    });
                }
                tm.sync();  // level 0 is rewritten by the next batch's repack
            }
        }
    }
}

template <typename T>
inline cudaError_t m_launch_fwd_t(const MPlan& pl, const KernelArgs& a, cudaStream_t stream) {
    static int configured = 0;  // benign race: idempotent
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(recconv_mfwd_kernel<T, 320>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(recconv_mfwd_kernel<T, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        configured = 1;
    }
    if (pl.threads <= 320) recconv_mfwd_kernel<T, 320><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
    else recconv_mfwd_kernel<T, 512><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
    return cudaGetLastError();
}

inline cudaError_t m_launch_fwd(const MPlan& pl, const KernelArgs& a, cudaStream_t stream) {
    if (pl.dtype == 1) return m_launch_fwd_t<__nv_bfloat16>(pl, a, stream);
    if (pl.dtype == 2) return m_launch_fwd_t<__half>(pl, a, stream);
    return cudaErrorInvalidValue;
}

}  // namespace recnext
