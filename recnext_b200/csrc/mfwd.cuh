// mfwd.cuh — TENSOR-CORE fused RecConv forward for 16-bit activations, k = 5 (see mplan.h for the formulation).
// Reference semantics: model/recnext.py:24-34 under autocast — every conv output, every `f + x` and every
// interpolate result is rounded to the activation dtype, accumulation is fp32.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "recconv_device.cuh"  // mbarrier / cp.async.bulk wrappers, KernelArgs, IdxLam tables
#include "mplan.h"
#include "devcfg.h"
#include <string.h>
#include <type_traits>

namespace recnext {

template <typename T> struct MmaT;
template <> struct MmaT<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
        __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    static __device__ __forceinline__ float2 unpack(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
    static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
    static __device__ __forceinline__ float lo(uint32_t u) { return __uint_as_float(u << 16); }       // element 0 of a pair
    static __device__ __forceinline__ float one(uint32_t h) { return __uint_as_float(h << 16); }      // a zero-extended element
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {                          // round(a + b) per element
        __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
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
    static __device__ __forceinline__ float lo(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u & 0xffffu))); }
    static __device__ __forceinline__ float one(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)h)); }
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {
        __half2 r = __hadd2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
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
// bfr: shared address of this lane's column of the channel's Toeplitz fragments for this conv (register (r, j) at
// bfr + (2r+j)*128).  FULL: all NTC n-tiles exist (no guards: straight-line code, no reconvergence barriers).
// Fragment rows 0..7 are image rows i0 + 0, 2, .., 14 and fragment rows 8..15 are i0 + 1, 3, .., 15: the 8 rows of
// every 8x8 ldmatrix phase then have ONE parity for every filter row r, i.e. they are consecutive rows of one
// parity array (odd chunk pitch): conflict free.
// epi(q, acc): acc[0..1] = (row i0 + 2(lane/4), columns 8q + 2(lane%4) + {0,1}), acc[2..3] = same columns of the next row.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void m_ldsm1(uint32_t& r0, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];\n" : "=r"(r0) : "r"(addr));
}

// row set E_r of an M-tile: the 8 image rows i0 + r + 0, 2, .., 14 (one parity: consecutive rows of one parity array,
// conflict free), k-blocks q0 .. q0 + NB - 1.  Lane L addresses row 2 (L % 8) of k-block L / 8 of each ldmatrix.
template <int NB>
__device__ __forceinline__ void m_load_rowset(uint32_t (&E)[NB], const MBuf& in, int row0, int rowMax, uint32_t colB, int lane) {
    int rr = row0 + 2 * (lane & 7);
    rr = rr < rowMax ? rr : rowMax;
    const uint32_t a = in.row(rr) + colB + (uint32_t)(lane >> 3) * 16u;
#pragma unroll
    for (int kb = 0; kb < NB; kb += 4) {
        if (NB - kb >= 4) m_ldsm4(E[kb], E[kb + 1], E[kb + 2], E[kb + 3], a + (uint32_t)kb * 16u);
        else if (NB - kb == 3) { m_ldsm2(E[kb], E[kb + 1], a + (uint32_t)kb * 16u); m_ldsm1(E[kb + 2], a - (uint32_t)(lane >> 3) * 16u + (uint32_t)(kb + 2) * 16u); }
        else if (NB - kb == 2) m_ldsm2(E[kb], E[kb + 1], a + (uint32_t)kb * 16u);
        else m_ldsm1(E[kb], a - (uint32_t)(lane >> 3) * 16u + (uint32_t)kb * 16u);
    }
}

template <typename T, int NTC, bool FULL, class Epi>
__device__ __forceinline__ void m_conv_s1_tile(const MBuf& in, int i0, int q0, int nt, uint32_t bfr, float bias, int lane, Epi epi) {
    float acc[NTC][4];
#pragma unroll
    for (int q = 0; q < NTC; ++q) { acc[q][0] = bias; acc[q][1] = bias; acc[q][2] = bias; acc[q][3] = bias; }
    // Fragment rows 0..7 are image rows i0 + 0, 2, .., 14 and fragment rows 8..15 are i0 + 1, 3, .., 15.  For filter row r
    // the A operand is therefore (E_r, E_{r+1}) with E_r = image rows i0 + r + {0, 2, .., 14}: the lower half of one
    // filter row is the upper half of the previous one, so an M-tile needs 6 row sets instead of 10 (40 % less
    // shared-memory traffic), each a conflict-free run of consecutive rows of one parity array.
    const int rowMax = in.H + 3;
    const uint32_t colB = (uint32_t)q0 * 16u;
    uint32_t E[2][NTC + 1];
    m_load_rowset<NTC + 1>(E[0], in, i0, rowMax, colB, lane);
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        m_load_rowset<NTC + 1>(E[(r + 1) & 1], in, i0 + r + 1, rowMax, colB, lane);
        const uint32_t b0 = m_lds32(bfr + (2 * r) * 128), b1 = m_lds32(bfr + (2 * r + 1) * 128);
#pragma unroll
        for (int q = 0; q < NTC; ++q)
            if (FULL || q0 + q < nt) MmaT<T>::mma16(acc[q], E[r & 1][q], E[(r + 1) & 1][q], E[r & 1][q + 1], E[(r + 1) & 1][q + 1], b0, b1);
    }
#pragma unroll
    for (int q = 0; q < NTC; ++q)
        if (FULL || q0 + q < nt) epi(q0 + q, acc[q]);
}

// 16 output rows x NTC n-tiles of the stride-2 5x5 depthwise conv (`down`): input rows 2i + r, columns 16q + k, k < 24.
// bfr: register (r, j) at bfr + (4r+j)*128, j = 0,1 -> k 0..15 (m16n8k16), j = 2 -> k 16..23 (m16n8k8).
// epi(q, acc): acc[0..1] = (row i0 + lane/4, columns 8q + 2(lane%4) + {0,1}), acc[2..3] = same columns of row + 8.
template <typename T, int NTC, bool FULL, class Epi>
__device__ __forceinline__ void m_conv_s2_tile(const MBuf& in, int i0, int q0, int nt, uint32_t bfr, float bias, int lane, Epi epi) {
    float acc[NTC][4];
#pragma unroll
    for (int q = 0; q < NTC; ++q) { acc[q][0] = bias; acc[q][1] = bias; acc[q][2] = bias; acc[q][3] = bias; }
    const int lrow = lane & 15;
    const uint32_t lcol = (uint32_t)(2 * q0 + (lane >> 4)) * 16u;
    const int rowMax = in.H + 3;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        int rr = 2 * (i0 + lrow) + r;
        rr = rr < rowMax ? rr : rowMax;
        const uint32_t arow = in.row(rr) + lcol;
        uint32_t A[2 * NTC + 1][2];
#pragma unroll
        for (int p = 0; p < NTC; ++p)
            if (FULL || q0 + p <= nt) m_ldsm4(A[2 * p][0], A[2 * p][1], A[2 * p + 1][0], A[2 * p + 1][1], arow + (uint32_t)(2 * p) * 16u);
        if (FULL || q0 + NTC <= nt) m_ldsm2(A[2 * NTC][0], A[2 * NTC][1], arow - (uint32_t)(lane >> 4) * 16u + (uint32_t)(2 * NTC) * 16u);
        const uint32_t b0 = m_lds32(bfr + (4 * r) * 128), b1 = m_lds32(bfr + (4 * r + 1) * 128), b2 = m_lds32(bfr + (4 * r + 2) * 128);
#pragma unroll
        for (int q = 0; q < NTC; ++q)
            if (FULL || q0 + q < nt) {
                MmaT<T>::mma16(acc[q], A[2 * q][0], A[2 * q][1], A[2 * q + 1][0], A[2 * q + 1][1], b0, b1);
                MmaT<T>::mma8(acc[q], A[2 * q + 2][0], A[2 * q + 2][1], b2);
            }
    }
#pragma unroll
    for (int q = 0; q < NTC; ++q)
        if (FULL || q0 + q < nt) epi(q0 + q, acc[q]);
}

template <typename T, bool S2, int NTC, class Epi>
__device__ __forceinline__ void m_conv_chunks(const MBuf& in, int i0, int nt, uint32_t bfr, float bias, int lane, Epi epi) {
    int q0 = 0;
    for (; q0 + NTC <= nt; q0 += NTC) {
        if (S2) m_conv_s2_tile<T, NTC, true>(in, i0, q0, nt, bfr, bias, lane, epi);
        else m_conv_s1_tile<T, NTC, true>(in, i0, q0, nt, bfr, bias, lane, epi);
    }
    if (NTC > 2 && q0 < nt) {  // ragged tail (NTC <= 2 is only chosen when it equals the n-tile count)
        if (S2) m_conv_s2_tile<T, NTC, false>(in, i0, q0, nt, bfr, bias, lane, epi);
        else m_conv_s1_tile<T, NTC, false>(in, i0, q0, nt, bfr, bias, lane, epi);
    }
}

template <typename T, bool S2, class Epi>
__device__ __forceinline__ void m_conv_rows(const MBuf& in, int ntc, int i0, int nt, uint32_t bfr, float bias, int lane, Epi epi) {
    if (ntc == 1) m_conv_chunks<T, S2, 1>(in, i0, nt, bfr, bias, lane, epi);
    else if (ntc == 2) m_conv_chunks<T, S2, 2>(in, i0, nt, bfr, bias, lane, epi);
    else if (S2 || ntc == 4) m_conv_chunks<T, S2, 4>(in, i0, nt, bfr, bias, lane, epi);  // `down`: at most 4 n-tiles per pass
    else m_conv_chunks<T, S2, 7>(in, i0, nt, bfr, bias, lane, epi);
}

// ---------------------------------------------------------------------------------------------------------
// team context
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct MTeam {
    const MPlan& pl;   // layout (compile-time constants in the specialised kernels)
    const MPlan& rt;   // run-time fields: B, C, n_cg, has_bias, wdtype, use_tma, dbg
    uint32_t smem32, tsm32;   // shared addresses of the CTA's dynamic shared memory and of the team slice
    int team, wt, lane, tl;
    __device__ __forceinline__ void sync() const {
        if (pl.TW == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(pl.team_lanes) : "memory");
    }
    __device__ __forceinline__ MBuf buf(int g, int l) const {
        MBuf b;
        const MLevel& lv = pl.lv[l];
        b.base = l == 0 ? tsm32 + (uint32_t)(g * pl.l0_bytes) : tsm32 + (uint32_t)(pl.off_upper + g * pl.upper_bytes + lv.off);
        b.parDelta = (uint32_t)lv.parDelta;
        b.pitchB = lv.pitchB; b.H = lv.H; b.W = lv.W;
        return b;
    }
    __device__ __forceinline__ uint32_t tbuf(int g, int l) const { return tsm32 + (uint32_t)(pl.off_upper + g * pl.upper_bytes + pl.lv[l].offT); }
    __device__ __forceinline__ uint32_t frag(int g, int reg0) const { return smem32 + (uint32_t)(pl.smFrag + ((g * pl.nregs + reg0) * 32 + lane) * 4); }
    __device__ __forceinline__ float bias(int g, int slot) const {
        float v = 0.f;
        if (rt.has_bias) asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(smem32 + (uint32_t)(pl.smBias + (g * (pl.L + 2) + slot) * 4)));
        return v;
    }
};

// Toeplitz fragments + biases of channel group cg -> shared table (all threads of the CTA)
template <typename T>
__device__ __forceinline__ void m_build_frags(const MPlan& pl, const MPlan& rt, const KernelArgs& a, unsigned char* smem, int cg, int tid, int nthreads) {
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + pl.smFrag);
    const int per = pl.nregs * 32;
    for (int idx = tid; idx < pl.G * per; idx += nthreads) {
        const int p = idx / per, rem = idx - p * per;
        const int reg = rem >> 5, ln = rem & 31;
        const int gq = ln >> 2, t4 = ln & 3;
        const long ch = (long)cg * pl.G + p;
        float v[2] = {0.f, 0.f};
        if (reg < 20) {
            if (pl.L > 0 && a.w[0] != nullptr && (reg & 3) != 3) {
                const int r = reg >> 2, kbase = 2 * t4 + 8 * (reg & 3);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int s = kbase + e - 2 * gq;
                    if (s >= 0 && s <= 4) v[e] = rc_load_param(a.w[0], rt.wdtype, ch * 25 + r * 5 + s);
                }
            }
        } else {
            const int rg = reg - 20, j = rg / 10, r = (rg % 10) >> 1, kbase = 2 * t4 + 8 * (rg & 1);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int s = kbase + e - gq;
                if (s >= 0 && s <= 4 && a.w[1 + j] != nullptr) v[e] = rc_load_param(a.w[1 + j], rt.wdtype, ch * 25 + r * 5 + s);
            }
        }
        tab[idx] = MmaT<T>::pack(v[0], v[1]);
    }
    if (rt.has_bias) {
        float* bt = reinterpret_cast<float*>(smem + pl.smBias);
        for (int idx = tid; idx < pl.G * (pl.L + 2); idx += nthreads) {
            const int p = idx / (pl.L + 2), slot = idx - p * (pl.L + 2);
            float v = 0.f;
            if (a.b[slot] && !(slot == 0 && pl.L == 0)) v = rc_load_param(a.b[slot], rt.wdtype, (long)cg * pl.G + p);
            bt[idx] = v;
        }
    }
}

// raw planes (dense, element type T, generic address: shared after a bulk copy, else global) -> padded level 0.
// A lane owns column pair j of the rows i = rg, rg + RG, ...; the two row parities are walked separately so that
// both the source and the destination address advance by a constant.
template <typename T>
__device__ __forceinline__ void m_repack(const MTeam<T>& tm, const T* __restrict__ src) {
    const MPlan& pl = tm.pl;
    const int H = pl.H, W = pl.W;
    const int LW = 1 << pl.rp_shift, RG = pl.team_lanes >> pl.rp_shift;
    const int j = tm.tl & (LW - 1), rg = tm.tl >> pl.rp_shift;
    const MLevel& l0 = pl.lv[0];
    if ((W & 1) == 0) {
        const int CP = W >> 1;
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
        const uint32_t dstep = (uint32_t)(RG * l0.pitchB);
        const int sstep = 2 * RG * CP;
        for (int g = 0; g < pl.G; ++g) {
            const MBuf b = tm.buf(g, 0);
            for (int jj = j; jj < CP; jj += LW) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int i = rg + h * RG;
                    uint32_t d = b.row(i + 2) + 4u + 4u * jj;
                    const uint32_t* sp = s32 + (g * H + i) * CP + jj;
                    for (; i < H; i += 2 * RG) { m_sts32(d, *sp); d += dstep; sp += sstep; }
                }
            }
        }
    } else {
        const unsigned short* s16 = reinterpret_cast<const unsigned short*>(src);
        for (int g = 0; g < pl.G; ++g) {
            const MBuf b = tm.buf(g, 0);
            for (int i = rg; i < H; i += RG) {
                const uint32_t drow = b.row(i + 2) + 4u;
                const unsigned short* srow = s16 + (g * H + i) * W;
                for (int jj = j; jj < W; jj += LW) m_sts16(drow + 2u * jj, srow[jj]);
            }
        }
    }
}

// variant 2: the external low-resolution operand z (dense planes in global memory) -> T of level 1, with the
// replicate border columns the interpolation stages expect
template <typename T>
__device__ __forceinline__ void m_load_z(const MTeam<T>& tm, const T* __restrict__ z) {
    const MPlan& pl = tm.pl;
    const MLevel& lz = pl.lv[1];
    const int LW = 1 << tm.rt.z_shift, RG = pl.team_lanes >> tm.rt.z_shift;
    const int j = tm.tl & (LW - 1), rg = tm.tl >> tm.rt.z_shift;
    const unsigned short* z16 = reinterpret_cast<const unsigned short*>(z);
    for (int g = 0; g < pl.G; ++g) {
        const uint32_t Tb = tm.tbuf(g, 1);
        for (int i = rg; i < lz.H; i += RG) {
            const uint32_t trow = Tb + (uint32_t)(i * lz.tpB) + 4u;
            const unsigned short* srow = z16 + ((long)g * lz.H + i) * lz.W;
            for (int c = j; c < lz.W; c += LW) {
                const uint32_t v = srow[c];
                m_sts16(trow + 2u * c, v);
                if (c == 0) m_sts16(trow - 2u, v);
                if (c == lz.W - 1) m_sts16(trow + 2u * c + 2u, v);
            }
        }
    }
}

// s_{l-1} = round(x_{l-1} + round(interpolate(t_l)))   (model/recnext.py:33 and the next `f + x`), table driven
template <typename T>
__device__ __forceinline__ void m_up_add(const MTeam<T>& tm, unsigned char* smem, int l) {
    const MPlan& pl = tm.pl;
    const MLevel& ls = pl.lv[l];
    const MLevel& ld = pl.lv[l - 1];
    const IdxLam* ytab = reinterpret_cast<const IdxLam*>(smem + pl.smTab + ls.tabY);
    const IdxLam* xtab = reinterpret_cast<const IdxLam*>(smem + pl.smTab + ls.tabX);
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
            const uint32_t Tb = tm.tbuf(g, l);
            for (int i = rg; i < ld.H; i += RG) {
                const IdxLam ty = ytab[i];
                const uint32_t t0 = Tb + (uint32_t)ty.i0 * ls.tpB;
                uint32_t u;
                if (nearest) {
                    u = m_lds16(t0 + xa0) | (v1 ? (m_lds16(t0 + xa1) << 16) : 0u);
                } else {
                    const uint32_t t1 = t0 + ((ty.i0 < ls.H - 1) ? (uint32_t)ls.tpB : 0u);
                    const float ly = ty.lam, hy = 1.f - ly;
                    const float a00 = MmaT<T>::one(m_lds16(t0 + xa0)), a01 = MmaT<T>::one(m_lds16(t0 + xb0));
                    const float a10 = MmaT<T>::one(m_lds16(t1 + xa0)), a11 = MmaT<T>::one(m_lds16(t1 + xb0));
                    const float b00 = MmaT<T>::one(m_lds16(t0 + xa1)), b01 = MmaT<T>::one(m_lds16(t0 + xb1));
                    const float b10 = MmaT<T>::one(m_lds16(t1 + xa1)), b11 = MmaT<T>::one(m_lds16(t1 + xb1));
                    const float u0 = hy * (hx0 * a00 + lx0 * a01) + ly * (hx0 * a10 + lx0 * a11);
                    const float u1 = hy * (hx1 * b00 + lx1 * b01) + ly * (hx1 * b10 + lx1 * b11);
                    u = MmaT<T>::pack(u0, v1 ? u1 : 0.f);
                }
                const uint32_t daddr = b.row(i + 2) + 4u + 4u * jj;
                m_sts32(daddr, MmaT<T>::add2(m_lds32(daddr), u));
            }
        }
    }
}

// exact-2x bilinear (align_corners=False): fixed 0.75 / 0.25 stencil with the source index clamped at the borders
// (ATen: even destination 2a -> sources a-1 (.25), a (.75); odd 2a+1 -> a (.75), a+1 (.25)).  A lane owns one column
// pair (2a, 2a+1) of level l-1 = source columns a-1, a, a+1 (replicate border kept in T) and walks down its block of
// source rows; destination rows 2m and 2m+1 live in different parity arrays, each advancing by one pitch per source row.
template <typename T>
__device__ __forceinline__ void m_up2x_add(const MTeam<T>& tm, int l) {
    const MPlan& pl = tm.pl;
    const MLevel& ls = pl.lv[l];
    const int LW = 1 << ls.up_shift;
    const int j = tm.tl & (LW - 1), rg = tm.tl >> ls.up_shift;
    const int Hs = ls.H, Ws = ls.W;
    const int rpg = ls.up_rpg;                    // source rows per row group
    const int m0 = rg * rpg, m1 = (m0 + rpg) < Hs ? (m0 + rpg) : Hs;
    if (m0 >= m1) return;
    const uint32_t tpB = (uint32_t)ls.tpB;
    for (int a = j; a < Ws; a += LW) {
        for (int g = 0; g < pl.G; ++g) {
            const MBuf b = tm.buf(g, l - 1);
            const uint32_t Tc = tm.tbuf(g, l) + 2u * (a + 1);   // element a - 1 (interior at 2)
            auto hrow = [&](uint32_t p, float& h0, float& h1) {
                const float ta = MmaT<T>::one(m_lds16(p)), tb = MmaT<T>::one(m_lds16(p + 2)), tc = MmaT<T>::one(m_lds16(p + 4));
                const float q = 0.75f * tb;
                h0 = fmaf(0.25f, ta, q);
                h1 = fmaf(0.25f, tc, q);
            };
            float p0, p1, c0, c1, n0, n1;
            hrow(Tc + (uint32_t)(m0 > 0 ? m0 - 1 : 0) * tpB, p0, p1);
            uint32_t tp = Tc + (uint32_t)m0 * tpB;
            hrow(tp, c0, c1);
            uint32_t d0 = b.row(2 * m0 + 2) + 4u + 4u * a, d1 = b.row(2 * m0 + 3) + 4u + 4u * a;
            for (int m = m0; m < m1; ++m) {
                if (m + 1 < Hs) tp += tpB;
                hrow(tp, n0, n1);
                const float q0 = 0.75f * c0, q1 = 0.75f * c1;
                const uint32_t e = MmaT<T>::pack(fmaf(0.25f, p0, q0), fmaf(0.25f, p1, q1));
                const uint32_t f = MmaT<T>::pack(fmaf(0.25f, n0, q0), fmaf(0.25f, n1, q1));
                m_sts32(d0, MmaT<T>::add2(m_lds32(d0), e));
                m_sts32(d1, MmaT<T>::add2(m_lds32(d1), f));
                d0 += (uint32_t)b.pitchB; d1 += (uint32_t)b.pitchB;
                p0 = c0; p1 = c1; c0 = n0; c1 = n1;
            }
        }
    }
}

// The same with TWO source columns (a, a+1; a even) per lane: three aligned 32-bit loads of T feed four destination
// columns, and the loop / address overhead is paid once per eight outputs.  Same arithmetic per output as above
// (bit-identical).  RECNEXT_MDBG=4; measured 8 % SLOWER at stage 0 (0.157 vs 0.145 ms): the two row groups a warp then
// needs collide in the banks — kept as a checked alternative (tools/ab_upsample.py).
template <typename T>
__device__ __forceinline__ void m_up2x_add2(const MTeam<T>& tm, int l) {
    const MPlan& pl = tm.pl;
    const MLevel& ls = pl.lv[l];
    const MLevel& ld = pl.lv[l - 1];
    const int LW = 1 << ls.up2_shift;
    const int j = tm.tl & (LW - 1), rg = tm.tl >> ls.up2_shift;
    const int Hs = ls.H, Ws = ls.W, Wd = ld.W;
    const int rpg = ls.up2_rpg;
    const int m0 = rg * rpg, m1 = (m0 + rpg) < Hs ? (m0 + rpg) : Hs;
    if (m0 >= m1) return;
    const uint32_t tpB = (uint32_t)ls.tpB;
    for (int a = 2 * j; a < Ws; a += 2 * LW) {
        const bool second = 2 * a + 2 < Wd;   // destination columns 2a+2, 2a+3 exist (source column a+1 < Ws)
        for (int g = 0; g < pl.G; ++g) {
            const MBuf b = tm.buf(g, l - 1);
            const uint32_t Tc = tm.tbuf(g, l) + 2u * a;   // elements a, a+1 = source columns a-2, a-1 (interior at 2)
            auto hrow = [&](uint32_t p, float (&h)[4]) {
                const uint32_t w0 = m_lds32(p), w1 = m_lds32(p + 4), w2 = m_lds32(p + 8);
                const float tm1 = MmaT<T>::unpack(w0).y;
                const float2 t01 = MmaT<T>::unpack(w1);
                const float tp2 = MmaT<T>::lo(w2);
                const float qa = 0.75f * t01.x, qb = 0.75f * t01.y;
                h[0] = fmaf(0.25f, tm1, qa);
                h[1] = fmaf(0.25f, t01.y, qa);
                h[2] = fmaf(0.25f, t01.x, qb);
                h[3] = fmaf(0.25f, tp2, qb);
            };
            float hp[4], hc[4], hn[4];
            hrow(Tc + (uint32_t)(m0 > 0 ? m0 - 1 : 0) * tpB, hp);
            uint32_t tp = Tc + (uint32_t)m0 * tpB;
            hrow(tp, hc);
            uint32_t d0 = b.row(2 * m0 + 2) + 4u + 4u * a, d1 = b.row(2 * m0 + 3) + 4u + 4u * a;
            for (int m = m0; m < m1; ++m) {
                if (m + 1 < Hs) tp += tpB;
                hrow(tp, hn);
                float q[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) q[c] = 0.75f * hc[c];
                m_sts32(d0, MmaT<T>::add2(m_lds32(d0), MmaT<T>::pack(fmaf(0.25f, hp[0], q[0]), fmaf(0.25f, hp[1], q[1]))));
                m_sts32(d1, MmaT<T>::add2(m_lds32(d1), MmaT<T>::pack(fmaf(0.25f, hn[0], q[0]), fmaf(0.25f, hn[1], q[1]))));
                if (second) {
                    m_sts32(d0 + 4u, MmaT<T>::add2(m_lds32(d0 + 4u), MmaT<T>::pack(fmaf(0.25f, hp[2], q[2]), fmaf(0.25f, hp[3], q[3]))));
                    m_sts32(d1 + 4u, MmaT<T>::add2(m_lds32(d1 + 4u), MmaT<T>::pack(fmaf(0.25f, hn[2], q[2]), fmaf(0.25f, hn[3], q[3]))));
                }
                d0 += (uint32_t)b.pitchB; d1 += (uint32_t)b.pitchB;
#pragma unroll
                for (int c = 0; c < 4; ++c) { hp[c] = hc[c]; hc[c] = hn[c]; }
            }
        }
    }
}

// Exact-2x bilinear upsample + add on the tensor cores (RECNEXT_MDBG=8; measured 5 % slower than the CUDA-core version
// above because its fragment-shaped read-modify-write is two-way bank conflicted — kept as a checked alternative).  The horizontal pass is one MMA per (16 source rows, 8
// destination columns): P[m, J] = sum_n T[m, n] * Ux[J, n] with A = 16 rows x 16 elements of T (ldmatrix) and B = the
// constant 0.25 / 0.75 band (exact in bf16; products exact, fp32 sums of two terms: P is the exact fp32 result).
// Fragment row g is source row m0 - 1 + 2g and fragment row g + 8 is m0 + 2g, so a thread holds two vertically
// adjacent source rows and gets the other two neighbours from lanes g -+ 1 with one shuffle each; the vertical
// 0.25 / 0.75 combination, the rounding and the `x + u` read-modify-write stay in fp32 / packed bf16 on the CUDA
// cores.  An M-tile yields the 14 source rows m0 .. m0 + 13 (28 destination rows); rows and the replicate border
// columns kept in T implement ATen's index clamping.  Same operation order as m_up2x_add: bit-identical results.
template <typename T>
__device__ __forceinline__ void m_up2x_mma(const MTeam<T>& tm, int l) {
    const MPlan& pl = tm.pl;
    const MLevel& ls = pl.lv[l];
    const MLevel& ld = pl.lv[l - 1];
    const int lane = tm.lane, g8 = lane >> 2, t4 = lane & 3;
    const int Hs = ls.H, Wd = ld.W;
    const int NTd = ld.NT;                     // destination n-tiles (8 columns)
    const int MTs = (Hs + 13) / 14;            // source M-tiles of 14 rows
    // constant B fragments: destination column j of an n-tile (0..7), window element k (0..15); q even / odd
    uint32_t Bc[2][2];
#pragma unroll
    for (int par = 0; par < 2; ++par)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 2 * t4 + 8 * h + e - 4 * par, j = g8, a = j >> 1;
                v[e] = (j & 1) ? (k == a + 2 ? 0.75f : (k == a + 3 ? 0.25f : 0.f)) : (k == a + 1 ? 0.25f : (k == a + 2 ? 0.75f : 0.f));
            }
            Bc[par][h] = MmaT<T>::pack(v[0], v[1]);
        }
    const int frow = 2 * (lane & 7) + ((lane >> 3) & 1) - 1;   // source row offset of this lane's ldmatrix address
    for (int g = 0, mt = tm.wt; g < pl.G; mt += pl.TW) {
        if (mt >= MTs) { mt -= MTs + pl.TW; ++g; continue; }
        const MBuf b = tm.buf(g, l - 1);
        const int m0 = 14 * mt;
        int rr = m0 + frow;
        rr = rr < 0 ? 0 : (rr > Hs - 1 ? Hs - 1 : rr);
        const uint32_t arow = tm.tbuf(g, l) + (uint32_t)(rr * ls.tpB) + (uint32_t)(lane >> 4) * 16u;
        const int r0 = m0 - 1 + 2 * g8, r1 = r0 + 1;             // the two source rows of this thread
        const bool v0 = g8 >= 1 && r0 < Hs, v1 = g8 <= 6 && r1 < Hs;
        const uint32_t d00 = b.row(2 * r0 + 2) + 4u + 4u * t4, d01 = b.row(2 * r0 + 3) + 4u + 4u * t4;
        const uint32_t d10 = b.row(2 * r1 + 2) + 4u + 4u * t4, d11 = b.row(2 * r1 + 3) + 4u + 4u * t4;
        for (int p = 0; 2 * p < NTd; ++p) {
            uint32_t a0, a1, a2, a3;
            m_ldsm4(a0, a1, a2, a3, arow + (uint32_t)p * 16u);
#pragma unroll
            for (int par = 0; par < 2; ++par) {
                const int q = 2 * p + par;
                if (q < NTd) {
                    float P[4] = {0.f, 0.f, 0.f, 0.f};
                    MmaT<T>::mma16(P, a0, a1, a2, a3, Bc[par][0], Bc[par][1]);
                    const float up0 = __shfl_up_sync(0xffffffffu, P[2], 4), up1 = __shfl_up_sync(0xffffffffu, P[3], 4);
                    const float dn0 = __shfl_down_sync(0xffffffffu, P[0], 4), dn1 = __shfl_down_sync(0xffffffffu, P[1], 4);
                    if (8 * q + 2 * t4 < Wd) {
                        const uint32_t co = 16u * q;
                        const float q00 = 0.75f * P[0], q01 = 0.75f * P[1], q10 = 0.75f * P[2], q11 = 0.75f * P[3];
                        if (v0) {
                            const uint32_t e = MmaT<T>::pack(fmaf(0.25f, up0, q00), fmaf(0.25f, up1, q01));
                            const uint32_t f = MmaT<T>::pack(fmaf(0.25f, P[2], q00), fmaf(0.25f, P[3], q01));
                            m_sts32(d00 + co, MmaT<T>::add2(m_lds32(d00 + co), e));
                            m_sts32(d01 + co, MmaT<T>::add2(m_lds32(d01 + co), f));
                        }
                        if (v1) {
                            const uint32_t e = MmaT<T>::pack(fmaf(0.25f, P[0], q10), fmaf(0.25f, P[1], q11));
                            const uint32_t f = MmaT<T>::pack(fmaf(0.25f, dn0, q10), fmaf(0.25f, dn1, q11));
                            m_sts32(d10 + co, MmaT<T>::add2(m_lds32(d10 + co), e));
                            m_sts32(d11 + co, MmaT<T>::add2(m_lds32(d11 + co), f));
                        }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// the kernel body.  pl = layout, rt = run-time fields (the same object in the generic kernel).  LS >= 0: the level
// count is a compile-time constant and every loop over levels is fully unrolled, so that with a constexpr `pl`
// all sizes, pitches and offsets become immediates.
// ---------------------------------------------------------------------------------------------------------
template <typename T, int LS, int VS = -1>
__device__ __forceinline__ void m_fwd_body(const MPlan& pl, const MPlan& rt, const KernelArgs& a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int team = warp / pl.TW, wt = warp - team * pl.TW;
    const uint32_t smem32 = rc_smem_u32(smem);
    MTeam<T> tm{pl, rt, smem32, smem32 + (uint32_t)(pl.smTeams + team * pl.team_bytes), team, wt, lane, wt * 32 + lane};
    const int L = LS >= 0 ? LS : pl.L, G = pl.G;
    const int variant = VS >= 0 ? VS : (LS >= 0 ? 0 : rt.variant);   // compile-time in the specialised kernels (RecConv2d: 0; RecAttn2d pieces: 1 / 2)
    const int lg = lane >> 2, lt = lane & 3;

    // ---- CTA init: zero the team slices (the borders of the padded buffers stay zero), interpolation tables, mbarriers
    {
        uint4* z = reinterpret_cast<uint4*>(smem + pl.smTeams);
        const int n16 = pl.NTEAM * pl.team_bytes / 16;
        const uint4 zero = {0u, 0u, 0u, 0u};
        for (int i = tid; i < n16; i += blockDim.x) z[i] = zero;
#pragma unroll
        for (int l = 1; l <= L; ++l) {
            if (pl.lv[l].tabY < 0) continue;
            rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + pl.smTab + pl.lv[l].tabY), pl.lv[l].H, pl.lv[l - 1].H, pl.mode, tid, blockDim.x);
            rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + pl.smTab + pl.lv[l].tabX), pl.lv[l].W, pl.lv[l - 1].W, pl.mode, tid, blockDim.x);
        }
    }
    const uint32_t bar = smem32 + (uint32_t)(pl.smBar + 8 * team);
    uint32_t phase = 0;
    if (rt.use_tma && tm.tl == 0) rc_mbar_init(bar, 1);
    rc_fence_proxy_async();  // the zero fill above also touched the bulk-copy buffers
    __syncthreads();

    const int plane_elems = pl.H * pl.W;
    const T* gx = reinterpret_cast<const T*>(a.x);
    T* gy = reinterpret_cast<T*>(a.out);
    const long total = (long)rt.n_cg * rt.B;
    const long start = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1) / gridDim.x;
    if (start >= end) return;
    long my = start + team;
    unsigned char* raw = smem + pl.smTeams + team * pl.team_bytes + pl.off_upper;   // aliases levels >= 1 and T
    auto plane0_of = [&](long idx) { const int cg = (int)(idx / rt.B), n = (int)(idx - (long)cg * rt.B); return ((long)n * rt.C + cg * G) * (long)plane_elems; };
    auto issue_load = [&](long idx) {
        if (tm.tl == 0) {
            rc_mbar_expect_tx(bar, (uint32_t)pl.raw_bytes);
            rc_bulk_g2s(raw, gx + plane0_of(idx), (uint32_t)pl.raw_bytes, bar);
        }
    };
    if (rt.use_tma && my < end) issue_load(my);
    const int cg_first = (int)(start / rt.B), cg_last = (int)((end - 1) / rt.B);
    for (int cg = cg_first; cg <= cg_last; ++cg) {
        __syncthreads();  // every team is done with the previous channel group's fragments
        m_build_frags<T>(pl, rt, a, smem, cg, tid, blockDim.x);
        __syncthreads();
        const long cg_end = ((long)(cg + 1) * rt.B) < end ? ((long)(cg + 1) * rt.B) : end;
        for (; my < cg_end; my += pl.NTEAM) {
            const long p0 = plane0_of(my);
            const long pidx = p0 / plane_elems;     // index of the batch's first (n, c) plane
            const long nxt = my + pl.NTEAM;
            // ---- x -> padded level 0
            if (rt.use_tma) {
                while (!rc_mbar_try_wait(bar, phase)) {}
                phase ^= 1u;
                m_repack<T>(tm, reinterpret_cast<const T*>(raw));
                tm.sync();
                if (L == 0 || variant == 1) {
                    rc_fence_proxy_async();  // generic reads of `raw` before the next bulk copy's writes
                    if (nxt < end) issue_load(nxt);
                } else if (variant == 2) {
                    // level 1 is not materialised; T is filled below
                }
            } else {
                m_repack<T>(tm, gx + p0);
                tm.sync();
            }
            if (variant == 0 && L > 0) {
                // the raw batch and the T buffers of the previous batch landed on top of levels >= 1: restore their zero borders
                const uint4 zero = {0u, 0u, 0u, 0u};
                for (int g = 0; g < G; ++g) {
                    uint4* z = reinterpret_cast<uint4*>(raw + g * pl.upper_bytes);
                    for (int i = tm.tl; i < pl.upper_bytes / 16; i += pl.team_lanes) z[i] = zero;
                }
                tm.sync();
            }
            // ---- down chain: x_l = down(x_{l-1})   (model/recnext.py:27-29)
            if (variant == 1) {  // RecAttn2d.down[0] (model/recattn.py:60): x_1 = down(x) straight to global memory
                const MLevel& lo = pl.lv[1];
                for (int g = 0, mt = wt; g < G; mt += pl.TW) {
                    if (mt >= lo.MT) { mt -= lo.MT + pl.TW; ++g; continue; }
                    const MBuf in = tm.buf(g, 0);
                    const int i0 = mt * 16, Ho = lo.H, Wo = lo.W;
                    const int ia = i0 + lg, ib = ia + 8;
                    T* dst = gy + (pidx + g) * (long)(Ho * Wo) + 2 * lt;
                    const bool even = (Wo & 1) == 0;
                    m_conv_rows<T, true>(in, lo.ntc, i0, lo.NT, tm.frag(g, 0), tm.bias(g, 0), lane, [&](int q, const float (&acc)[4]) {
                        const int c = 8 * q + 2 * lt;
                        if (c < Wo) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int i = h ? ib : ia;
                                if (i < Ho) {
                                    const uint32_t w = MmaT<T>::pack(acc[2 * h], acc[2 * h + 1]);
                                    T* d = dst + i * Wo + 8 * q;
                                    if (even) *reinterpret_cast<uint32_t*>(d) = w;
                                    else {
                                        *reinterpret_cast<unsigned short*>(d) = (unsigned short)(w & 0xffffu);
                                        if (c + 1 < Wo) *reinterpret_cast<unsigned short*>(d + 1) = (unsigned short)(w >> 16);
                                    }
                                }
                            }
                        }
                    });
                }
                tm.sync();  // level 0 is rewritten by the next batch's repack
                continue;
            }
            if (variant == 2) {  // RecAttn2d tail (model/recattn.py:67): s_0 = x + interpolate(z); y = conv(s_0)
                m_load_z<T>(tm, reinterpret_cast<const T*>(a.gy) + pidx * (long)(pl.lv[1].H * pl.lv[1].W));
                tm.sync();
                const MLevel& lv = pl.lv[1];
                if (lv.tabY < 0) { if (rt.dbg & 4) m_up2x_add2<T>(tm, 1); else if (rt.dbg & 8) m_up2x_mma<T>(tm, 1); else m_up2x_add<T>(tm, 1); }
                else m_up_add<T>(tm, smem, 1);
                tm.sync();
            }
            const int Lp = variant == 2 ? 0 : L;  // levels of the pyramid proper
#pragma unroll
            for (int l = 1; l <= Lp; ++l) {
                const MLevel& lo = pl.lv[l];
                for (int g = 0, mt = wt; g < G; mt += pl.TW) {
                    if (mt >= lo.MT) { mt -= lo.MT + pl.TW; ++g; continue; }
                    const MBuf in = tm.buf(g, l - 1);
                    const MBuf out = tm.buf(g, l);
                    const int i0 = mt * 16, Wo = lo.W;
                    const int ia = i0 + lg, ib = ia + 8;
                    const bool va = ia < lo.H, vb = ib < lo.H;
                    const uint32_t da = out.row(ia + 2) + 4u + 4u * lt, db = out.row(ib + 2) + 4u + 4u * lt;
                    m_conv_rows<T, true>(in, lo.ntc, i0, lo.NT, tm.frag(g, 0), tm.bias(g, 0), lane, [&](int q, const float (&acc)[4]) {
                        const int c = 8 * q + 2 * lt;
                        if (c < Wo) {
                            const bool pair = c + 1 < Wo;
                            if (va) m_sts32(da + 16u * q, MmaT<T>::pack(acc[0], pair ? acc[1] : 0.f));
                            if (vb) m_sts32(db + 16u * q, MmaT<T>::pack(acc[2], pair ? acc[3] : 0.f));
                        }
                    });
                }
                tm.sync();
            }
            // ---- up pass: t_l = convs[L-l](s_l); s_{l-1} = x_{l-1} + interpolate(t_l)   (model/recnext.py:31-33)
#pragma unroll
            for (int l = Lp; l >= 1; --l) {
                const MLevel& lv = pl.lv[l];
                for (int g = 0, mt = wt; g < G; mt += pl.TW) {
                    if (mt >= lv.MT) { mt -= lv.MT + pl.TW; ++g; continue; }
                    const MBuf in = tm.buf(g, l);
                    const int i0 = mt * 16, Wl = lv.W;
                    const int ia = i0 + 2 * lg;
                    const bool va = ia < lv.H, vb = ia + 1 < lv.H;
                    const uint32_t ta = tm.tbuf(g, l) + (uint32_t)(ia * lv.tpB) + 4u + 4u * lt, tpB = (uint32_t)lv.tpB;
                    m_conv_rows<T, false>(in, lv.ntc, i0, lv.NT, tm.frag(g, 20 + 10 * (L - l)), tm.bias(g, 1 + (L - l)), lane,
                                          [&](int q, const float (&acc)[4]) {
                        const int c = 8 * q + 2 * lt;
                        if (c < Wl) {
                            const bool pair = c + 1 < Wl;  // else column c is the last one: its pair slot is the replicate border
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (h ? vb : va) {
                                    const uint32_t w = MmaT<T>::pack(acc[2 * h], pair ? acc[2 * h + 1] : acc[2 * h]);
                                    const uint32_t ad = ta + h * tpB + 16u * q;
                                    m_sts32(ad, w);
                                    if (c == 0) m_sts16(ad - 2u, w);                       // left replicate border
                                    if (c + 1 == Wl - 1) m_sts16(ad + 4u, w >> 16);        // right replicate border
                                }
                            }
                        }
                    });
                }
                tm.sync();
                if (lv.tabY < 0) { if (rt.dbg & 4) m_up2x_add2<T>(tm, l); else if (rt.dbg & 8) m_up2x_mma<T>(tm, l); else m_up2x_add<T>(tm, l); }
                else m_up_add<T>(tm, smem, l);
                tm.sync();
            }
            if (rt.use_tma && L > 0) {  // levels >= 1 and T are dead: fetch the next batch under the final conv
                rc_fence_proxy_async();
                tm.sync();
                if (nxt < end) issue_load(nxt);
            }
            // ---- y = convs[L](s_0) -> global   (model/recnext.py:34)
            {
                const MLevel& lv = pl.lv[0];
                const int H = pl.H, W = pl.W;
                for (int g = 0, mt = wt; g < G; mt += pl.TW) {
                    if (mt >= lv.MT) { mt -= lv.MT + pl.TW; ++g; continue; }
                    const MBuf in = tm.buf(g, 0);
                    const int i0 = mt * 16;
                    const int ia = i0 + 2 * lg;
                    const bool va = ia < H, vb = ia + 1 < H;
                    T* da = gy + p0 + (long)g * plane_elems + ia * W + 2 * lt;
                    const bool even = (W & 1) == 0;
                    m_conv_rows<T, false>(in, lv.ntc, i0, lv.NT, tm.frag(g, 20 + 10 * L), tm.bias(g, 1 + L), lane,
                                          [&](int q, const float (&acc)[4]) {
                        const int c = 8 * q + 2 * lt;
                        if (c < W) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (h ? vb : va) {
                                    const uint32_t w = MmaT<T>::pack(acc[2 * h], acc[2 * h + 1]);
                                    T* d = da + h * W + 8 * q;
                                    if (even) *reinterpret_cast<uint32_t*>(d) = w;
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

// generic kernel: everything from the run-time plan.  Up to 16 warps at 128 registers.
template <typename T>
__global__ void __launch_bounds__(512, 1) recconv_mfwd_kernel(const __grid_constant__ MPlan pl, const __grid_constant__ KernelArgs a) {
    m_fwd_body<T, -1>(pl, pl, a);
}
// specialised kernel: plane geometry (H0 x W0, L0 levels, G0 planes per batch), variant (0 RecConv2d, 1 / 2 the RecAttn2d pieces) and
// interpolation mode fixed at compile time
template <typename T, int H0, int W0, int L0, int G0, int V0 = 0, int MODE0 = 0>
__global__ void __launch_bounds__(32 * m_static_max_warps(H0, W0), 1) recconv_mfwd_static_kernel(const __grid_constant__ MPlan rt, const __grid_constant__ KernelArgs a) {
    constexpr MPlan sp = m_static_plan(H0, W0, L0, G0, std::is_same<T, __half>::value ? 2 : 1, V0, MODE0);
    m_fwd_body<T, L0, V0>(sp, rt, a);
}

template <typename T, int H0, int W0, int L0, int G0, int V0 = 0, int MODE0 = 0>
inline bool m_try_static(const MPlan& pl, const KernelArgs& a, cudaStream_t stream, cudaError_t& err) {
    constexpr MPlan sp = m_static_plan(H0, W0, L0, G0, std::is_same<T, __half>::value ? 2 : 1, V0, MODE0);
    if (pl.variant != V0 || pl.mode != MODE0 || pl.H != H0 || pl.W != W0 || pl.L != L0 || pl.G != G0) return false;
    const MPlan st = m_static_patched(sp, pl);
    if (memcmp(&st, &pl, sizeof(MPlan)) != 0) return false;
    static DeviceOnce configured = {};
    err = rc_once_per_device(configured, [] {
        return cudaFuncSetAttribute(recconv_mfwd_static_kernel<T, H0, W0, L0, G0, V0, MODE0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (err != cudaSuccess) return true;
    recconv_mfwd_static_kernel<T, H0, W0, L0, G0, V0, MODE0><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
    err = cudaGetLastError();
    return true;
}

// true if the run-time plan is one of the compile-time geometries (introspection: recconv_plan_describe)
template <int H0, int W0, int L0, int G0>
inline bool m_matches_static(const MPlan& pl) {
    constexpr MPlan sp = m_static_plan(H0, W0, L0, G0, 1);
    if (pl.variant != 0 || pl.dtype != 1 || pl.H != H0 || pl.W != W0 || pl.L != L0 || pl.G != G0) return false;
    const MPlan st = m_static_patched(sp, pl);
    return memcmp(&st, &pl, sizeof(MPlan)) == 0;
}
inline bool m_is_static(const MPlan& pl) {
    return !(pl.dbg & 2) && (m_matches_static<56, 56, 4, 1>(pl) || m_matches_static<28, 28, 3, 1>(pl) || m_matches_static<14, 14, 2, 4>(pl) ||
                             m_matches_static<7, 7, 1, 8>(pl));
}

template <typename T>
inline cudaError_t m_launch_fwd_t(const MPlan& pl, const KernelArgs& a, cudaStream_t stream, bool allow_static) {
    cudaError_t err = cudaSuccess;
    // the RecNeXt stage shapes at 224 px (model/recnext.py:152: level = 4 - stage)
    if constexpr (std::is_same<T, __nv_bfloat16>::value) if (allow_static) {
        if (m_try_static<T, 56, 56, 4, 1>(pl, a, stream, err)) return err;
        if (m_try_static<T, 28, 28, 3, 1>(pl, a, stream, err)) return err;
        if (m_try_static<T, 14, 14, 2, 4>(pl, a, stream, err)) return err;
        if (m_try_static<T, 7, 7, 1, 8>(pl, a, stream, err)) return err;
        // the RecAttn2d pieces of the A-series at the same stage shapes (model/recattn.py:54-67: one level, mode 'nearest')
        if (pl.variant == 1) {
            if (m_try_static<T, 56, 56, 1, 1, 1, 1>(pl, a, stream, err)) return err;
            if (m_try_static<T, 28, 28, 1, 1, 1, 1>(pl, a, stream, err)) return err;
            if (m_try_static<T, 14, 14, 1, 4, 1, 1>(pl, a, stream, err)) return err;
            if (m_try_static<T, 7, 7, 1, 8, 1, 1>(pl, a, stream, err)) return err;
        } else if (pl.variant == 2) {
            if (m_try_static<T, 56, 56, 1, 1, 2, 1>(pl, a, stream, err)) return err;
            if (m_try_static<T, 28, 28, 1, 1, 2, 1>(pl, a, stream, err)) return err;
            if (m_try_static<T, 14, 14, 1, 4, 2, 1>(pl, a, stream, err)) return err;
            if (m_try_static<T, 7, 7, 1, 8, 2, 1>(pl, a, stream, err)) return err;
        }
    }
    static DeviceOnce configured = {};
    err = rc_once_per_device(configured, [] { return cudaFuncSetAttribute(recconv_mfwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    if (err != cudaSuccess) return err;
    if (pl.threads > 512) return cudaErrorInvalidConfiguration;  // planned for a specialised kernel that did not match
    recconv_mfwd_kernel<T><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
    return cudaGetLastError();
}

inline cudaError_t m_launch_fwd(const MPlan& pl, const KernelArgs& a, cudaStream_t stream) {
    const bool allow_static = !(pl.dbg & 2);
    if (pl.dtype == 1) return m_launch_fwd_t<__nv_bfloat16>(pl, a, stream, allow_static);
    if (pl.dtype == 2) return m_launch_fwd_t<__half>(pl, a, stream, allow_static);
    return cudaErrorInvalidValue;
}

}  // namespace recnext
