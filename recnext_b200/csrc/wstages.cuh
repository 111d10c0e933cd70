// wstages.cuh — per-lane stage work of the team-resident RecConv kernels (see wplan.h).
//
// Every function is a parallel-for over the ITEMS of one stage: lane `lane` of `nl` takes items lane, lane + nl, ...
// An item is (plane g of the batch, 4-column strip, block of rows); stages never communicate inside a stage, so
// the same code runs on the CPU for the schedule tests (tests/emu) — the CUDA build never uses the host versions.
// Reference semantics: model/recnext.py:21-34 (see recconv_stages.cuh for the ATen index contract).
#pragma once
#include "recconv_stages.cuh"
#include "wplan.h"

namespace recnext {

struct WItem { int r0, c0, nrows; bool valid; };

// item `jl + round * LPP` of this lane's plane (see WGrid)
RC_HD WItem w_item(const WGrid& gr, int jl, int LPP, int round, int rows_total) {
    WItem it;
    const int j = jl + round * LPP;
    const int rb = rc_fastdiv(j, gr.m_spr), st = j - rb * gr.spr;
    it.valid = j < gr.ipp && st < gr.strips;
    it.c0 = st * kStripW;
    it.r0 = rb * gr.rpb;
    const int left = rows_total - it.r0;
    it.nrows = left < gr.rpb ? left : gr.rpb;
    return it;
}

// ---------------------------------------------------------------------------------------------------------
// depthwise KxK stride-1 cross-correlation over a padded level buffer of THIS LANE'S plane; FLIP = transpose
// (input gradient).  wslot: the plane's filter slot in shared memory (bias at [K*K]).  epi(row, c0, acc[4]).
// ---------------------------------------------------------------------------------------------------------
template <int K, bool FLIP, class Epi>
RC_HD void w_conv_s1(const float* __restrict__ src, int pitch, int Ho, const WGrid& gr, const float* __restrict__ wslot,
                     bool use_bias, int jl, int LPP, Epi epi) {
    constexpr int WL = (kStripW + 2 * (K / 2) + 3) & ~3;
    if (jl >= gr.ipp) return;
    float w[K * K];
    rc_load_filter<K * K>(w, wslot, FLIP);
    const float bias = use_bias ? wslot[K * K] : 0.f;
    for (int rd = 0; rd < gr.rounds; ++rd) {
        const WItem it = w_item(gr, jl, LPP, rd, Ho);
        if (!it.valid) continue;
        const float* base = src + it.r0 * pitch + it.c0;
        float win[K][WL];
#pragma unroll
        for (int r = 0; r < K - 1; ++r) rc_load_row<WL>(win[r], base + r * pitch);
        for (int o = 0; o < it.nrows; o += K) {
#pragma unroll
            for (int ph = 0; ph < K; ++ph) {
                if (o + ph < it.nrows) {
                    rc_load_row<WL>(win[(ph + K - 1) % K], base + (o + ph + K - 1) * pitch);
                    float acc[kStripW] = {bias, bias, bias, bias};
#pragma unroll
                    for (int r = 0; r < K; ++r)
#pragma unroll
                        for (int s = 0; s < K; ++s)
#pragma unroll
                            for (int c = 0; c < kStripW; ++c) acc[c] = fmaf(w[r * K + s], win[(ph + r) % K][c + s], acc[c]);
                    epi(it.r0 + o + ph, it.c0, acc);
                }
            }
        }
    }
}

// depthwise KxK STRIDE-2 cross-correlation (the shared `down` filter): level l-1 (padded) -> rows of level l
template <int K, class Epi>
RC_HD void w_conv_s2(const float* __restrict__ src, int pitch, int Ho, const WGrid& gr, const float* __restrict__ wslot,
                     bool use_bias, int jl, int LPP, Epi epi) {
    constexpr int WL = (2 * (kStripW - 1) + K + 3) & ~3;
    if (jl >= gr.ipp) return;
    float w[K * K];
    rc_load_filter<K * K>(w, wslot, false);
    const float bias = use_bias ? wslot[K * K] : 0.f;
    for (int rd = 0; rd < gr.rounds; ++rd) {
        const WItem it = w_item(gr, jl, LPP, rd, Ho);
        if (!it.valid) continue;
        const float* base = src + 2 * it.r0 * pitch + 2 * it.c0;
        float win[K][WL];
#pragma unroll
        for (int r = 0; r < K - 2; ++r) rc_load_row<WL>(win[r], base + r * pitch);
        for (int o = 0; o < it.nrows; o += K) {
#pragma unroll
            for (int ph = 0; ph < K; ++ph) {
                if (o + ph < it.nrows) {
                    const float* rowp = base + (2 * (o + ph) + K - 2) * pitch;
                    rc_load_row<WL>(win[(2 * ph + K - 2) % K], rowp);
                    rc_load_row<WL>(win[(2 * ph + K - 1) % K], rowp + pitch);
                    float acc[kStripW] = {bias, bias, bias, bias};
#pragma unroll
                    for (int r = 0; r < K; ++r)
#pragma unroll
                        for (int s = 0; s < K; ++s)
#pragma unroll
                            for (int c = 0; c < kStripW; ++c)
                                acc[c] = fmaf(w[r * K + s], win[(2 * ph + r) % K][2 * c + s], acc[c]);
                    epi(it.r0 + o + ph, it.c0, acc);
                }
            }
        }
    }
}

// raw plane (element type T, unpadded, W even) -> padded fp32 interior.  An item is a column PAIR x a block of rows:
// consecutive lanes read consecutive 4/8-byte words of a raw row and write consecutive 8-byte words.
template <typename T>
RC_HD void w_unpack_pairs(const T* __restrict__ raw, int H, int W, float* __restrict__ dst_interior, int pitch, const WGrid& gr,
                          int jl, int LPP) {
    for (int rd = 0; rd < gr.rounds; ++rd) {
        const int j = jl + rd * LPP;
        const int rb = rc_fastdiv(j, gr.m_spr), cp = j - rb * gr.spr;
        if (j >= gr.ipp || cp >= gr.strips) continue;
        const int r0 = rb * gr.rpb;
        const int r1 = (r0 + gr.rpb) < H ? (r0 + gr.rpb) : H;
        const T* s = raw + (long)r0 * W + 2 * cp;
        float* d = dst_interior + r0 * pitch + 2 * cp;
        for (int r = r0; r < r1; ++r) {
            alignas(8) T v[2];
            if (sizeof(T) == 4) *reinterpret_cast<float2*>(v) = *reinterpret_cast<const float2*>(s);
            else *reinterpret_cast<float*>(v) = *reinterpret_cast<const float*>(s);
            *reinterpret_cast<float2*>(d) = make_float2(Elem<T>::to_f(v[0]), Elem<T>::to_f(v[1]));
            s += W; d += pitch;
        }
    }
}

// masked store of 4 results into a padded level buffer (interior coordinates row, c0); pads stay zero
RC_HD void w_store_level(float* __restrict__ dst_interior, int pitch, int W, int row, int c0, const float (&v)[kStripW]) {
    float* d = dst_interior + row * pitch + c0;
    if (c0 + kStripW <= W) {
        if ((reinterpret_cast<uintptr_t>(d) & 7) == 0) {  // even pad and pitch: two aligned pairs
            reinterpret_cast<float2*>(d)[0] = make_float2(v[0], v[1]);
            reinterpret_cast<float2*>(d)[1] = make_float2(v[2], v[3]);
        } else {
#pragma unroll
            for (int c = 0; c < kStripW; ++c) d[c] = v[c];
        }
    } else {
#pragma unroll
        for (int c = 0; c < kStripW; ++c)
            if (c0 + c < W) d[c] = v[c];
    }
}

// ---------------------------------------------------------------------------------------------------------
// s_{l-1} += interpolate(t_l, size of level l-1)   (model/recnext.py:33 with the next `f + x` folded in).
// T: unpadded rows of tp floats.  Exact-2x bilinear: the align_corners=False weights are the fixed 0.75 / 0.25
// stencil with source indices clamped at the border (ATen gives i0 = a-1 (lambda .75) for even, a (lambda .25)
// for odd destinations; the clamped ends reduce to the border pixel).
// ---------------------------------------------------------------------------------------------------------
RC_HD void w_hrow2x(float (&h)[kStripW], const float* __restrict__ trow, int ca, int cb, int cc, int cd) {
    const float a = trow[ca], b = trow[cb], c = trow[cc], d = trow[cd];
    h[0] = fmaf(0.75f, b, 0.25f * a);
    h[1] = fmaf(0.75f, b, 0.25f * c);
    h[2] = fmaf(0.75f, c, 0.25f * b);
    h[3] = fmaf(0.75f, c, 0.25f * d);
}

// dstS: INTERIOR origin of this lane's level l-1 buffer; T: this lane's plane of the T buffer
RC_HD void w_up2x_add(float* __restrict__ dstS, int pitch, int Wd, const float* __restrict__ T, int tp, int Hl, int Wl,
                      const WGrid& gr, int jl, int LPP) {
    for (int rd = 0; rd < gr.rounds; ++rd) {
        const WItem it = w_item(gr, jl, LPP, rd, Hl);  // rows = SOURCE rows, strips = destination strips
        if (!it.valid) continue;
        const int j0 = it.c0, cb = j0 >> 1, ca = cb > 0 ? cb - 1 : 0;
        const int cc = cb + 1 < Wl ? cb + 1 : Wl - 1, cd = cb + 2 < Wl ? cb + 2 : Wl - 1;
        const int m0 = it.r0, m1 = it.r0 + it.nrows;
        float hp[kStripW], hc[kStripW], hn[kStripW];
        w_hrow2x(hp, T + (m0 > 0 ? m0 - 1 : 0) * tp, ca, cb, cc, cd);
        w_hrow2x(hc, T + m0 * tp, ca, cb, cc, cd);
        float* d = dstS + 2 * m0 * pitch + j0;
        const bool vec = j0 + kStripW <= Wd && ((reinterpret_cast<uintptr_t>(d) | (uintptr_t)(pitch * 4)) & 7) == 0;
        for (int m = m0; m < m1; ++m) {
            w_hrow2x(hn, T + (m + 1 < Hl ? m + 1 : Hl - 1) * tp, ca, cb, cc, cd);
            if (vec) {
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const float (&ho)[kStripW] = rr ? hn : hp;
                    float2* d2 = reinterpret_cast<float2*>(d + rr * pitch);
                    float2 u = d2[0], v = d2[1];
                    u.x += fmaf(0.75f, hc[0], 0.25f * ho[0]); u.y += fmaf(0.75f, hc[1], 0.25f * ho[1]);
                    v.x += fmaf(0.75f, hc[2], 0.25f * ho[2]); v.y += fmaf(0.75f, hc[3], 0.25f * ho[3]);
                    d2[0] = u; d2[1] = v;
                }
            } else {
#pragma unroll
                for (int c = 0; c < kStripW; ++c)
                    if (j0 + c < Wd) {
                        d[c] += fmaf(0.75f, hc[c], 0.25f * hp[c]);
                        d[pitch + c] += fmaf(0.75f, hc[c], 0.25f * hn[c]);
                    }
            }
            d += 2 * pitch;
#pragma unroll
            for (int c = 0; c < kStripW; ++c) { hp[c] = hc[c]; hc[c] = hn[c]; }
        }
    }
}

// Aligned exact-2x bilinear path (even pad, K = 5): an item is 4 PADDED columns [4q, 4q+3] of level l-1 (16-byte
// aligned, so the read-modify-write is one LDS.128 + STS.128 per row and consecutive lanes touch consecutive
// banks) x a block of source rows.  The 4 columns are interior columns j0 = 4q - pad .. j0 + 3 and read source
// columns a0 - 1 .. a0 + 2, a0 = j0 / 2: two aligned pairs of the T row, whose interior starts at column 2 behind
// a replicate border of 2 (written by the conv epilogue, w_store_T_fast).  Cells outside the interior get +0.
RC_HD void w_hrow2x_fast(float (&h)[kStripW], const float* __restrict__ trow) {
    const float2 ab = *reinterpret_cast<const float2*>(trow);      // source columns a0 - 1, a0
    const float2 cd = *reinterpret_cast<const float2*>(trow + 2);  // a0 + 1, a0 + 2
    h[0] = fmaf(0.75f, ab.y, 0.25f * ab.x);
    h[1] = fmaf(0.75f, ab.y, 0.25f * cd.x);
    h[2] = fmaf(0.75f, cd.x, 0.25f * ab.y);
    h[3] = fmaf(0.75f, cd.x, 0.25f * cd.y);
}
// dstS_padded: origin of the PADDED level l-1 buffer row `pad` (i.e. buffer + pad * pitch); T: plane's T buffer
RC_HD void w_up2x_add_fast(float* __restrict__ dstS_rows, int pitch, int pad, int Wd, const float* __restrict__ T, int tp, int Hl,
                           const WGrid& gr, int jl, int LPP) {
    for (int rd = 0; rd < gr.rounds; ++rd) {
        const WItem it = w_item(gr, jl, LPP, rd, Hl);  // rows = SOURCE rows, strips = groups of 4 padded columns
        if (!it.valid) continue;
        const int pc0 = it.c0;            // first padded column (multiple of 4)
        const int j0 = pc0 - pad;         // first interior column (even)
        const float* tcol = T + (j0 >> 1) + 1;  // T column of source a0 - 1 (interior offset 2): (j0/2 - 1) + 2
        float msk[kStripW];
#pragma unroll
        for (int c = 0; c < kStripW; ++c) msk[c] = (j0 + c >= 0 && j0 + c < Wd) ? 1.f : 0.f;
        const int m0 = it.r0, m1 = it.r0 + it.nrows;
        float hp[kStripW], hc[kStripW], hn[kStripW];
        w_hrow2x_fast(hp, tcol + (m0 > 0 ? m0 - 1 : 0) * tp);
        w_hrow2x_fast(hc, tcol + m0 * tp);
#pragma unroll
        for (int c = 0; c < kStripW; ++c) { hp[c] *= msk[c]; hc[c] *= msk[c]; }
        float* d = dstS_rows + 2 * m0 * pitch + pc0;
        for (int m = m0; m < m1; ++m) {
            w_hrow2x_fast(hn, tcol + (m + 1 < Hl ? m + 1 : Hl - 1) * tp);
#pragma unroll
            for (int c = 0; c < kStripW; ++c) hn[c] *= msk[c];
            float4 u = *reinterpret_cast<float4*>(d);
            float4 v = *reinterpret_cast<float4*>(d + pitch);
            u.x += fmaf(0.75f, hc[0], 0.25f * hp[0]); u.y += fmaf(0.75f, hc[1], 0.25f * hp[1]);
            u.z += fmaf(0.75f, hc[2], 0.25f * hp[2]); u.w += fmaf(0.75f, hc[3], 0.25f * hp[3]);
            v.x += fmaf(0.75f, hc[0], 0.25f * hn[0]); v.y += fmaf(0.75f, hc[1], 0.25f * hn[1]);
            v.z += fmaf(0.75f, hc[2], 0.25f * hn[2]); v.w += fmaf(0.75f, hc[3], 0.25f * hn[3]);
            *reinterpret_cast<float4*>(d) = u;
            *reinterpret_cast<float4*>(d + pitch) = v;
            d += 2 * pitch;
#pragma unroll
            for (int c = 0; c < kStripW; ++c) { hp[c] = hc[c]; hc[c] = hn[c]; }
        }
    }
}
// conv epilogue of the fast path: 4 results of row `row` into T (interior at column 2) + the replicate border
RC_HD void w_store_T_fast(float* __restrict__ T, int tp, int Wl, int row, int c0, const float (&v)[kStripW]) {
    float* r = T + row * tp + 2 + c0;
    reinterpret_cast<float2*>(r)[0] = make_float2(v[0], v[1]);  // columns >= Wl of the last strip are overwritten below
    reinterpret_cast<float2*>(r)[1] = make_float2(v[2], v[3]);  // or never read
    if (c0 == 0) { r[-2] = v[0]; r[-1] = v[0]; }
    const int last = Wl - 1 - c0;  // 0..3 if this strip holds the last column
    if (last >= 0 && last < kStripW) {
        const float vl = last == 0 ? v[0] : (last == 1 ? v[1] : (last == 2 ? v[2] : v[3]));
        r[last + 1] = vl; r[last + 2] = vl;
    }
}

// generic sizes (7 -> 4, odd detection sizes) and nearest mode: table driven; items = 4 columns x rows of level l-1
RC_HD void w_up_add(float* __restrict__ dstS, int pitch, int Hd, int Wd, const float* __restrict__ T, int tp, int Hl, int Wl,
                    const IdxLam* __restrict__ ytab, const IdxLam* __restrict__ xtab, int mode, const WGrid& gr, int jl, int LPP) {
    for (int rd = 0; rd < gr.rounds; ++rd) {
        const WItem it = w_item(gr, jl, LPP, rd, Hd);
        if (!it.valid) continue;
        const int j0 = it.c0;
        int x0[kStripW], x1[kStripW];
        float lx[kStripW];
#pragma unroll
        for (int c = 0; c < kStripW; ++c) {
            const int j = (j0 + c) < Wd ? (j0 + c) : (Wd - 1);
            const IdxLam t = xtab[j];
            x0[c] = t.i0;
            x1[c] = (mode == 1) ? t.i0 : t.i0 + (t.i0 < Wl - 1 ? 1 : 0);
            lx[c] = t.lam;
        }
        for (int i = it.r0; i < it.r0 + it.nrows; ++i) {
            const IdxLam ty = ytab[i];
            float* d = dstS + i * pitch + j0;
            const float* t0 = T + ty.i0 * tp;
            if (mode == 1) {
#pragma unroll
                for (int c = 0; c < kStripW; ++c)
                    if (j0 + c < Wd) d[c] += t0[x0[c]];
            } else {
                const float* t1 = T + (ty.i0 + (ty.i0 < Hl - 1 ? 1 : 0)) * tp;
                const float ly = ty.lam, hy = 1.f - ly;
#pragma unroll
                for (int c = 0; c < kStripW; ++c)
                    if (j0 + c < Wd) {
                        const float hx = 1.f - lx[c];
                        d[c] += hy * (hx * t0[x0[c]] + lx[c] * t0[x1[c]]) + ly * (hx * t1[x0[c]] + lx[c] * t1[x1[c]]);
                    }
            }
        }
    }
}

// 4 results of one output row -> global memory (element type T), straight from registers
template <typename T>
RC_HD void w_store_global4(T* __restrict__ dst_plane, int W, int row, int c0, const float (&v)[kStripW]) {
    T* d = dst_plane + (long)row * W + c0;
    if (c0 + kStripW <= W && (W & 3) == 0) {
        alignas(16) T tmp[kStripW];
#pragma unroll
        for (int c = 0; c < kStripW; ++c) tmp[c] = Elem<T>::from_f(v[c]);
        if (sizeof(T) == 4) *reinterpret_cast<float4*>(d) = *reinterpret_cast<const float4*>(tmp);
        else *reinterpret_cast<float2*>(d) = *reinterpret_cast<const float2*>(tmp);
    } else {
#pragma unroll
        for (int c = 0; c < kStripW; ++c)
            if (c0 + c < W) d[c] = Elem<T>::from_f(v[c]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward stages
// ---------------------------------------------------------------------------------------------------------
// transpose of the interpolation as a GATHER over the source pixels of level l of this lane's plane
// (deterministic; see rc_build_gather_table).  gsrc: INTERIOR origin of the level l-1 gradient (pitch gpitch; reads
// may run up to 3 elements past a row / the last row: finite cells with weight 0).  dstGT: padded buffer origin.
RC_HD void w_up_bwd(float* __restrict__ dstGT, int pitch, int pad, int Hl, int Wl, const float* __restrict__ gsrc, int gpitch,
                    const GatherEntry* __restrict__ gy, const GatherEntry* __restrict__ gx, unsigned magic_W, int jl, int LPP) {
    const int n = Hl * Wl;
    for (int idx = jl; idx < n; idx += LPP) {
        const int iy = rc_fastdiv(idx, magic_W), ix = idx - iy * Wl;
        const GatherEntry ey = gy[iy], ex = gx[ix];
        const float* p = gsrc + ey.d0 * gpitch + ex.d0;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float* row = p + a * gpitch;
            const float r = ex.w[0] * row[0] + ex.w[1] * row[1] + ex.w[2] * row[2] + ex.w[3] * row[3];
            acc = fmaf(ey.w[a], r, acc);
        }
        dstGT[(iy + pad) * pitch + ix + pad] = acc;
    }
}

// filter gradient of a stride-1 depthwise conv over this lane's items:
//   acc[r*K+s] += sum S[i+r, j+s] * G[i, j]; acc[K*K] += sum G      (S, G: padded buffers of one level geometry)
template <int K>
RC_HD void w_wgrad_s1(const float* __restrict__ S, const float* __restrict__ G, int pitch, int Ho, const WGrid& gr, int jl, int LPP,
                      float (&acc)[K * K + 1]) {
    constexpr int WL = (kStripW + 2 * (K / 2) + 3) & ~3;
    constexpr int PAD = K / 2;
    for (int rd = 0; rd < gr.rounds; ++rd) {
        const WItem it = w_item(gr, jl, LPP, rd, Ho);
        if (!it.valid) continue;
        const float* base = S + it.r0 * pitch + it.c0;
        const float* gbase = G + (it.r0 + PAD) * pitch + it.c0 + PAD;
        float win[K][WL];
#pragma unroll
        for (int r = 0; r < K - 1; ++r) rc_load_row<WL>(win[r], base + r * pitch);
        for (int o = 0; o < it.nrows; o += K) {
#pragma unroll
            for (int ph = 0; ph < K; ++ph) {
                if (o + ph < it.nrows) {
                    rc_load_row<WL>(win[(ph + K - 1) % K], base + (o + ph + K - 1) * pitch);
                    float gv[kStripW];
#pragma unroll
                    for (int c = 0; c < kStripW; ++c) gv[c] = gbase[(o + ph) * pitch + c];  // zero beyond Wo (padding)
#pragma unroll
                    for (int r = 0; r < K; ++r)
#pragma unroll
                        for (int s = 0; s < K; ++s)
#pragma unroll
                            for (int c = 0; c < kStripW; ++c)
                                acc[r * K + s] = fmaf(win[(ph + r) % K][c + s], gv[c], acc[r * K + s]);
                    acc[K * K] += (gv[0] + gv[1]) + (gv[2] + gv[3]);
                }
            }
        }
    }
}

// filter gradient of the stride-2 `down` conv: X = padded level l-1 input, Gr = padded total gradient of x_l
template <int K>
RC_HD void w_wgrad_s2(const float* __restrict__ X, int xpitch, const float* __restrict__ Gr, int gpitch, int Ho, const WGrid& gr,
                      int jl, int LPP, float (&acc)[K * K + 1]) {
    constexpr int WL = (2 * (kStripW - 1) + K + 3) & ~3;
    constexpr int PAD = K / 2;
    for (int rd = 0; rd < gr.rounds; ++rd) {
        const WItem it = w_item(gr, jl, LPP, rd, Ho);
        if (!it.valid) continue;
        const float* base = X + 2 * it.r0 * xpitch + 2 * it.c0;
        const float* gbase = Gr + (it.r0 + PAD) * gpitch + it.c0 + PAD;
        float win[K][WL];
#pragma unroll
        for (int r = 0; r < K - 2; ++r) rc_load_row<WL>(win[r], base + r * xpitch);
        for (int o = 0; o < it.nrows; o += K) {
#pragma unroll
            for (int ph = 0; ph < K; ++ph) {
                if (o + ph < it.nrows) {
                    const float* rowp = base + (2 * (o + ph) + K - 2) * xpitch;
                    rc_load_row<WL>(win[(2 * ph + K - 2) % K], rowp);
                    rc_load_row<WL>(win[(2 * ph + K - 1) % K], rowp + xpitch);
                    float gv[kStripW];
#pragma unroll
                    for (int c = 0; c < kStripW; ++c) gv[c] = gbase[(o + ph) * gpitch + c];
#pragma unroll
                    for (int r = 0; r < K; ++r)
#pragma unroll
                        for (int s = 0; s < K; ++s)
#pragma unroll
                            for (int c = 0; c < kStripW; ++c)
                                acc[r * K + s] = fmaf(win[(2 * ph + r) % K][2 * c + s], gv[c], acc[r * K + s]);
                    acc[K * K] += (gv[0] + gv[1]) + (gv[2] + gv[3]);
                }
            }
        }
    }
}

// transpose of the stride-2 `down` conv, gathered per 2x2 block of level l-1 so that every tap parity is static:
//   out[i, j] = sum_{r,s : (i+PAD-r), (j+PAD-s) even} w[r,s] * G[(i+PAD-r)/2, (j+PAD-s)/2].
// An item is a column pair b and a block of row pairs; the (PAD+1) x (PAD+1) window of G rolls down the rows.
// Gr: padded total gradient of x_l (origin).  epi(i, j, value) for every valid output of level l-1 (Ho x Wo).
template <int K, class Epi>
RC_HD void w_convT_s2(const float* __restrict__ Gr, int gpitch, const float* __restrict__ wslot, int Ho, int Wo, const WGrid& gr,
                      int jl, int LPP, Epi epi) {
    constexpr int PAD = K / 2;
    constexpr int LO = -(PAD / 2);
    constexpr int NW = PAD + 1;
    if (jl >= gr.ipp) return;
    float w[K * K];
    rc_load_filter<K * K>(w, wslot, false);
    const int nrp = (Ho + 1) / 2;
    for (int rd = 0; rd < gr.rounds; ++rd) {
        const int j = jl + rd * LPP;
        if (j >= gr.ipp) break;
        const int rb = rc_fastdiv(j, gr.m_spr), b = j - rb * gr.spr;
        if (b >= gr.strips) continue;
        const int a0 = rb * gr.rpb, a1 = (a0 + gr.rpb) < nrp ? (a0 + gr.rpb) : nrp;
        const float* gp = Gr + (a0 + LO + PAD) * gpitch + (b + LO + PAD);
        float gw[NW][NW];
#pragma unroll
        for (int r = 0; r < NW - 1; ++r)
#pragma unroll
            for (int s = 0; s < NW; ++s) gw[r + 1][s] = gp[r * gpitch + s];
        for (int a = a0; a < a1; ++a) {
#pragma unroll
            for (int r = 0; r < NW - 1; ++r)
#pragma unroll
                for (int s = 0; s < NW; ++s) gw[r][s] = gw[r + 1][s];
#pragma unroll
            for (int s = 0; s < NW; ++s) gw[NW - 1][s] = gp[(a - a0 + NW - 1) * gpitch + s];
#pragma unroll
            for (int di = 0; di < 2; ++di)
#pragma unroll
                for (int dj = 0; dj < 2; ++dj) {
                    float sum = 0.f;
#pragma unroll
                    for (int r = 0; r < K; ++r) {
                        if (((di + PAD - r) & 1) != 0) continue;
#pragma unroll
                        for (int s = 0; s < K; ++s) {
                            if (((dj + PAD - s) & 1) != 0) continue;
                            sum = fmaf(w[r * K + s], gw[(di + PAD - r) / 2 - LO][(dj + PAD - s) / 2 - LO], sum);
                        }
                    }
                    const int i = 2 * a + di, jj = 2 * b + dj;
                    if (i < Ho && jj < Wo) epi(i, jj, sum);
                }
        }
    }
}

}  // namespace recnext
