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
    it.valid = j < gr.ipp;
    const int rb = rc_fastdiv(j, gr.m_strips);
    it.c0 = (j - rb * gr.strips) * kStripW;
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
        if (!it.valid) break;
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
        if (!it.valid) break;
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

// masked store of 4 results into a padded level buffer (interior coordinates row, c0); pads stay zero
RC_HD void w_store_level(float* __restrict__ dst_interior, int pitch, int W, int row, int c0, const float (&v)[kStripW]) {
    float* d = dst_interior + row * pitch + c0;
    if (c0 + kStripW <= W) {
#pragma unroll
        for (int c = 0; c < kStripW; ++c) d[c] = v[c];
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
        if (!it.valid) break;
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

// generic sizes (7 -> 4, odd detection sizes) and nearest mode: table driven; items = 4 columns x rows of level l-1
RC_HD void w_up_add(float* __restrict__ dstS, int pitch, int Hd, int Wd, const float* __restrict__ T, int tp, int Hl, int Wl,
                    const IdxLam* __restrict__ ytab, const IdxLam* __restrict__ xtab, int mode, const WGrid& gr, int jl, int LPP) {
    for (int rd = 0; rd < gr.rounds; ++rd) {
        const WItem it = w_item(gr, jl, LPP, rd, Hd);
        if (!it.valid) break;
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
        if (!it.valid) break;
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
        if (!it.valid) break;
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
        const int rb = rc_fastdiv(j, gr.m_strips), b = j - rb * gr.strips;
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
