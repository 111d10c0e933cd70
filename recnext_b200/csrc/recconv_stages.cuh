// recconv_stages.cuh — the per-stage work of the fused RecConv kernels, written once for device and host.
//
// Every function here is a "parallel-for over the g lanes of one plane" (or over all threads of the CTA) with
// no communication inside a stage; stages are separated by barriers in recconv_body.cuh.  They are RC_HD so
// that tests can run exactly this code on the CPU (tests/emu) — the CUDA build never uses the host versions.
//
// Reference semantics (file:line relative to the reference tree):
//   model/recnext.py:21     down  = depthwise KxK, stride 2, pad K/2 (one filter shared by all levels)
//   model/recnext.py:22     convs = depthwise KxK, stride 1, pad K/2
//   model/recnext.py:27-34  forward chain; F.interpolate(size=s, mode) with ATen's align_corners=False index
//                           math (ATen/native/UpSample.h:259-311, 441-476; cuda/UpSample.cuh:96-145).
#pragma once
#include "recconv_plan.h"

#if defined(__CUDACC__)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#endif
#include <math.h>

namespace recnext {

struct IdxLam { int i0; float lam; };

// ---------------------------------------------------------------------------------------------------------
// interpolation source indices — bit-exact contract with ATen.
//   bilinear: scale = (float)in/out; src = fma(scale, dst + 0.5, -0.5) clamped at 0; i0 = (int)src;
//             i1 = i0 + (i0 < in-1); lambda = src - i0.   torch evaluates the source coordinate with ONE fused
//             multiply-add on both CPU and CUDA builds (probe in DESIGN.md), hence the explicit fmaf.
//   nearest : min((int)floorf(dst * scale), in - 1), with ATen's exact-2x and identity shortcuts.
// ---------------------------------------------------------------------------------------------------------
RC_HD void rc_bilinear_src(int in_size, int out_size, int dst, int& i0, int& i1, float& lam) {
    if (in_size == out_size) { i0 = dst; i1 = dst; lam = 0.f; return; }
    const float scale = (float)in_size / (float)out_size;
    float src = fmaf(scale, (float)dst + 0.5f, -0.5f);
    src = src < 0.f ? 0.f : src;
    int i = (int)src;  // src >= 0: truncation == floor
    if (i > in_size - 1) i = in_size - 1;
    float l = src - (float)i;
    l = l < 0.f ? 0.f : (l > 1.f ? 1.f : l);
    i0 = i; i1 = i + (i < in_size - 1 ? 1 : 0); lam = l;
}
RC_HD int rc_nearest_src(int in_size, int out_size, int dst) {
    if (in_size == out_size) return dst;
    if (out_size == 2 * in_size) return dst >> 1;
    const float scale = (float)in_size / (float)out_size;
#if defined(__CUDA_ARCH__)
    const int s = (int)floorf(__fmul_rn((float)dst, scale));
#else
    volatile float prod = (float)dst * scale;  // one rounding, never contracted
    const int s = (int)floorf(prod);
#endif
    return s < in_size - 1 ? s : in_size - 1;
}

// ---------------------------------------------------------------------------------------------------------
// element conversion
// ---------------------------------------------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
    static RC_HD float to_f(float v) { return v; }
    static RC_HD float from_f(float v) { return v; }
};
#if defined(__CUDACC__)
template <> struct Elem<__nv_bfloat16> {
    static RC_HD float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static RC_HD __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Elem<__half> {
    static RC_HD float to_f(__half v) { return __half2float(v); }
    static RC_HD __half from_f(float v) { return __float2half_rn(v); }
};
#endif

// filter element `i` of a parameter tensor stored as wdtype (0 f32, 1 bf16, 2 f16)
RC_HD float rc_load_param(const void* p, int wdtype, long i) {
#if defined(__CUDACC__)
    if (wdtype == 1) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    if (wdtype == 2) return __half2float(reinterpret_cast<const __half*>(p)[i]);
#endif
    return reinterpret_cast<const float*>(p)[i];
}

template <int N>
RC_HD void rc_load_row(float (&dst)[N], const float* __restrict__ src) {
    static_assert(N % 4 == 0, "row windows are whole float4s");
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(src + 4 * q);
        dst[4 * q + 0] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Stage: raw planes of one unit (element type T, unpadded, planes back to back) -> padded fp32 interiors.
// Rows are cut into chunks of V elements (V = largest power of two dividing W, <= 16 bytes), so every chunk is
// one aligned vector load and never straddles a row.  `nl` lanes of the unit share the chunks.
// ---------------------------------------------------------------------------------------------------------
template <typename T, int V>
RC_HD void rc_unpack_rows(const T* __restrict__ raw, int nrows_total, int H, int W, unsigned magic_cpr, unsigned magic_H,
                          float* __restrict__ planes, int plane_floats, int off, int pitch, int pad, int lane, int nl) {
    const int cpr = W / V;
    const int nchunks = nrows_total * cpr;
    for (int ci = lane; ci < nchunks; ci += nl) {
        const int r = rc_fastdiv(ci, magic_cpr);       // row within the unit (plane-major)
        const int col = (ci - r * cpr) * V;
        const int q = rc_fastdiv(r, magic_H);          // plane within the unit
        const int i = r - q * H;
        alignas(16) T vals[V];
        if (V * sizeof(T) == 16) *reinterpret_cast<float4*>(vals) = *reinterpret_cast<const float4*>(raw + (long)r * W + col);
        else if (V * sizeof(T) == 8) *reinterpret_cast<float2*>(vals) = *reinterpret_cast<const float2*>(raw + (long)r * W + col);
        else if (V * sizeof(T) == 4) *reinterpret_cast<float*>(vals) = *reinterpret_cast<const float*>(raw + (long)r * W + col);
        else vals[0] = raw[(long)r * W + col];
        float* dst = planes + (long)q * plane_floats + off + (i + pad) * pitch + pad + col;
#pragma unroll
        for (int e = 0; e < V; ++e) dst[e] = Elem<T>::to_f(vals[e]);
    }
}

template <typename T>
RC_HD void rc_unpack_unit(const T* raw, int nplanes, int H, int W, int vec, unsigned magic_cpr, unsigned magic_H, float* planes,
                          int plane_floats, int off, int pitch, int pad, int lane, int nl) {
    const int rows = nplanes * H;
    switch (vec * (int)sizeof(T)) {
        case 16: rc_unpack_rows<T, 16 / (int)sizeof(T)>(raw, rows, H, W, magic_cpr, magic_H, planes, plane_floats, off, pitch, pad, lane, nl); break;
        case 8: rc_unpack_rows<T, 8 / (int)sizeof(T)>(raw, rows, H, W, magic_cpr, magic_H, planes, plane_floats, off, pitch, pad, lane, nl); break;
        case 4: rc_unpack_rows<T, 4 / (int)sizeof(T)>(raw, rows, H, W, magic_cpr, magic_H, planes, plane_floats, off, pitch, pad, lane, nl); break;
        default: rc_unpack_rows<T, 1>(raw, rows, H, W, magic_cpr, magic_H, planes, plane_floats, off, pitch, pad, lane, nl); break;
    }
}

template <int N>
RC_HD void rc_load_filter(float (&w)[N], const float* __restrict__ wsm, bool flip) {
    // wsm slots are 16-byte aligned and padded to a multiple of 4 floats (plan.wstride)
    constexpr int N4 = (N + 3) / 4;
    float tmp[N4 * 4];
#pragma unroll
    for (int q = 0; q < N4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(wsm + 4 * q);
        tmp[4 * q + 0] = v.x; tmp[4 * q + 1] = v.y; tmp[4 * q + 2] = v.z; tmp[4 * q + 3] = v.w;
    }
    if (flip) {
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] = tmp[N - 1 - i];
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] = tmp[i];
    }
}

// ---------------------------------------------------------------------------------------------------------
// Stage: depthwise KxK stride-1 cross-correlation over a padded level buffer (rolling K-row register window).
// FLIP = true gives the transpose (input-gradient) of the stride-1 conv.  epi(row, col0, acc[4]) stores.
// ---------------------------------------------------------------------------------------------------------
template <int K, bool FLIP, class Epi>
RC_HD void rc_conv_s1(const float* __restrict__ src, int pitch, const float* __restrict__ wsm, bool use_bias, int Ho,
                      const StripGrid& sg, int lane, int g, Epi epi) {
    constexpr int WL = (kStripW + 2 * (K / 2) + 3) & ~3;
    const int nstrips = sg.nstrips, nitems = sg.nitems, rpi = sg.rpi;
    const unsigned magic_strips = sg.magic;
    if (lane >= nitems) return;
    float w[K * K];
    rc_load_filter<K * K>(w, wsm, FLIP);
    const float bias = use_bias ? wsm[K * K] : 0.f;
    for (int item = lane; item < nitems; item += g) {
        const int rb = rc_fastdiv(item, magic_strips), st = item - rb * nstrips;
        const int r0 = rb * rpi, c0 = st * kStripW;
        const int nrows = (Ho - r0) < rpi ? (Ho - r0) : rpi;
        const float* base = src + r0 * pitch + c0;
        float win[K][WL];
#pragma unroll
        for (int r = 0; r < K - 1; ++r) rc_load_row<WL>(win[r], base + r * pitch);
        for (int o = 0; o < nrows; o += K) {
#pragma unroll
            for (int ph = 0; ph < K; ++ph) {
                if (o + ph < nrows) {
                    rc_load_row<WL>(win[(ph + K - 1) % K], base + (o + ph + K - 1) * pitch);
                    float acc[kStripW] = {bias, bias, bias, bias};
#pragma unroll
                    for (int r = 0; r < K; ++r)
#pragma unroll
                        for (int s = 0; s < K; ++s)
#pragma unroll
                            for (int c = 0; c < kStripW; ++c) acc[c] = fmaf(w[r * K + s], win[(ph + r) % K][c + s], acc[c]);
                    epi(r0 + o + ph, c0, acc);
                }
            }
        }
    }
}

// Stage: depthwise KxK STRIDE-2 cross-correlation (the shared `down` filter): level l-1 (padded) -> level l.
template <int K, class Epi>
RC_HD void rc_conv_s2(const float* __restrict__ src, int pitch, const float* __restrict__ wsm, bool use_bias, int Ho,
                      const StripGrid& sg, int lane, int g, Epi epi) {
    constexpr int WL = (2 * (kStripW - 1) + K + 3) & ~3;
    const int nstrips = sg.nstrips, nitems = sg.nitems, rpi = sg.rpi;
    const unsigned magic_strips = sg.magic;
    if (lane >= nitems) return;
    float w[K * K];
    rc_load_filter<K * K>(w, wsm, false);
    const float bias = use_bias ? wsm[K * K] : 0.f;
    for (int item = lane; item < nitems; item += g) {
        const int rb = rc_fastdiv(item, magic_strips), st = item - rb * nstrips;
        const int r0 = rb * rpi, c0 = st * kStripW;
        const int nrows = (Ho - r0) < rpi ? (Ho - r0) : rpi;
        const float* base = src + 2 * r0 * pitch + 2 * c0;
        float win[K][WL];
#pragma unroll
        for (int r = 0; r < K - 2; ++r) rc_load_row<WL>(win[r], base + r * pitch);
        for (int o = 0; o < nrows; o += K) {
#pragma unroll
            for (int ph = 0; ph < K; ++ph) {
                if (o + ph < nrows) {
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
                    epi(r0 + o + ph, c0, acc);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// T buffer: the conv output t_l awaiting interpolation, (Hl+2) x tpitch floats with the interior at (+1,+1)
// and a REPLICATE border of one pixel.  With that border the exact-2x bilinear case of F.interpolate
// (align_corners=False) is the fixed 0.25/0.75 stencil, including ATen's clamping at the image edges.
// ---------------------------------------------------------------------------------------------------------
RC_HD void rc_store_T(float* __restrict__ T, int tp, int Hl, int Wl, int row, int c0, const float (&v)[kStripW]) {
    float* r = T + (row + 1) * tp + 1 + c0;
    const int last = Wl - 1 - c0;  // 0..3 if this strip holds the last column
    const float vl = last == 0 ? v[0] : (last == 1 ? v[1] : (last == 2 ? v[2] : v[3]));
    const bool top = row == 0, bot = row == Hl - 1;
#pragma unroll
    for (int c = 0; c < kStripW; ++c)
        if (c0 + c < Wl) {
            r[c] = v[c];
            if (top) r[c - tp] = v[c];
            if (bot) r[c + tp] = v[c];
        }
    if (c0 == 0) {
        r[-1] = v[0];
        if (top) r[-1 - tp] = v[0];
        if (bot) r[-1 + tp] = v[0];
    }
    if (last >= 0 && last < kStripW) {
        r[last + 1] = vl;
        if (top) r[last + 1 - tp] = vl;
        if (bot) r[last + 1 + tp] = vl;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Stage: s_{l-1} = x_{l-1} + interpolate(t_l, size = level l-1)   (model/recnext.py:33, with the `f + x` of the
// next loop iteration folded in).  dstS is the padded level l-1 buffer, updated in place.
// Fast path: both dimensions exactly 2x, bilinear.  Items = (4 output columns) x (rpu source rows).
// ---------------------------------------------------------------------------------------------------------
RC_HD void rc_hinterp2x(float (&h)[kStripW], const float* __restrict__ trow) {
    // trow points at padded column m0 (= source column m0 - 1); m0 is even, so two aligned float2 loads
    const float2 a = *reinterpret_cast<const float2*>(trow);
    const float2 b = *reinterpret_cast<const float2*>(trow + 2);
    h[0] = 0.75f * a.y + 0.25f * a.x;
    h[1] = 0.75f * a.y + 0.25f * b.x;
    h[2] = 0.75f * b.x + 0.25f * a.y;
    h[3] = 0.75f * b.x + 0.25f * b.y;
}

RC_HD void rc_upsample2x_add(float* __restrict__ dstS, int pitch, int pad, int Hd, int Wd, const float* __restrict__ T, int tp,
                             int Hl, const StripGrid& sg, int lane, int g) {
    const int nstrips = sg.nstrips, nitems = sg.nitems, rpu = sg.rpi;
    const unsigned magic_strips = sg.magic;
    for (int item = lane; item < nitems; item += g) {
        const int rb = rc_fastdiv(item, magic_strips), q = item - rb * nstrips;
        const int m_begin = rb * rpu;
        const int m_end = (m_begin + rpu) < Hl ? (m_begin + rpu) : Hl;
        const int j0 = q * kStripW;
        const float* tcol = T + 2 * q;  // padded column index of source column 2q - 1
        float hp[kStripW], hc[kStripW], hn[kStripW];
        rc_hinterp2x(hp, tcol + (m_begin) * tp);      // source row m_begin - 1 (padded row m_begin)
        rc_hinterp2x(hc, tcol + (m_begin + 1) * tp);  // source row m_begin
        float* d = dstS + (2 * m_begin + pad) * pitch + pad + j0;
        for (int m = m_begin; m < m_end; ++m) {
            rc_hinterp2x(hn, tcol + (m + 2) * tp);    // source row m + 1
            if (((pad | pitch) & 1) == 0 && j0 + kStripW <= Wd) {  // aligned float2 read-modify-write
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const float (&ho)[kStripW] = rr ? hn : hp;
                    float2* d2 = reinterpret_cast<float2*>(d + rr * pitch);
                    float2 u = d2[0], v = d2[1];
                    u.x += 0.75f * hc[0] + 0.25f * ho[0]; u.y += 0.75f * hc[1] + 0.25f * ho[1];
                    v.x += 0.75f * hc[2] + 0.25f * ho[2]; v.y += 0.75f * hc[3] + 0.25f * ho[3];
                    d2[0] = u; d2[1] = v;
                }
            } else {
#pragma unroll
                for (int c = 0; c < kStripW; ++c)
                    if (j0 + c < Wd) {
                        d[c] += 0.75f * hc[c] + 0.25f * hp[c];          // output row 2m
                        d[pitch + c] += 0.75f * hc[c] + 0.25f * hn[c];  // output row 2m + 1
                    }
            }
            d += 2 * pitch;
#pragma unroll
            for (int c = 0; c < kStripW; ++c) { hp[c] = hc[c]; hc[c] = hn[c]; }
        }
    }
}

// Generic path (odd sizes such as 7 -> 4, nearest mode): table driven.  Items = 4 columns x rpi rows of level l-1.
RC_HD void rc_upsample_add(float* __restrict__ dstS, int pitch, int pad, int Hd, int Wd, const float* __restrict__ T, int tp,
                           int Hl, int Wl, const IdxLam* __restrict__ ytab, const IdxLam* __restrict__ xtab, int mode,
                           const StripGrid& sg, int lane, int g) {
    const int nstrips = sg.nstrips, nitems = sg.nitems, rpi = sg.rpi;
    const unsigned magic_strips = sg.magic;
    for (int item = lane; item < nitems; item += g) {
        const int rb = rc_fastdiv(item, magic_strips), q = item - rb * nstrips;
        const int j0 = q * kStripW;
        const int i_end = (rb * rpi + rpi) < Hd ? (rb * rpi + rpi) : Hd;
        int x0[kStripW], x1[kStripW];
        float lx[kStripW];
#pragma unroll
        for (int c = 0; c < kStripW; ++c) {
            const int j = (j0 + c) < Wd ? (j0 + c) : (Wd - 1);
            const IdxLam t = xtab[j];
            x0[c] = t.i0 + 1;  // +1: T interior starts at padded column 1
            x1[c] = (mode == 1) ? x0[c] : t.i0 + (t.i0 < Wl - 1 ? 1 : 0) + 1;
            lx[c] = t.lam;
        }
        for (int i = rb * rpi; i < i_end; ++i) {
            const IdxLam ty = ytab[i];
            float* d = dstS + (i + pad) * pitch + pad + j0;
            const float* t0 = T + (ty.i0 + 1) * tp;
            if (mode == 1) {
#pragma unroll
                for (int c = 0; c < kStripW; ++c)
                    if (j0 + c < Wd) d[c] += t0[x0[c]];
            } else {
                const float* t1 = T + (ty.i0 + (ty.i0 < Hl - 1 ? 1 : 0) + 1) * tp;
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

// ---------------------------------------------------------------------------------------------------------
// Stage (bwd): transpose of the interpolation as a GATHER (deterministic): every source pixel of level l sums
// the (<= 4 x 4) destinations of level l-1 that read it, with the weights tabulated once per CTA.
// gsrc points at the INTERIOR origin of the level l-1 gradient (pitch gpitch; reads may run up to 3 elements
// past a row / the last row — those cells are finite and carry weight 0).  dst is the padded GT_l buffer.
// ---------------------------------------------------------------------------------------------------------
RC_HD void rc_upsample_bwd(float* __restrict__ dstGT, int pitch, int pad, int Hl, int Wl, const float* __restrict__ gsrc,
                           int gpitch, const GatherEntry* __restrict__ gy, const GatherEntry* __restrict__ gx,
                           unsigned magic_W, int lane, int g) {
    const int n = Hl * Wl;
    for (int idx = lane; idx < n; idx += g) {
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

// ---------------------------------------------------------------------------------------------------------
// Stage (bwd): weight gradient of a stride-1 depthwise conv: acc[r*K+s] += sum S[i+r, j+s] * G[i, j] over this
// lane's items; acc[K*K] += sum G (bias gradient).  S and G are padded buffers of the same level geometry.
// ---------------------------------------------------------------------------------------------------------
template <int K>
RC_HD void rc_wgrad_s1(const float* __restrict__ S, const float* __restrict__ G, int pitch, int Ho, const StripGrid& sg,
                       int lane, int g, float (&acc)[K * K + 1]) {
    constexpr int WL = (kStripW + 2 * (K / 2) + 3) & ~3;
    constexpr int PAD = K / 2;
    const int nstrips = sg.nstrips, nitems = sg.nitems, rpi = sg.rpi;
    const unsigned magic_strips = sg.magic;
    for (int item = lane; item < nitems; item += g) {
        const int rb = rc_fastdiv(item, magic_strips), st = item - rb * nstrips;
        const int r0 = rb * rpi, c0 = st * kStripW;
        const int nrows = (Ho - r0) < rpi ? (Ho - r0) : rpi;
        const float* base = S + r0 * pitch + c0;
        const float* gbase = G + (r0 + PAD) * pitch + c0 + PAD;
        float win[K][WL];
#pragma unroll
        for (int r = 0; r < K - 1; ++r) rc_load_row<WL>(win[r], base + r * pitch);
        for (int o = 0; o < nrows; o += K) {
#pragma unroll
            for (int ph = 0; ph < K; ++ph) {
                if (o + ph < nrows) {
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

// Weight gradient of the stride-2 `down` conv: X = padded level l-1 input, G = padded total gradient of x_l.
template <int K>
RC_HD void rc_wgrad_s2(const float* __restrict__ X, int xpitch, const float* __restrict__ G, int gpitch, int Ho,
                       const StripGrid& sg, int lane, int g, float (&acc)[K * K + 1]) {
    constexpr int WL = (2 * (kStripW - 1) + K + 3) & ~3;
    constexpr int PAD = K / 2;
    const int nstrips = sg.nstrips, nitems = sg.nitems, rpi = sg.rpi;
    const unsigned magic_strips = sg.magic;
    for (int item = lane; item < nitems; item += g) {
        const int rb = rc_fastdiv(item, magic_strips), st = item - rb * nstrips;
        const int r0 = rb * rpi, c0 = st * kStripW;
        const int nrows = (Ho - r0) < rpi ? (Ho - r0) : rpi;
        const float* base = X + 2 * r0 * xpitch + 2 * c0;
        const float* gbase = G + (r0 + PAD) * gpitch + c0 + PAD;
        float win[K][WL];
#pragma unroll
        for (int r = 0; r < K - 2; ++r) rc_load_row<WL>(win[r], base + r * xpitch);
        for (int o = 0; o < nrows; o += K) {
#pragma unroll
            for (int ph = 0; ph < K; ++ph) {
                if (o + ph < nrows) {
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

// ---------------------------------------------------------------------------------------------------------
// Stage (bwd): transpose of the stride-2 `down` conv, gathered per 2x2 output block so that every tap parity
// is static:  out[i, j] = base(i, j) + sum_{r,s : (i+PAD-r), (j+PAD-s) even} w[r,s] * G[(i+PAD-r)/2, (j+PAD-s)/2].
// G is the padded total gradient of x_l; outputs cover level l-1 (Ho x Wo).  epi(i, j, value_without_base).
// ---------------------------------------------------------------------------------------------------------
template <int K, class Epi>
RC_HD void rc_convT_s2(const float* __restrict__ G, int gpitch, const float* __restrict__ wsm, int Ho, int Wo,
                       const StripGrid& sg, int lane, int g, Epi epi) {
    constexpr int PAD = K / 2;
    constexpr int LO = -(PAD / 2);
    constexpr int NW = PAD + 1;
    const int nb = sg.nstrips, nitems = sg.nitems;
    if (lane >= nitems) return;
    float w[K * K];
    rc_load_filter<K * K>(w, wsm, false);
    for (int item = lane; item < nitems; item += g) {
        const int a = rc_fastdiv(item, sg.magic), b = item - a * nb;
        float gw[NW][NW];
        const float* gp = G + (a + LO + PAD) * gpitch + (b + LO + PAD);
#pragma unroll
        for (int r = 0; r < NW; ++r)
#pragma unroll
            for (int s = 0; s < NW; ++s) gw[r][s] = gp[r * gpitch + s];
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
                const int i = 2 * a + di, j = 2 * b + dj;
                if (i < Ho && j < Wo) epi(i, j, sum);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Tables (built once per CTA): forward {i0, lambda} per destination and, for the backward gather, the
// destination range per source.
// ---------------------------------------------------------------------------------------------------------
RC_HD void rc_build_fwd_table(IdxLam* tab, int in_size, int out_size, int mode, int tid, int nthreads) {
    for (int d = tid; d < out_size; d += nthreads) {
        IdxLam t;
        if (mode == 1) { t.i0 = rc_nearest_src(in_size, out_size, d); t.lam = 0.f; }
        else { int i1; rc_bilinear_src(in_size, out_size, d, t.i0, i1, t.lam); }
        tab[d] = t;
    }
}
// bwd: for every source index s of level l, the destinations d0..d0+n-1 (n <= 4 because out <= 2*in) of level
// l-1 whose interpolation reads s, and the weight each of them gives to s.  Returns false if n > 4.
RC_HD void rc_build_gather_table(GatherEntry* gat, const IdxLam* tab, int in_size, int out_size, int mode, int tid, int nthreads) {
    for (int s = tid; s < in_size; s += nthreads) {
        GatherEntry e;
        e.d0 = 0; e.n = 0; e.w[0] = e.w[1] = e.w[2] = e.w[3] = 0.f; e.pad_[0] = e.pad_[1] = 0;
        bool found = false;
        for (int d = 0; d < out_size; ++d) {
            const int i0 = tab[d].i0;
            const int i1 = mode == 1 ? i0 : i0 + (i0 < in_size - 1 ? 1 : 0);
            const float w = mode == 1 ? (i0 == s ? 1.f : 0.f) : ((i0 == s ? 1.f - tab[d].lam : 0.f) + (i1 == s ? tab[d].lam : 0.f));
            if (w != 0.f) {  // (a destination whose lambda is exactly 0 reads i1 with weight 0: not a reader)
                if (!found) { e.d0 = d; found = true; }
                const int k = d - e.d0;
                if (k < 4) {
                    if (k == 0) e.w[0] = w; else if (k == 1) e.w[1] = w; else if (k == 2) e.w[2] = w; else e.w[3] = w;
                }
                e.n = k + 1;
            }
        }
        // all four reads d0..d0+3 must stay inside the level (or inside the 4-row/col minimum the plan allocates)
        const int limit = out_size >= 4 ? out_size - 4 : 0;
        if (e.d0 > limit) {
            const int sh = e.d0 - limit;  // 1..3
            float w4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float wk = k == 0 ? e.w[0] : (k == 1 ? e.w[1] : (k == 2 ? e.w[2] : e.w[3]));
                const int t = k + sh;
                if (t == 1) w4[1] = wk; else if (t == 2) w4[2] = wk; else if (t == 3) w4[3] = wk;
            }
            e.w[0] = w4[0]; e.w[1] = w4[1]; e.w[2] = w4[2]; e.w[3] = w4[3];
            e.d0 = limit;
        }
        gat[s] = e;
    }
}

}  // namespace recnext
