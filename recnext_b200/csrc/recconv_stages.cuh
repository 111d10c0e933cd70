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
struct Range { int lo, hi; };

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
// Stage: raw plane group (element type T, unpadded, planes back to back) -> padded fp32 interiors.
// Flat over all threads of the CTA (any thread may touch any plane).
// ---------------------------------------------------------------------------------------------------------
template <typename T>
RC_HD void rc_unpack_group(const T* __restrict__ raw, int nplanes, int H, int W, float* __restrict__ planes,
                           int plane_floats, int off, int pitch, int pad, int tid, int nthreads) {
    constexpr int V = 16 / (int)sizeof(T);
    const int HW = H * W, total = nplanes * HW;
    for (int v = tid; v * V < total; v += nthreads) {
        const int idx = v * V;
        int q = idx / HW;
        const int rem = idx - q * HW;
        int i = rem / W, j = rem - i * W;
        alignas(16) T vals[V];
        *reinterpret_cast<float4*>(vals) = *reinterpret_cast<const float4*>(raw + idx);  // buffers are 128-B padded
        float* dst = planes + (long)q * plane_floats + off + (i + pad) * pitch + pad;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if (idx + e < total) {
                dst[j] = Elem<T>::to_f(vals[e]);
                if (++j == W) {
                    j = 0; dst += pitch;
                    if (++i == H) { i = 0; ++q; dst = planes + (long)q * plane_floats + off + pad * pitch + pad; }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Stage: depthwise KxK stride-1 cross-correlation over a padded level buffer (rolling K-row register window).
// FLIP = true gives the transpose (input-gradient) of the stride-1 conv.  epi(row, col0, acc[4]) stores.
// ---------------------------------------------------------------------------------------------------------
template <int K, bool FLIP, class Epi>
RC_HD void rc_conv_s1(const float* __restrict__ src, int pitch, const float* __restrict__ wsm, bool use_bias, int Ho,
                      int Wo, int rpi, int lane, int g, Epi epi) {
    constexpr int WL = (kStripW + 2 * (K / 2) + 3) & ~3;
    const int nstrips = (Wo + kStripW - 1) / kStripW;
    const int nitems = nstrips * ((Ho + rpi - 1) / rpi);
    if (lane >= nitems) return;
    float w[K * K];
#pragma unroll
    for (int i = 0; i < K * K; ++i) w[i] = FLIP ? wsm[K * K - 1 - i] : wsm[i];
    const float bias = use_bias ? wsm[K * K] : 0.f;
    for (int item = lane; item < nitems; item += g) {
        const int rb = item / nstrips, st = item - rb * nstrips;
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
                      int Wo, int rpi, int lane, int g, Epi epi) {
    constexpr int WL = (2 * (kStripW - 1) + K + 3) & ~3;
    const int nstrips = (Wo + kStripW - 1) / kStripW;
    const int nitems = nstrips * ((Ho + rpi - 1) / rpi);
    if (lane >= nitems) return;
    float w[K * K];
#pragma unroll
    for (int i = 0; i < K * K; ++i) w[i] = wsm[i];
    const float bias = use_bias ? wsm[K * K] : 0.f;
    for (int item = lane; item < nitems; item += g) {
        const int rb = item / nstrips, st = item - rb * nstrips;
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
// Stage: s_{l-1} = x_{l-1} + interpolate(t_l, size = level l-1)   (model/recnext.py:33, the `f + x` of the
// next iteration folded in).  T is unpadded [Hl x Wl]; dst is the padded level l-1 buffer, updated in place.
// ---------------------------------------------------------------------------------------------------------
RC_HD void rc_upsample_add(float* __restrict__ dstS, int pitch, int pad, int Hd, int Wd, const float* __restrict__ T,
                           int Hl, int Wl, const IdxLam* __restrict__ ytab, const IdxLam* __restrict__ xtab, int mode,
                           int lane, int g) {
    const int nstrips = (Wd + kStripW - 1) / kStripW;
    const int nitems = nstrips * Hd;
    for (int item = lane; item < nitems; item += g) {
        const int i = item / nstrips, c0 = (item - i * nstrips) * kStripW;
        const IdxLam ty = ytab[i];
        float* d = dstS + (i + pad) * pitch + pad;
        if (mode == 1) {
            const float* t0 = T + ty.i0 * Wl;
#pragma unroll
            for (int c = 0; c < kStripW; ++c) {
                const int j = c0 + c;
                if (j < Wd) d[j] += t0[xtab[j].i0];
            }
        } else {
            const int y1 = ty.i0 + (ty.i0 < Hl - 1 ? 1 : 0);
            const float ly = ty.lam, hy = 1.f - ly;
            const float* t0 = T + ty.i0 * Wl;
            const float* t1 = T + y1 * Wl;
#pragma unroll
            for (int c = 0; c < kStripW; ++c) {
                const int j = c0 + c;
                if (j < Wd) {
                    const IdxLam tx = xtab[j];
                    const int x1 = tx.i0 + (tx.i0 < Wl - 1 ? 1 : 0);
                    const float lx = tx.lam, hx = 1.f - lx;
                    const float v = hy * (hx * t0[tx.i0] + lx * t0[x1]) + ly * (hx * t1[tx.i0] + lx * t1[x1]);
                    d[j] += v;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Stage (bwd): transpose of the interpolation as a GATHER (deterministic): for every source pixel of level l
// sum the destinations of level l-1 that read it.  rng tables give the contiguous destination ranges.
// gsrc points at the INTERIOR origin of the level l-1 gradient (pitch gpitch); dst is the padded GT_l buffer.
// ---------------------------------------------------------------------------------------------------------
RC_HD float rc_interp_weight(const IdxLam t, int src, int in_size, int mode) {
    if (mode == 1) return t.i0 == src ? 1.f : 0.f;
    const int i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    return (t.i0 == src ? 1.f - t.lam : 0.f) + (i1 == src ? t.lam : 0.f);
}

RC_HD void rc_upsample_bwd(float* __restrict__ dstGT, int pitch, int pad, int Hl, int Wl, const float* __restrict__ gsrc,
                           int gpitch, const IdxLam* __restrict__ ytab, const IdxLam* __restrict__ xtab,
                           const Range* __restrict__ yr, const Range* __restrict__ xr, int mode, int lane, int g) {
    const int n = Hl * Wl;
    for (int idx = lane; idx < n; idx += g) {
        const int iy = idx / Wl, ix = idx - iy * Wl;
        const Range ry = yr[iy], rx = xr[ix];
        float acc = 0.f;
        for (int dy = ry.lo; dy <= ry.hi; ++dy) {
            const float wy = rc_interp_weight(ytab[dy], iy, Hl, mode);
            const float* grow = gsrc + dy * gpitch;
            float rowacc = 0.f;
            for (int dx = rx.lo; dx <= rx.hi; ++dx) rowacc = fmaf(rc_interp_weight(xtab[dx], ix, Wl, mode), grow[dx], rowacc);
            acc = fmaf(wy, rowacc, acc);
        }
        dstGT[(iy + pad) * pitch + ix + pad] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Stage (bwd): weight gradient of a stride-1 depthwise conv: acc[r*K+s] += sum S[i+r, j+s] * G[i, j] over this
// lane's items; acc[K*K] += sum G (bias gradient).  S and G are padded buffers of the same level geometry.
// ---------------------------------------------------------------------------------------------------------
template <int K>
RC_HD void rc_wgrad_s1(const float* __restrict__ S, const float* __restrict__ G, int pitch, int Ho, int Wo, int rpi,
                       int lane, int g, float (&acc)[K * K + 1]) {
    constexpr int WL = (kStripW + 2 * (K / 2) + 3) & ~3;
    constexpr int PAD = K / 2;
    const int nstrips = (Wo + kStripW - 1) / kStripW;
    const int nitems = nstrips * ((Ho + rpi - 1) / rpi);
    for (int item = lane; item < nitems; item += g) {
        const int rb = item / nstrips, st = item - rb * nstrips;
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
RC_HD void rc_wgrad_s2(const float* __restrict__ X, int xpitch, const float* __restrict__ G, int gpitch, int Ho, int Wo,
                       int rpi, int lane, int g, float (&acc)[K * K + 1]) {
    constexpr int WL = (2 * (kStripW - 1) + K + 3) & ~3;
    constexpr int PAD = K / 2;
    const int nstrips = (Wo + kStripW - 1) / kStripW;
    const int nitems = nstrips * ((Ho + rpi - 1) / rpi);
    for (int item = lane; item < nitems; item += g) {
        const int rb = item / nstrips, st = item - rb * nstrips;
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
RC_HD void rc_convT_s2(const float* __restrict__ G, int gpitch, const float* __restrict__ wsm, int Ho, int Wo, int lane,
                       int g, Epi epi) {
    constexpr int PAD = K / 2;
    constexpr int LO = -(PAD / 2);
    constexpr int NW = PAD + 1;
    const int na = (Ho + 1) >> 1, nb = (Wo + 1) >> 1;
    const int nitems = na * nb;
    if (lane >= nitems) return;
    float w[K * K];
#pragma unroll
    for (int i = 0; i < K * K; ++i) w[i] = wsm[i];
    for (int item = lane; item < nitems; item += g) {
        const int a = item / nb, b = item - a * nb;
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
RC_HD void rc_build_range_table(Range* rng, const IdxLam* tab, int in_size, int out_size, int mode, int tid, int nthreads) {
    for (int s = tid; s < in_size; s += nthreads) {
        Range r; r.lo = 0; r.hi = -1;
        bool found = false;
        for (int d = 0; d < out_size; ++d) {
            const int i0 = tab[d].i0;
            const int i1 = mode == 1 ? i0 : i0 + (i0 < in_size - 1 ? 1 : 0);
            if (i0 == s || i1 == s) { if (!found) { r.lo = d; found = true; } r.hi = d; }
        }
        rng[s] = r;
    }
}

}  // namespace recnext
