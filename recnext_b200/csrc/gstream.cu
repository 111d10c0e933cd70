// gstream.cu — STREAMED RecConv forward / backward for planes whose pyramid does not fit in one SM's shared memory
// (detection stages 0-1: [2,64,200,336] level 4, [2,128,100,168] level 3, in fp32 the forward too).
//
// The fused kernels keep a plane's whole pyramid on chip; a fused spatial tiling would need a halo of ~90 pixels per
// side at level 4 (SURVEY.md "Hard parts"), so planes that do not fit are walked level by level through a caller-owned
// fp32 workspace instead: one launch per stage of reference model/recnext.py:24-34 (down loop :27-29, up loop :31-33,
// final conv :34) and of its autograd graph.  Every stage is a plain grid-stride kernel over the
// elements of one pyramid level; intermediates are fp32 (so the fp32 1e-5 bar holds and the 16-bit results are at
// least as accurate as the reference's autocast graph).  Filter gradients are reduced per (image, channel) plane by
// one CTA and summed over images and levels in a fixed order by a finalize kernel: deterministic.
// The levels below the first one are small (1/4, 1/16, ..): traffic is ~14 N bytes instead of the fused 5 N e.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "recconv_stages.cuh"
#include "recconv_body.cuh"
#include "gstream.h"

namespace recnext {

namespace {

__device__ __forceinline__ float g_load(const void* p, int dtype, long i) {
    if (dtype == 1) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    if (dtype == 2) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    return reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void g_store(void* p, int dtype, long i, float v) {
    if (dtype == 1) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    else if (dtype == 2) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
    else reinterpret_cast<float*>(p)[i] = v;
}

// filters / biases of all convs as fp32: wf[(slot * C + c) * (KK + 1) + e], e == KK is the bias (0 when absent)
__global__ void g_prep_params(KernelArgs a, int nslots, int C, int KK, int wdtype, int has_bias, float* __restrict__ wf) {
    const long total = (long)nslots * C * (KK + 1);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int e = (int)(i % (KK + 1));
        const long sc = i / (KK + 1);
        const int c = (int)(sc % C), slot = (int)(sc / C);
        float v = 0.f;
        if (a.w[slot]) {
            if (e < KK) v = rc_load_param(a.w[slot], wdtype, (long)c * KK + e);
            else if (has_bias && a.b[slot]) v = rc_load_param(a.b[slot], wdtype, c);
        }
        wf[i] = v;
    }
}

// out[p, i, j] = bias + sum_{r,s} in[p, i*S + r - pad, j*S + s - pad] * w[c(p), r, s]     (zero padding)
template <int K, int S>
__global__ void g_conv(const void* __restrict__ in, int in_dtype, void* __restrict__ out, int out_dtype, const float* __restrict__ wf, int C, long planes,
                       int Hi, int Wi, int Ho, int Wo) {
    constexpr int PAD = K / 2, KK = K * K;
    const long total = planes * Ho * Wo;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % Wo);
        const long t = idx / Wo;
        const int i = (int)(t % Ho);
        const long p = t / Ho;
        const float* w = wf + (p % C) * (KK + 1);
        const long base = p * (long)Hi * Wi;
        float acc = w[KK];
#pragma unroll
        for (int r = 0; r < K; ++r) {
            const int y = i * S + r - PAD;
            if (y < 0 || y >= Hi) continue;
#pragma unroll
            for (int s = 0; s < K; ++s) {
                const int x = j * S + s - PAD;
                if (x >= 0 && x < Wi) acc = fmaf(g_load(in, in_dtype, base + (long)y * Wi + x), w[r * K + s], acc);
            }
        }
        g_store(out, out_dtype, idx, acc);
    }
}

// transpose of g_conv w.r.t. its input: gin[p, y, x] (+)= sum over (r, s, i, j) with i*S + r - pad == y, j*S + s - pad == x of gout[p,i,j] * w[r,s]
// add0 (nullable): an fp32 tensor of the shape of gin added to the result
template <int K, int S>
__global__ void g_convT(const void* __restrict__ gout, int gout_dtype, void* gin, int gin_dtype, const float* add0 /* may alias gin */,
                        const float* __restrict__ wf, int C, long planes, int Hi, int Wi, int Ho, int Wo) {
    constexpr int PAD = K / 2, KK = K * K;
    const long total = planes * Hi * Wi;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % Wi);
        const long t = idx / Wi;
        const int y = (int)(t % Hi);
        const long p = t / Hi;
        const float* w = wf + (p % C) * (KK + 1);
        const long base = p * (long)Ho * Wo;
        float acc = add0 ? add0[idx] : 0.f;
#pragma unroll
        for (int r = 0; r < K; ++r) {
            const int yy = y + PAD - r;
            if (yy < 0 || (yy % S) != 0) continue;
            const int i = yy / S;
            if (i >= Ho) continue;
#pragma unroll
            for (int s = 0; s < K; ++s) {
                const int xx = x + PAD - s;
                if (xx < 0 || (xx % S) != 0) continue;
                const int j = xx / S;
                if (j < Wo) acc = fmaf(g_load(gout, gout_dtype, base + (long)i * Wo + j), w[r * K + s], acc);
            }
        }
        g_store(gin, gin_dtype, idx, acc);
    }
}

// s[p, i, j] = base[p, i, j] + interpolate(t)[i, j]      (model/recnext.py:33 and the next `f + x`)
__global__ void g_upadd(const void* __restrict__ base, int base_dtype, const float* __restrict__ t, float* __restrict__ s, long planes, int Hs, int Ws,
                        int Hd, int Wd, int mode) {
    const long total = planes * Hd * Wd;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % Wd);
        const long q = idx / Wd;
        const int i = (int)(q % Hd);
        const long p = q / Hd;
        const float* tp = t + p * (long)Hs * Ws;
        float u;
        if (mode == 1) {
            u = tp[(long)rc_nearest_src(Hs, Hd, i) * Ws + rc_nearest_src(Ws, Wd, j)];
        } else {
            int y0, y1, x0, x1; float ly, lx;
            rc_bilinear_src(Hs, Hd, i, y0, y1, ly);
            rc_bilinear_src(Ws, Wd, j, x0, x1, lx);
            const float hy = 1.f - ly, hx = 1.f - lx;
            u = hy * (hx * tp[(long)y0 * Ws + x0] + lx * tp[(long)y0 * Ws + x1]) + ly * (hx * tp[(long)y1 * Ws + x0] + lx * tp[(long)y1 * Ws + x1]);
        }
        s[idx] = g_load(base, base_dtype, idx) + u;
    }
}

// weight of source index `src` in destination `d` of a 1-D interpolation (0 when d does not read src)
__device__ __forceinline__ float g_w1d(int in_size, int out_size, int d, int src, int mode) {
    if (mode == 1) return rc_nearest_src(in_size, out_size, d) == src ? 1.f : 0.f;
    int i0, i1; float lam;
    rc_bilinear_src(in_size, out_size, d, i0, i1, lam);
    return (i0 == src ? 1.f - lam : 0.f) + (i1 == src ? lam : 0.f);
}
// gt[p, a, b] = sum over destinations (i, j) reading source (a, b) of gs[p, i, j] * wy * wx     (transpose of interpolate, as a gather)
__global__ void g_upT(const float* __restrict__ gs, float* __restrict__ gt, long planes, int Hs, int Ws, int Hd, int Wd, int mode) {
    const long total = planes * Hs * Ws;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int b = (int)(idx % Ws);
        const long q = idx / Ws;
        const int a = (int)(q % Hs);
        const long p = q / Hs;
        // destinations that can read source a lie in ((a - 1) Hd / Hs - 1, (a + 2) Hd / Hs + 1) for either mode
        const int i_lo = max(0, (int)(((long)(a - 1) * Hd) / Hs) - 2), i_hi = min(Hd - 1, (int)(((long)(a + 2) * Hd + Hs - 1) / Hs) + 1);
        const int j_lo = max(0, (int)(((long)(b - 1) * Wd) / Ws) - 2), j_hi = min(Wd - 1, (int)(((long)(b + 2) * Wd + Ws - 1) / Ws) + 1);
        const float* gp = gs + p * (long)Hd * Wd;
        float acc = 0.f;
        for (int i = i_lo; i <= i_hi; ++i) {
            const float wy = g_w1d(Hs, Hd, i, a, mode);
            if (wy == 0.f) continue;
            float row = 0.f;
            for (int j = j_lo; j <= j_hi; ++j) {
                const float wx = g_w1d(Ws, Wd, j, b, mode);
                if (wx != 0.f) row = fmaf(wx, gp[(long)i * Wd + j], row);
            }
            acc = fmaf(wy, row, acc);
        }
        gt[idx] = acc;
    }
}

// partial filter gradient of one plane: part[p][e] = sum_{i,j} in[p, i*S + r - pad, j*S + s - pad] * g[p, i, j]  (e = r*K+s),
// part[p][KK] = sum g (bias).  One CTA per plane, fixed summation order.
template <int K, int S>
__global__ void __launch_bounds__(256) g_wgrad(const void* __restrict__ in, int in_dtype, const void* __restrict__ g, int g_dtype, float* __restrict__ part,
                                               int Hi, int Wi, int Ho, int Wo) {
    constexpr int PAD = K / 2, KK = K * K;
    const long p = blockIdx.x;
    const long ibase = p * (long)Hi * Wi, gbase = p * (long)Ho * Wo;
    float acc[KK + 1];
#pragma unroll
    for (int e = 0; e <= KK; ++e) acc[e] = 0.f;
    for (int idx = threadIdx.x; idx < Ho * Wo; idx += blockDim.x) {
        const int i = idx / Wo, j = idx - i * Wo;
        const float gv = g_load(g, g_dtype, gbase + idx);
        acc[KK] += gv;
#pragma unroll
        for (int r = 0; r < K; ++r) {
            const int y = i * S + r - PAD;
            if (y < 0 || y >= Hi) continue;
#pragma unroll
            for (int s = 0; s < K; ++s) {
                const int x = j * S + s - PAD;
                if (x >= 0 && x < Wi) acc[r * K + s] = fmaf(g_load(in, in_dtype, ibase + (long)y * Wi + x), gv, acc[r * K + s]);
            }
        }
    }
    __shared__ float red[8][KK + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e <= KK; ++e) {
        float v = acc[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][e] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e <= KK; e += blockDim.x) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][e];
        part[p * (KK + 1) + e] = v;
    }
}

// gw[slot][c][e] = sum over images (and, for slot 0 = `down`, over levels) of the plane partials, in a fixed order.
// partial sets: set 0..L-1 = `down` applied to level l (producing level l+1); set L + j = convs[j]
__global__ void g_wgrad_finalize(const float* __restrict__ part, float* __restrict__ gw, float* __restrict__ gb, int L, int B, int C, int KK) {
    const long total = (long)(L + 2) * C * (KK + 1);
    const long set_stride = (long)B * C * (KK + 1);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int e = (int)(i % (KK + 1));
        const long sc = i / (KK + 1);
        const int c = (int)(sc % C), slot = (int)(sc / C);
        const int set0 = slot == 0 ? 0 : L + slot - 1, nset = slot == 0 ? L : 1;
        float s = 0.f;
        for (int q = 0; q < nset; ++q)
            for (int n = 0; n < B; ++n) s += part[(set0 + q) * set_stride + ((long)n * C + c) * (KK + 1) + e];
        if (e < KK) gw[((long)slot * C + c) * KK + e] = s;
        else if (gb) gb[(long)slot * C + c] = s;
    }
}

inline int g_blocks(long total, int sms) {
    long b = (total + 255) / 256;
    const long cap = (long)sms * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

struct GLayout {
    int L, H[kMaxLevel + 1], W[kMaxLevel + 1];
    long n[kMaxLevel + 1];        // elements of level l over all planes
    long offX[kMaxLevel + 1];     // x_l, l >= 1
    long offS[kMaxLevel + 1];     // s_l = x_l + u_l, l < L (s_L is x_L)
    long offGS[kMaxLevel + 1];    // bwd: gradient w.r.t. s_l, later the total gradient of x_l
    long offT, offGT, offW, offPart, total;
};

GLayout g_layout(const GStreamDesc& d, bool bwd) {
    GLayout g{};
    g.L = d.L;
    g.H[0] = d.H; g.W[0] = d.W;
    for (int l = 1; l <= d.L; ++l) { g.H[l] = rc_down_size(g.H[l - 1], d.K); g.W[l] = rc_down_size(g.W[l - 1], d.K); }
    const long planes = (long)d.B * d.C;
    long off = 0;
    auto take = [&](long n) { const long o = off; off += (n + 63) / 64 * 64; return o; };
    for (int l = 0; l <= d.L; ++l) g.n[l] = planes * g.H[l] * g.W[l];
    for (int l = 1; l <= d.L; ++l) g.offX[l] = take(g.n[l]);
    for (int l = 0; l < d.L; ++l) g.offS[l] = take(g.n[l]);
    g.offS[d.L] = d.L > 0 ? g.offX[d.L] : -1;
    g.offT = take(d.L > 0 ? g.n[1] : 0);
    g.offW = take((long)(d.L + 2) * d.C * (d.K * d.K + 1));
    if (bwd) {
        for (int l = 0; l <= d.L; ++l) g.offGS[l] = take(g.n[l]);
        g.offGT = take(d.L > 0 ? g.n[1] : 0);
        g.offPart = take((long)(2 * d.L + 1) * planes * (d.K * d.K + 1));
    }
    g.total = off;
    return g;
}

template <int K>
cudaError_t g_run(const GStreamDesc& d, const KernelArgs& a, float* ws, bool bwd, float* gw, float* gb, cudaStream_t st) {
    constexpr int KK = K * K;
    const GLayout g = g_layout(d, bwd);
    const int L = d.L, C = d.C, sms = d.num_sms > 0 ? d.num_sms : 148;
    const long planes = (long)d.B * C;
    float* wf = ws + g.offW;
    auto WF = [&](int slot) { return wf + (long)slot * C * (KK + 1); };
    g_prep_params<<<g_blocks((long)(L + 2) * C * (KK + 1), sms), 256, 0, st>>>(a, L + 2, C, KK, d.wdtype, d.has_bias, wf);
    // ---- down loop (model/recnext.py:27-29)
    for (int l = 1; l <= L; ++l) {
        const void* in = l == 1 ? a.x : (const void*)(ws + g.offX[l - 1]);
        g_conv<K, 2><<<g_blocks(g.n[l], sms), 256, 0, st>>>(in, l == 1 ? d.dtype : 0, ws + g.offX[l], 0, WF(0), C, planes, g.H[l - 1], g.W[l - 1], g.H[l], g.W[l]);
    }
    // ---- up loop (:31-33): t_l = convs[L-l](s_l); s_{l-1} = x_{l-1} + interpolate(t_l)
    for (int l = L; l >= 1; --l) {
        g_conv<K, 1><<<g_blocks(g.n[l], sms), 256, 0, st>>>(ws + g.offS[l], 0, ws + g.offT, 0, WF(1 + L - l), C, planes, g.H[l], g.W[l], g.H[l], g.W[l]);
        const void* base = l == 1 ? a.x : (const void*)(ws + g.offX[l - 1]);
        g_upadd<<<g_blocks(g.n[l - 1], sms), 256, 0, st>>>(base, l == 1 ? d.dtype : 0, ws + g.offT, ws + g.offS[l - 1], planes, g.H[l], g.W[l], g.H[l - 1], g.W[l - 1], d.mode);
    }
    const void* s0 = L > 0 ? (const void*)(ws + g.offS[0]) : a.x;
    const int s0_dtype = L > 0 ? 0 : d.dtype;
    if (!bwd) {
        // ---- y = convs[L](s_0)  (:34)
        g_conv<K, 1><<<g_blocks(g.n[0], sms), 256, 0, st>>>(s0, s0_dtype, a.out, d.dtype, WF(1 + L), C, planes, d.H, d.W, d.H, d.W);
        return cudaGetLastError();
    }
    // ---- backward: final conv
    float* part = ws + g.offPart;
    const long pstride = planes * (KK + 1);
    auto PART = [&](int set) { return part + (long)set * pstride; };
    g_wgrad<K, 1><<<(unsigned)planes, 256, 0, st>>>(s0, s0_dtype, a.gy, d.dtype, PART(L + L), d.H, d.W, d.H, d.W);
    if (L == 0) {
        g_convT<K, 1><<<g_blocks(g.n[0], sms), 256, 0, st>>>(a.gy, d.dtype, a.out, d.dtype, nullptr, WF(1), C, planes, d.H, d.W, d.H, d.W);
    } else {
        g_convT<K, 1><<<g_blocks(g.n[0], sms), 256, 0, st>>>(a.gy, d.dtype, ws + g.offGS[0], 0, nullptr, WF(1 + L), C, planes, d.H, d.W, d.H, d.W);
        // per level: gt_l = interpolate^T(gs_{l-1}); dK_{L-l} = corr(s_l, gt_l); gs_l = K_{L-l}^T(gt_l)
        for (int l = 1; l <= L; ++l) {
            g_upT<<<g_blocks(g.n[l], sms), 256, 0, st>>>(ws + g.offGS[l - 1], ws + g.offGT, planes, g.H[l], g.W[l], g.H[l - 1], g.W[l - 1], d.mode);
            g_wgrad<K, 1><<<(unsigned)planes, 256, 0, st>>>(ws + g.offS[l], 0, ws + g.offGT, 0, PART(L + L - l), g.H[l], g.W[l], g.H[l], g.W[l]);
            g_convT<K, 1><<<g_blocks(g.n[l], sms), 256, 0, st>>>(ws + g.offGT, 0, ws + g.offGS[l], 0, nullptr, WF(1 + L - l), C, planes, g.H[l], g.W[l], g.H[l], g.W[l]);
        }
        // down chain: G_L = gs_L; dD += corr_s2(x_{l-1}, G_l); G_{l-1} = gs_{l-1} + D^T(G_l); gx = G_0
        for (int l = L; l >= 1; --l) {
            const void* xin = l == 1 ? a.x : (const void*)(ws + g.offX[l - 1]);
            g_wgrad<K, 2><<<(unsigned)planes, 256, 0, st>>>(xin, l == 1 ? d.dtype : 0, ws + g.offGS[l], 0, PART(l - 1), g.H[l - 1], g.W[l - 1], g.H[l], g.W[l]);
            void* dst = l == 1 ? a.out : (void*)(ws + g.offGS[l - 1]);
            g_convT<K, 2><<<g_blocks(g.n[l - 1], sms), 256, 0, st>>>(ws + g.offGS[l], 0, dst, l == 1 ? d.dtype : 0, ws + g.offGS[l - 1], WF(0), C, planes,
                                                                  g.H[l - 1], g.W[l - 1], g.H[l], g.W[l]);
        }
    }
    g_wgrad_finalize<<<g_blocks((long)(L + 2) * C * (KK + 1), sms), 256, 0, st>>>(part, gw, gb, L, d.B, C, KK);
    return cudaGetLastError();
}


// ---- RecAttn2d pieces for activations the tensor-core kernels do not take (fp32: the 1e-5 bar).  No workspace: the filter is read
// from the parameter tensor, and the tail evaluates s = x + interpolate(z) at each of the K*K taps of an output on the fly
// (correctness and coverage, not speed: 4 K*K loads of z per output in bilinear mode).
//   variant 1: out = conv_s2(x) + b                          model/recattn.py:60 (`down[0]`, BatchNorm folded by the caller)
//   variant 2: out = conv_s1(x + interpolate(z)) + b         model/recattn.py:67
template <int K>
__global__ void g_recattn(int variant, const void* __restrict__ x, const void* __restrict__ z, void* __restrict__ out, int dtype, const void* __restrict__ w,
                          const void* __restrict__ b, int wdtype, int C, long planes, int H, int W, int zH, int zW, int Ho, int Wo, int mode) {
    constexpr int PAD = K / 2, KK = K * K;
    const int S = variant == 1 ? 2 : 1;
    const long total = planes * Ho * Wo;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % Wo);
        const long t = idx / Wo;
        const int i = (int)(t % Ho);
        const long p = t / Ho;
        const int c = (int)(p % C);
        const long xb = p * (long)H * W, zb = p * (long)zH * zW;
        float acc = b ? rc_load_param(b, wdtype, c) : 0.f;
        for (int r = 0; r < K; ++r) {
            const int yy = i * S + r - PAD;
            if (yy < 0 || yy >= H) continue;
            for (int q = 0; q < K; ++q) {
                const int xx = j * S + q - PAD;
                if (xx < 0 || xx >= W) continue;
                float v = g_load(x, dtype, xb + (long)yy * W + xx);
                if (variant == 2) {
                    if (mode == 1) {
                        v += g_load(z, dtype, zb + (long)rc_nearest_src(zH, H, yy) * zW + rc_nearest_src(zW, W, xx));
                    } else {
                        int y0, y1, x0, x1; float ly, lx;
                        rc_bilinear_src(zH, H, yy, y0, y1, ly);
                        rc_bilinear_src(zW, W, xx, x0, x1, lx);
                        const float hy = 1.f - ly, hx = 1.f - lx;
                        v += hy * (hx * g_load(z, dtype, zb + (long)y0 * zW + x0) + lx * g_load(z, dtype, zb + (long)y0 * zW + x1)) +
                             ly * (hx * g_load(z, dtype, zb + (long)y1 * zW + x0) + lx * g_load(z, dtype, zb + (long)y1 * zW + x1));
                    }
                }
                acc = fmaf(v, rc_load_param(w, wdtype, (long)c * KK + r * K + q), acc);
            }
        }
        g_store(out, dtype, idx, acc);
    }
}

}  // namespace

size_t gstream_workspace_bytes(const GStreamDesc& d, bool bwd) { return (size_t)g_layout(d, bwd).total * sizeof(float); }

cudaError_t gstream_launch(const GStreamDesc& d, const KernelArgs& a, void* ws, bool bwd, float* gw, float* gb, cudaStream_t st) {
    float* w = reinterpret_cast<float*>(ws);
    switch (d.K) {
        case 3: return g_run<3>(d, a, w, bwd, gw, gb, st);
        case 5: return g_run<5>(d, a, w, bwd, gw, gb, st);
        case 7: return g_run<7>(d, a, w, bwd, gw, gb, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t gstream_recattn(const GStreamDesc& d, int variant, const void* w, const void* b, const void* x, const void* z, int zH, int zW, void* out,
                            cudaStream_t st) {
    const long planes = (long)d.B * d.C;
    const int Ho = variant == 1 ? rc_down_size(d.H, d.K) : d.H, Wo = variant == 1 ? rc_down_size(d.W, d.K) : d.W;
    const unsigned blocks = g_blocks(planes * Ho * Wo, d.num_sms);
    switch (d.K) {
        case 3: g_recattn<3><<<blocks, 256, 0, st>>>(variant, x, z, out, d.dtype, w, b, d.wdtype, d.C, planes, d.H, d.W, zH, zW, Ho, Wo, d.mode); break;
        case 5: g_recattn<5><<<blocks, 256, 0, st>>>(variant, x, z, out, d.dtype, w, b, d.wdtype, d.C, planes, d.H, d.W, zH, zW, Ho, Wo, d.mode); break;
        case 7: g_recattn<7><<<blocks, 256, 0, st>>>(variant, x, z, out, d.dtype, w, b, d.wdtype, d.C, planes, d.H, d.W, zH, zW, Ho, Wo, d.mode); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace recnext
