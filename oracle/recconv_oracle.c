/*
 * oracle/recconv_oracle.c — CPU restatement of RecNeXt's RecConv2d, forward and backward.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.  The product path
 * (recnext_b200/) never links or calls it and has no CPU fallback.
 *
 * Parity pinning: the reference ships NO golden vectors for this path.
 * This restatement is pinned instead against outputs of the UNMODIFIED reference module
 * (/root/reference/model/recnext.py, imported in the build container through
 * oracle/timm_shim) — see oracle/gen_golden.py, tests/golden/ and tests/test_oracle.py.
 *
 * What it follows (reference file:line, relative to /root/reference):
 *   model/recnext.py:9-22   parameters: `down` = depthwise k×k stride 2 pad k/2 (ONE filter shared by
 *                           all levels), `convs[0..L]` = depthwise k×k stride 1 pad k/2, optional bias
 *   model/recnext.py:27-29  down pass   x_l = down(x_{l-1}),  l = 1..L, remembering size s_{l-1}
 *   model/recnext.py:31-33  up pass     u_{l-1} = interpolate(convs[L-l](x_l + u_l), size = s_{l-1}),  u_L = 0
 *   model/recnext.py:34     final       y = convs[L](x_0 + u_0)
 * The arithmetic itself lives in PyTorch ATen (third party, torch 2.11.0 in this image); the index
 * math restated here follows the installed headers
 *   ATen/native/UpSample.h:259-311  (compute_scales_value, area_pixel_compute_source_index)
 *   ATen/native/UpSample.h:313-358  (nearest_neighbor_compute_source_index, nearest_idx)
 *   ATen/native/UpSample.h:441-476  (guard_index_and_lambda, compute_source_index_and_lambda)
 * i.e. scale = (float)in / out, src = scale*(dst+0.5)-0.5 clamped at 0, all in fp32.
 *
 * Numerics: tensors are fp32 at every op boundary, exactly like the eager reference; each
 * convolution tap-sum is accumulated in double and rounded once (the reference accumulates in
 * fp32 in oneDNN/cuDNN order, so it sits within ~1e-6 of this).  With round_bf16 != 0 every
 * intermediate tensor (and the weights) is rounded to bfloat16 (RNE), which is what the eager
 * reference does under bf16 autocast.
 *
 * Layout: NCHW contiguous, every (n, c) plane independent (all convs are groups = C).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_LEVEL 8

/* ---- bf16 rounding (round-to-nearest-even), used only when emulating autocast ---- */
static float round_bf16(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return v; /* inf / nan unchanged */
    u += 0x7fffu + ((u >> 16) & 1u);
    u &= 0xffff0000u;
    memcpy(&v, &u, 4);
    return v;
}
static void round_buf(float* p, long n, int on) {
    if (!on) return;
    for (long i = 0; i < n; ++i) p[i] = round_bf16(p[i]);
}

/* ---- interpolation index math, fp32 exactly as ATen (UpSample.h:259-311, 441-476) ---- */
void recconv_oracle_bilinear_index(int in_size, int out_size, int dst, int* i0, int* i1, float* lambda1) {
    if (out_size == in_size) { *i0 = dst; *i1 = dst; *lambda1 = 0.0f; return; }
    const float scale = (float)in_size / (float)out_size;
    /* torch 2.11's ATen evaluates scale*(dst+0.5)-0.5 as ONE fused multiply-add, on CPU (probe:
     * 129->257, dst 128 gives i0=63, lambda=0.999996) and on CUDA (nvcc contracts it), so fmaf it is. */
    float src = fmaf(scale, (float)dst + 0.5f, -0.5f);
    if (src < 0.0f) src = 0.0f;
    int idx = (int)floorf(src);
    if (idx > in_size - 1) idx = in_size - 1;
    float lam = src - (float)idx;
    if (lam < 0.0f) lam = 0.0f;
    if (lam > 1.0f) lam = 1.0f;
    *i0 = idx;
    *i1 = idx + ((idx < in_size - 1) ? 1 : 0);
    *lambda1 = lam;
}

/* UpSample.h:313-358 — legacy "nearest": floor(dst * scale), special cases for 1x and 2x */
int recconv_oracle_nearest_index(int in_size, int out_size, int dst) {
    if (out_size == in_size) return dst;
    if (out_size == 2 * in_size) return dst >> 1;
    const float scale = (float)in_size / (float)out_size;
    int s = (int)floorf((float)dst * scale);
    return s < in_size - 1 ? s : in_size - 1;
}

/* ---- single-plane primitives ---- */

/* depthwise cross-correlation (what nn.Conv2d computes), pad k/2; out size given by caller */
static void dwconv_plane(const float* in, int Hi, int Wi, const float* w, float bias, int k, int stride,
                         float* out, int Ho, int Wo) {
    const int p = k / 2;
    for (int oy = 0; oy < Ho; ++oy)
        for (int ox = 0; ox < Wo; ++ox) {
            double acc = bias;
            for (int r = 0; r < k; ++r) {
                const int iy = oy * stride + r - p;
                if (iy < 0 || iy >= Hi) continue;
                for (int s = 0; s < k; ++s) {
                    const int ix = ox * stride + s - p;
                    if (ix < 0 || ix >= Wi) continue;
                    acc += (double)in[iy * Wi + ix] * (double)w[r * k + s];
                }
            }
            out[oy * Wo + ox] = (float)acc;
        }
}

/* transpose of dwconv_plane w.r.t. its input: gin[iy,ix] = sum over (oy,ox,r,s) hitting it */
static void dwconv_plane_bwd_input(const float* gout, int Ho, int Wo, const float* w, int k, int stride,
                                   float* gin, int Hi, int Wi) {
    const int p = k / 2;
    for (int iy = 0; iy < Hi; ++iy)
        for (int ix = 0; ix < Wi; ++ix) {
            double acc = 0.0;
            for (int r = 0; r < k; ++r) {
                const int ty = iy + p - r;
                if (ty < 0 || ty % stride) continue;
                const int oy = ty / stride;
                if (oy >= Ho) continue;
                for (int s = 0; s < k; ++s) {
                    const int tx = ix + p - s;
                    if (tx < 0 || tx % stride) continue;
                    const int ox = tx / stride;
                    if (ox >= Wo) continue;
                    acc += (double)gout[oy * Wo + ox] * (double)w[r * k + s];
                }
            }
            gin[iy * Wi + ix] = (float)acc;
        }
}

/* weight gradient, ACCUMULATED into gw (double), bias gradient accumulated into *gb */
static void dwconv_plane_bwd_weight(const float* in, int Hi, int Wi, const float* gout, int Ho, int Wo, int k,
                                    int stride, double* gw, double* gb) {
    const int p = k / 2;
    for (int r = 0; r < k; ++r)
        for (int s = 0; s < k; ++s) {
            double acc = 0.0;
            for (int oy = 0; oy < Ho; ++oy) {
                const int iy = oy * stride + r - p;
                if (iy < 0 || iy >= Hi) continue;
                for (int ox = 0; ox < Wo; ++ox) {
                    const int ix = ox * stride + s - p;
                    if (ix < 0 || ix >= Wi) continue;
                    acc += (double)in[iy * Wi + ix] * (double)gout[oy * Wo + ox];
                }
            }
            gw[r * k + s] += acc;
        }
    double sb = 0.0;
    for (long i = 0; i < (long)Ho * Wo; ++i) sb += gout[i];
    *gb += sb;
}

static void upsample_plane(const float* in, int Hi, int Wi, float* out, int Ho, int Wo, int mode) {
    if (mode == 1) { /* nearest */
        for (int oy = 0; oy < Ho; ++oy) {
            const int sy = recconv_oracle_nearest_index(Hi, Ho, oy);
            for (int ox = 0; ox < Wo; ++ox)
                out[oy * Wo + ox] = in[sy * Wi + recconv_oracle_nearest_index(Wi, Wo, ox)];
        }
        return;
    }
    for (int oy = 0; oy < Ho; ++oy) {
        int y0, y1; float ly;
        recconv_oracle_bilinear_index(Hi, Ho, oy, &y0, &y1, &ly);
        const float hy = 1.0f - ly;
        for (int ox = 0; ox < Wo; ++ox) {
            int x0, x1; float lx;
            recconv_oracle_bilinear_index(Wi, Wo, ox, &x0, &x1, &lx);
            const float hx = 1.0f - lx;
            out[oy * Wo + ox] = hy * (hx * in[y0 * Wi + x0] + lx * in[y0 * Wi + x1]) +
                                ly * (hx * in[y1 * Wi + x0] + lx * in[y1 * Wi + x1]);
        }
    }
}

/* transpose of upsample_plane: gin (Hi×Wi) = Upᵀ gout (Ho×Wo) */
static void upsample_plane_bwd(const float* gout, int Ho, int Wo, float* gin, int Hi, int Wi, int mode) {
    double* acc = (double*)calloc((size_t)Hi * Wi, sizeof(double));
    for (int oy = 0; oy < Ho; ++oy)
        for (int ox = 0; ox < Wo; ++ox) {
            const double g = gout[oy * Wo + ox];
            if (mode == 1) {
                acc[recconv_oracle_nearest_index(Hi, Ho, oy) * Wi + recconv_oracle_nearest_index(Wi, Wo, ox)] += g;
            } else {
                int y0, y1, x0, x1; float ly, lx;
                recconv_oracle_bilinear_index(Hi, Ho, oy, &y0, &y1, &ly);
                recconv_oracle_bilinear_index(Wi, Wo, ox, &x0, &x1, &lx);
                const float hy = 1.0f - ly, hx = 1.0f - lx;
                acc[y0 * Wi + x0] += (double)(hy * hx) * g;
                acc[y0 * Wi + x1] += (double)(hy * lx) * g;
                acc[y1 * Wi + x0] += (double)(ly * hx) * g;
                acc[y1 * Wi + x1] += (double)(ly * lx) * g;
            }
        }
    for (long i = 0; i < (long)Hi * Wi; ++i) gin[i] = (float)acc[i];
    free(acc);
}

/* ---- pyramid geometry: H_l = floor((H_{l-1} + 2*(k/2) - k)/2) + 1 (nn.Conv2d stride 2) ---- */
static int down_size(int n, int k) { return (n + 2 * (k / 2) - k) / 2 + 1; }

typedef struct {
    int L, k, mode, rb;
    int Hs[ORACLE_MAX_LEVEL + 1], Ws[ORACLE_MAX_LEVEL + 1];
    float* x[ORACLE_MAX_LEVEL + 1]; /* x_l  (x[0] is NOT owned) */
    float* s[ORACLE_MAX_LEVEL + 1]; /* s_l = x_l + u_l (conv inputs of the up pass) */
    float* t;                       /* scratch: conv output before interpolation */
    float* u;                       /* scratch: interpolated tensor */
} plane_ws;

static void ws_alloc(plane_ws* ws, int H, int W, int k, int L, int mode, int rb) {
    ws->L = L; ws->k = k; ws->mode = mode; ws->rb = rb;
    ws->Hs[0] = H; ws->Ws[0] = W;
    for (int l = 1; l <= L; ++l) { ws->Hs[l] = down_size(ws->Hs[l - 1], k); ws->Ws[l] = down_size(ws->Ws[l - 1], k); }
    for (int l = 0; l <= L; ++l) {
        ws->x[l] = l ? (float*)malloc(sizeof(float) * ws->Hs[l] * ws->Ws[l]) : NULL;
        ws->s[l] = (float*)malloc(sizeof(float) * ws->Hs[l] * ws->Ws[l]);
    }
    ws->t = (float*)malloc(sizeof(float) * H * W);
    ws->u = (float*)malloc(sizeof(float) * H * W);
}
static void ws_free(plane_ws* ws) {
    for (int l = 0; l <= ws->L; ++l) { if (l) free(ws->x[l]); free(ws->s[l]); }
    free(ws->t); free(ws->u);
}

/* forward for one plane; fills ws->x[1..L], ws->s[0..L]; y may be NULL (backward recompute) */
static void plane_forward(plane_ws* ws, const float* x0, const float* wd, float bd, const float* const* wc,
                          const float* bc, float* y) {
    const int L = ws->L, k = ws->k;
    ws->x[0] = (float*)x0;
    for (int l = 1; l <= L; ++l) { /* model/recnext.py:27-29 */
        dwconv_plane(ws->x[l - 1], ws->Hs[l - 1], ws->Ws[l - 1], wd, bd, k, 2, ws->x[l], ws->Hs[l], ws->Ws[l]);
        round_buf(ws->x[l], (long)ws->Hs[l] * ws->Ws[l], ws->rb);
    }
    /* model/recnext.py:31-33 ; convs[0] acts on the deepest level */
    memcpy(ws->s[L], ws->x[L], sizeof(float) * ws->Hs[L] * ws->Ws[L]); /* f + 0 */
    for (int l = L; l >= 1; --l) {
        const long nl = (long)ws->Hs[l] * ws->Ws[l], nu = (long)ws->Hs[l - 1] * ws->Ws[l - 1];
        dwconv_plane(ws->s[l], ws->Hs[l], ws->Ws[l], wc[L - l], bc ? bc[L - l] : 0.0f, k, 1, ws->t, ws->Hs[l], ws->Ws[l]);
        round_buf(ws->t, nl, ws->rb);
        upsample_plane(ws->t, ws->Hs[l], ws->Ws[l], ws->u, ws->Hs[l - 1], ws->Ws[l - 1], ws->mode);
        round_buf(ws->u, nu, ws->rb);
        for (long i = 0; i < nu; ++i) ws->s[l - 1][i] = ws->x[l - 1][i] + ws->u[i];
        round_buf(ws->s[l - 1], nu, ws->rb);
    }
    if (y) { /* model/recnext.py:34 */
        dwconv_plane(ws->s[0], ws->Hs[0], ws->Ws[0], wc[L], bc ? bc[L] : 0.0f, k, 1, y, ws->Hs[0], ws->Ws[0]);
        round_buf(y, (long)ws->Hs[0] * ws->Ws[0], ws->rb);
    }
}

static int check_args(int B, int C, int H, int W, int k, int level, int mode) {
    if (B < 0 || C < 0 || H < 1 || W < 1) return 1;
    if (k < 1 || !(k & 1)) return 2;
    if (level < 0 || level > ORACLE_MAX_LEVEL) return 3;
    if (mode != 0 && mode != 1) return 4;
    return 0;
}

/*
 * Forward.  w_down [C,k,k], b_down [C] or NULL, w_convs [(L+1),C,k,k] (convs[j] at j*C*k*k),
 * b_convs [(L+1),C] or NULL.  mode 0 = bilinear, 1 = nearest.
 */
int recconv_oracle_forward(const float* x, const float* w_down, const float* b_down, const float* w_convs,
                           const float* b_convs, float* y, int B, int C, int H, int W, int k, int level, int mode,
                           int round_bf16_flag) {
    int rc = check_args(B, C, H, W, k, level, mode);
    if (rc) return rc;
    const int kk = k * k, L = level;
    plane_ws ws;
    ws_alloc(&ws, H, W, k, L, mode, round_bf16_flag);
    float* wbuf = (float*)malloc(sizeof(float) * kk * (L + 2));
    for (int n = 0; n < B; ++n)
        for (int c = 0; c < C; ++c) {
            const float* wc[ORACLE_MAX_LEVEL + 1];
            float bc[ORACLE_MAX_LEVEL + 1];
            memcpy(wbuf, w_down + (long)c * kk, sizeof(float) * kk);
            for (int j = 0; j <= L; ++j) {
                memcpy(wbuf + (j + 1) * kk, w_convs + ((long)j * C + c) * kk, sizeof(float) * kk);
                wc[j] = wbuf + (j + 1) * kk;
                bc[j] = b_convs ? b_convs[(long)j * C + c] : 0.0f;
                if (round_bf16_flag) bc[j] = round_bf16(bc[j]);
            }
            round_buf(wbuf, (long)kk * (L + 2), round_bf16_flag);
            float bd = b_down ? b_down[c] : 0.0f;
            if (round_bf16_flag) bd = round_bf16(bd);
            const long off = ((long)n * C + c) * H * W;
            plane_forward(&ws, x + off, wbuf, bd, wc, b_convs ? bc : NULL, y + off);
        }
    free(wbuf);
    ws_free(&ws);
    return 0;
}

/*
 * Backward (autograd of the forward above).  Outputs: gx [B,C,H,W];
 * gw_down [C,k,k] (summed over ALL levels — the filter is shared), gb_down [C] or NULL,
 * gw_convs [(L+1),C,k,k], gb_convs [(L+1),C] or NULL.  Weight/bias grads are summed over the batch.
 */
int recconv_oracle_backward(const float* x, const float* gy, const float* w_down, const float* b_down,
                            const float* w_convs, const float* b_convs, float* gx, float* gw_down, float* gb_down,
                            float* gw_convs, float* gb_convs, int B, int C, int H, int W, int k, int level, int mode) {
    int rc = check_args(B, C, H, W, k, level, mode);
    if (rc) return rc;
    const int kk = k * k, L = level;
    plane_ws ws;
    ws_alloc(&ws, H, W, k, L, mode, 0);
    float* G[ORACLE_MAX_LEVEL + 1]; /* G[l] = gradient reaching x_l through the up pass (= grad of s_l) */
    for (int l = 0; l <= L; ++l) G[l] = (float*)malloc(sizeof(float) * ws.Hs[l] * ws.Ws[l]);
    float* gt = (float*)malloc(sizeof(float) * H * W);
    float* tmp = (float*)malloc(sizeof(float) * H * W);
    double* aw = (double*)calloc((size_t)C * kk * (L + 2), sizeof(double)); /* [0]=down, [1+j]=convs[j] */
    double* ab = (double*)calloc((size_t)C * (L + 2), sizeof(double));
    for (int n = 0; n < B; ++n)
        for (int c = 0; c < C; ++c) {
            const float* wc[ORACLE_MAX_LEVEL + 1];
            float bc[ORACLE_MAX_LEVEL + 1];
            for (int j = 0; j <= L; ++j) {
                wc[j] = w_convs + ((long)j * C + c) * kk;
                bc[j] = b_convs ? b_convs[(long)j * C + c] : 0.0f;
            }
            const float* wd = w_down + (long)c * kk;
            const long off = ((long)n * C + c) * H * W;
            plane_forward(&ws, x + off, wd, b_down ? b_down[c] : 0.0f, wc, b_convs ? bc : NULL, NULL);

            /* y = convs[L](s_0) */
            dwconv_plane_bwd_weight(ws.s[0], H, W, gy + off, H, W, k, 1, aw + ((long)(1 + L) * C + c) * kk,
                                    ab + (long)(1 + L) * C + c);
            dwconv_plane_bwd_input(gy + off, H, W, wc[L], k, 1, G[0], H, W);
            /* s_{l-1} = x_{l-1} + Up(convs[L-l](s_l)) */
            for (int l = 1; l <= L; ++l) {
                upsample_plane_bwd(G[l - 1], ws.Hs[l - 1], ws.Ws[l - 1], gt, ws.Hs[l], ws.Ws[l], mode);
                dwconv_plane_bwd_weight(ws.s[l], ws.Hs[l], ws.Ws[l], gt, ws.Hs[l], ws.Ws[l], k, 1,
                                        aw + ((long)(1 + L - l) * C + c) * kk, ab + (long)(1 + L - l) * C + c);
                dwconv_plane_bwd_input(gt, ws.Hs[l], ws.Ws[l], wc[L - l], k, 1, G[l], ws.Hs[l], ws.Ws[l]);
            }
            /* x_l = down(x_{l-1}): total grad T_l = G[l] + downᵀ(T_{l+1}); accumulate in place in G */
            for (int l = L; l >= 1; --l) {
                dwconv_plane_bwd_weight(ws.x[l - 1], ws.Hs[l - 1], ws.Ws[l - 1], G[l], ws.Hs[l], ws.Ws[l], k, 2,
                                        aw + (long)c * kk, ab + c);
                dwconv_plane_bwd_input(G[l], ws.Hs[l], ws.Ws[l], wd, k, 2, tmp, ws.Hs[l - 1], ws.Ws[l - 1]);
                for (long i = 0; i < (long)ws.Hs[l - 1] * ws.Ws[l - 1]; ++i) G[l - 1][i] += tmp[i];
            }
            memcpy(gx + off, G[0], sizeof(float) * H * W);
        }
    for (long i = 0; i < (long)C * kk; ++i) gw_down[i] = (float)aw[i];
    for (long i = 0; i < (long)(L + 1) * C * kk; ++i) gw_convs[i] = (float)aw[(long)C * kk + i];
    if (gb_down) for (int c = 0; c < C; ++c) gb_down[c] = (float)ab[c];
    if (gb_convs) for (long i = 0; i < (long)(L + 1) * C; ++i) gb_convs[i] = (float)ab[C + i];
    for (int l = 0; l <= L; ++l) free(G[l]);
    free(gt); free(tmp); free(aw); free(ab);
    ws_free(&ws);
    return 0;
}

/* pyramid sizes, for tests: fills Hs/Ws[0..level] */
int recconv_oracle_pyramid(int H, int W, int k, int level, int* Hs, int* Ws) {
    if (level < 0 || level > ORACLE_MAX_LEVEL) return 3;
    Hs[0] = H; Ws[0] = W;
    for (int l = 1; l <= level; ++l) { Hs[l] = down_size(Hs[l - 1], k); Ws[l] = down_size(Ws[l - 1], k); }
    return 0;
}
