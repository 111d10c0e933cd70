// recconv_plan.h — launch plan / shared-memory layout of the fused RecConv kernels.
//
// One CTA owns P consecutive (n, c) planes of ONE image ("plane group" = P consecutive channels, which are
// contiguous in NCHW) and keeps the whole pyramid of those planes in shared memory; it walks over a chunk of
// images for its channel group, so filters are loaded once and weight-gradient partials stay on chip.
// Each plane is worked on by `g` lanes.  All geometry is decided on the host and passed by value.
//
// Padded level buffers: level l is stored as rows_l x pitch_l floats, interior at (row + pad, col + pad),
// pad = k/2, everything outside the interior is zero for the whole kernel (zeroed once), which is what makes
// the stencil loops free of bounds checks (nn.Conv2d zero padding, reference model/recnext.py:18).
#pragma once
#include <stdint.h>

#ifndef RECNEXT_MAX_LEVEL
#define RECNEXT_MAX_LEVEL 6
#endif

#if defined(__CUDACC__)
#define RC_HD __host__ __device__ __forceinline__
#define RC_H __host__ inline
#else
#define RC_HD inline
#define RC_H inline
#endif

namespace recnext {

constexpr int kMaxLevel = RECNEXT_MAX_LEVEL;
constexpr int kStripW = 4;  // output columns per work item (one float4 of accumulators)

struct LevelGeo {
    int H, W;        // level size (level 0 = input)
    int pitch;       // floats per padded row
    int rows;        // padded rows
    int offS;        // float offset of S_l (x_l, then x_l + u_l) inside the plane block
    int offX;        // bwd: copy of x_l (levels 1..L-1), -1 if absent
    int offGT;       // bwd: gradient w.r.t. t_l = convs[L-l](s_l) after upsample-backward (padded), levels 1..L
    int offGS;       // bwd: gradient w.r.t. s_l, later total gradient of x_l (padded), levels 1..L
    int rpi;         // rows per work item for stride-1 stencils ON this level
    int rpi_down;    // rows per work item for the stride-2 stencil PRODUCING this level (l >= 1)
    int tabY, tabX;  // byte offsets (in the table region) of the forward interpolation tables that map
                     // level l-1 coordinates to level l sources: {int i0; float lambda}[H_{l-1}] / [W_{l-1}]
    int rngY, rngX;  // bwd: byte offsets of {int lo; int hi}[H_l] / [W_l]: destinations touching source i
};

struct Plan {
    int B, C, H, W, K, L, mode, dtype, wdtype, has_bias, backward;
    int P;             // planes (channels) per CTA
    int g;             // lanes per plane (power of two <= 32, or a multiple of 32)
    int T;             // threads per CTA = P * g
    int n_cg;          // channel groups = ceil(C / P)
    int n_chunk;       // image chunks; grid = n_cg * n_chunk
    int img_per_chunk;
    int use_tma;       // 1: cp.async.bulk loads/stores of whole plane groups (needs 16-byte alignment)
    int share_raw;     // 1: the raw output buffer aliases a raw input buffer (big planes): no prefetch of the next image
    int esize;         // bytes per element of x
    LevelGeo lv[kMaxLevel + 1];
    int offT;          // float offset of T (unpadded conv output awaiting interpolation), fwd and bwd
    int offGY;         // bwd: padded gy (level-0 geometry)
    int offG0;         // bwd: gradient w.r.t. s_0 (unpadded, pitch = pitchG0)
    int pitchG0;
    int plane_floats;  // floats per plane block
    // byte offsets inside dynamic shared memory
    int smTab, smW, smWG, smRawX, smRawG, smRawOut, smPlanes, smem_bytes;
    int wstride;       // floats per (plane, conv) filter slot = K*K + 1 (bias last)
    int nslots;        // bwd: weight-gradient accumulation slots per CTA
    int raw_plane_bytes;
    int ws_partial_floats;  // bwd: floats of per-chunk partials in the workspace
};

RC_HD int rc_down_size(int n, int k) { return (n + 2 * (k / 2) - k) / 2 + 1; }
RC_HD int rc_round_up(int v, int m) { return (v + m - 1) / m * m; }
RC_HD int rc_div_up(int a, int b) { return (a + b - 1) / b; }

// window floats loaded per input row by a 4-column strip
RC_HD int rc_win_s1(int k) { return rc_round_up(kStripW + 2 * (k / 2), 4); }
RC_HD int rc_win_s2(int k) { return rc_round_up(2 * (kStripW - 1) + k, 4); }

// rows-per-item so that the `g` lanes of a plane are busy and few rounds are needed
RC_H int rc_pick_rpi(int Ho, int Wo, int g, int halo_rows) {
    const int nstrips = rc_div_up(Wo, kStripW);
    int best = 1;
    long best_cost = -1;
    for (int rpi = 1; rpi <= Ho; ++rpi) {
        const int items = nstrips * rc_div_up(Ho, rpi);
        const int rounds = rc_div_up(items, g);
        const long cost = (long)rounds * (rpi * 8 + halo_rows * 2 + 6);  // ~instruction slots per lane
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = rpi; }
    }
    return best;
}

struct PlanOptions {
    int force_P = 0, force_g = 0, force_chunks = 0, force_no_tma = 0;
    int num_sms = 148;
    int smem_limit = 227 * 1024;
};

// Returns 0 on success, 1 if the plane pyramid does not fit in shared memory, 2 on bad arguments.
RC_H int rc_make_plan(Plan& pl, int B, int C, int H, int W, int K, int L, int mode, int dtype, int wdtype, int has_bias,
                      int backward, const PlanOptions& opt) {
    if (B < 1 || C < 1 || H < 1 || W < 1 || !(K == 3 || K == 5 || K == 7) || L < 0 || L > kMaxLevel) return 2;
    pl = Plan();
    pl.B = B; pl.C = C; pl.H = H; pl.W = W; pl.K = K; pl.L = L; pl.mode = mode; pl.dtype = dtype; pl.wdtype = wdtype;
    pl.has_bias = has_bias; pl.backward = backward;
    pl.esize = dtype == 0 ? 4 : 2;
    const int pad = K / 2;
    pl.lv[0].H = H; pl.lv[0].W = W;
    for (int l = 1; l <= L; ++l) { pl.lv[l].H = rc_down_size(pl.lv[l - 1].H, K); pl.lv[l].W = rc_down_size(pl.lv[l - 1].W, K); }
    int off = 0;
    for (int l = 0; l <= L; ++l) {
        LevelGeo& g = pl.lv[l];
        int need = rc_round_up(g.W, kStripW) - kStripW + rc_win_s1(K);  // stride-1 readers of this level
        if (l < L) {  // the stride-2 stencil producing level l+1 reads this level
            const int n2 = 2 * (rc_round_up(pl.lv[l + 1].W, kStripW) - kStripW) + rc_win_s2(K);
            if (n2 > need) need = n2;
        }
        if (need < g.W + 2 * pad) need = g.W + 2 * pad;
        g.pitch = rc_round_up(need, 4);
        if ((g.pitch & 31) == 0) g.pitch += 4;  // keep row-to-row bank offsets non-zero
        g.rows = g.H + 2 * pad;
        if (l < L) {  // stride-2 reader touches rows up to 2*(H_{l+1}-1) + K - 1
            const int r2 = 2 * (pl.lv[l + 1].H - 1) + K;
            if (r2 > g.rows) g.rows = r2;
        }
        g.offS = off; off += g.rows * g.pitch;
        g.offX = g.offGT = g.offGS = -1;
    }
    const int nT = L >= 1 ? rc_round_up(pl.lv[1].H * pl.lv[1].W, 4) : 0;
    pl.offT = off; off += nT;
    if (backward) {
        for (int l = 1; l <= L; ++l) {
            LevelGeo& g = pl.lv[l];
            if (l < L) { g.offX = off; off += g.rows * g.pitch; }
            g.offGT = off; off += g.rows * g.pitch;
            g.offGS = off; off += g.rows * g.pitch;
        }
        pl.offGY = off; off += pl.lv[0].rows * pl.lv[0].pitch;
        pl.pitchG0 = rc_round_up(W, 4);
        pl.offG0 = off; off += rc_round_up(H * pl.pitchG0, 4);
    }
    pl.plane_floats = rc_round_up(off, 4);
    pl.wstride = K * K + 1;
    pl.raw_plane_bytes = H * W * pl.esize;

    // table region (shared by all planes of the CTA)
    int tb = 0;
    for (int l = 1; l <= L; ++l) {
        pl.lv[l].tabY = tb; tb += 8 * pl.lv[l - 1].H;
        pl.lv[l].tabX = tb; tb += 8 * pl.lv[l - 1].W;
        pl.lv[l].rngY = tb; tb += 8 * pl.lv[l].H;
        pl.lv[l].rngX = tb; tb += 8 * pl.lv[l].W;
    }
    tb = rc_round_up(tb, 16);

    // ---- choose lanes per plane g and planes per CTA P ----
    const int items0 = rc_div_up(W, kStripW) * H;  // finest split of level 0 (one row per item)
    int g = 1;
    while (g < 32 && g * 2 * 4 <= items0) g *= 2;            // aim at >= ~4 rows per lane ...
    if (g == 32) { while (g < 256 && (g + 32) * 7 <= items0) g += 32; }  // ... and ~7+ rows per lane for big planes
    if (g > 32) { int w = g / 32; while (w & (w - 1)) --w; g = w * 32; }  // whole power-of-two warps
    if (opt.force_g) g = opt.force_g;
    const long per_plane_bytes = (long)pl.plane_floats * 4 + (long)pl.raw_plane_bytes * (backward ? 3 : 2) +
                                 (long)(L + 2) * pl.wstride * 4;
    auto smem_for = [&](int P, int T) -> long {
        const int nslots = backward ? (g >= 32 ? T / 32 : P) : 0;
        return 64 + tb + per_plane_bytes * P + (long)nslots * (L + 2) * pl.wstride * 4 + 3 * 128 + 256;
    };
    int P = 0;
    if (opt.force_P) {
        P = opt.force_P;
    } else {
        // target ~128..256 threads per CTA, several CTAs per SM, and 16-byte aligned plane groups for TMA
        int Pmax = C;
        const int t_target = g >= 128 ? g : 128;
        int Pt = t_target / g; if (Pt < 1) Pt = 1;
        if (Pt > Pmax) Pt = Pmax;
        P = Pt;
        // shrink until at least 2 CTAs fit per SM (if possible at all)
        while (P > 1 && smem_for(P, P * g) * 2 > opt.smem_limit) --P;
        // prefer a P that divides C and keeps groups 16-byte aligned
        for (int cand = P; cand >= 1; --cand) {
            if (C % cand == 0 && ((long)cand * pl.raw_plane_bytes) % 16 == 0) { if (cand * 2 > P) P = cand; break; }
        }
    }
    if (P < 1) P = 1;
    if (P > C) P = C;
    pl.P = P; pl.g = g; pl.T = rc_round_up(P * g, 32);  // whole warps (surplus lanes own no plane)
    if (pl.T > 1024) return 1;
    pl.nslots = backward ? (g >= 32 ? pl.T / 32 : P) : 0;
    pl.n_cg = rc_div_up(C, P);
    pl.use_tma = (!opt.force_no_tma && C % P == 0 && ((long)P * pl.raw_plane_bytes) % 16 == 0) ? 1 : 0;

    // rows per item on every level
    for (int l = 0; l <= L; ++l) {
        pl.lv[l].rpi = rc_pick_rpi(pl.lv[l].H, pl.lv[l].W, g, K - 1);
        pl.lv[l].rpi_down = l >= 1 ? rc_pick_rpi(pl.lv[l].H, pl.lv[l].W, g, K - 2) : 0;
    }

    // shared memory map
    int sm = 64;  // two mbarriers + padding
    pl.smTab = sm; sm += tb;
    pl.smW = sm; sm += P * (L + 2) * pl.wstride * 4;
    pl.smWG = sm; sm += pl.nslots * (L + 2) * pl.wstride * 4;
    sm = rc_round_up(sm, 128);
    pl.smRawX = sm; sm += rc_round_up(P * pl.raw_plane_bytes, 128);
    pl.smRawG = sm; if (backward) sm += rc_round_up(P * pl.raw_plane_bytes, 128);
    pl.smRawOut = sm; sm += rc_round_up(P * pl.raw_plane_bytes, 128);
    pl.smPlanes = sm; sm += P * pl.plane_floats * 4;
    if (sm > opt.smem_limit) {  // big planes: write the result over a raw input buffer that is dead by then
        pl.share_raw = 1;
        const int raw = rc_round_up(P * pl.raw_plane_bytes, 128);
        pl.smRawOut = backward ? pl.smRawG : pl.smRawX;
        pl.smPlanes -= raw; sm -= raw;
    }
    pl.smem_bytes = sm;
    if (sm > opt.smem_limit) return 1;

    // grid: enough CTAs for a few waves, but never more chunks than images
    int ctas_per_sm = opt.smem_limit / (sm + 1024);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const int max_by_threads = 2048 / pl.T > 0 ? 2048 / pl.T : 1;
    if (ctas_per_sm > max_by_threads) ctas_per_sm = max_by_threads;
    if (ctas_per_sm > 32) ctas_per_sm = 32;
    const long resident = (long)opt.num_sms * ctas_per_sm;
    int n_chunk = (int)((resident * (backward ? 1 : 2) + pl.n_cg - 1) / pl.n_cg);
    if (n_chunk > B) n_chunk = B;
    if (n_chunk < 1) n_chunk = 1;
    if (opt.force_chunks) n_chunk = opt.force_chunks > B ? B : opt.force_chunks;
    pl.img_per_chunk = rc_div_up(B, n_chunk);
    pl.n_chunk = rc_div_up(B, pl.img_per_chunk);
    pl.ws_partial_floats = backward ? pl.n_chunk * (L + 2) * C * pl.wstride : 0;
    return 0;
}

}  // namespace recnext
