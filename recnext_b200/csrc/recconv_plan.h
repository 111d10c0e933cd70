// recconv_plan.h — launch plan / shared-memory layout of the fused RecConv kernels.
//
// Work decomposition.  Every (n, c) plane of RecConv2d is independent (all convs are depthwise, reference
// model/recnext.py:13-22), so a plane is given to a TEAM of g lanes that keeps the plane's whole pyramid in
// shared memory.  Teams never talk to each other: a team of g <= 32 lanes lives inside one warp and is
// synchronised with __syncwarp(), a bigger team (g = 64..256) owns a named barrier.  A UNIT is what moves
// through TMA together: one warp's teams (32/g consecutive planes) or one multi-warp team (one plane).
// A CTA is P planes = n_units units of ONE image and walks over a chunk of images for its channel group, so
// filters are loaded once and weight-gradient partials stay on chip.
//
// Padded level buffers: level l is stored as rows_l x pitch_l floats, interior at (row + pad, col + pad),
// pad = k/2; everything outside the interior is zero for the whole kernel (zeroed once), which is what makes
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
constexpr int kMaxThreads = 256;

struct StripGrid {  // work items of one stage: 4-column strips x blocks of `rpi` rows, item = rb * nstrips + strip
    int nstrips;
    unsigned magic;  // item / nstrips by multiplication
    int rpi;         // rows per item
    int nitems;
    int lanes;       // leading lanes of the team that take part (power of two >= 32 when the team spans warps)
};

struct LevelGeo {
    int H, W;        // level size (level 0 = input)
    int pitch;       // floats per padded row
    int rows;        // padded rows
    int offS;        // float offset of S_l (x_l, then x_l + u_l) inside the plane block
    int offX;        // bwd: copy of x_l (levels 1..L-1), -1 if absent
    int offGT;       // bwd: gradient w.r.t. t_l = convs[L-l](s_l) (padded), levels 1..L
    int offGS;       // bwd: gradient w.r.t. s_l, later total gradient of x_l (padded), levels 1..L
    StripGrid g1;    // stride-1 stencils ON this level (rows of this level)
    StripGrid g2;    // l >= 1: the stride-2 stencil PRODUCING this level (rows of this level)
    StripGrid gu;    // l >= 1: upsample of this level INTO level l-1: strips of level l-1; rows are SOURCE rows of
                     //         this level on the exact-2x path, destination rows of level l-1 on the generic path
    StripGrid gt;    // l >= 1 (bwd): 2x2 blocks of level l-1 for the transpose of the stride-2 conv (nstrips = blocks per row)
    unsigned magic_W;  // magic for division by W (backward gather)
    int gather_lanes;  // bwd: lanes taking part in the gather of this level
    int tpitch;      // l >= 1: pitch of the T buffer of this level ((H+2) x tpitch, replicate border of 1)
    int exact2x;     // l >= 1: level l-1 is exactly 2x this level in both dimensions (bilinear fast path)
    int tabY, tabX;  // byte offsets (table region) of {int i0; float lambda}[H_{l-1}] / [W_{l-1}] (level l-1 -> l)
    int gatY, gatX;  // bwd: byte offsets of GatherEntry[H_l] / [W_l]: destinations of level l-1 reading source i
};

struct Plan {
    int B, C, H, W, K, L, mode, dtype, wdtype, has_bias, backward;
    int P;             // planes (channels) per CTA
    int g;             // lanes per plane (team size): power of two <= 256
    int T;             // threads per CTA = P * g, a multiple of 32
    int unit_lanes;    // lanes per unit = max(g, 32)
    int ppu;           // planes per unit = 32 / g (g < 32) or 1
    int n_units;       // units per CTA
    int n_cg;          // channel groups = ceil(C / P)
    int n_chunk;       // image chunks; grid = n_cg * n_chunk
    int img_per_chunk;
    int use_tma;       // 1: cp.async.bulk loads/stores of whole units (needs 16-byte alignment)
    int share_raw;     // 1: raw output aliases a raw input buffer (big planes): no prefetch of the next image
    int esize;         // bytes per element of x
    int vec;           // elements per unpack chunk: largest power of two dividing W, <= 16 / esize
    unsigned magic_cpr;  // magic for division by chunks-per-row (W / vec)
    LevelGeo lv[kMaxLevel + 1];
    int offT;          // float offset of T inside the plane block
    int offGY;         // bwd: padded gy (level-0 geometry)
    int offG0;         // bwd: gradient w.r.t. s_0 (unpadded, pitch = pitchG0)
    int pitchG0;
    int plane_floats;  // floats per plane block
    // byte offsets inside dynamic shared memory
    int smBar, smTab, smW, smWG, smRawX, smRawG, smRawOut, smPlanes, smem_bytes;
    int raw_unit_bytes;  // bytes of one unit's raw buffer (padded to 128)
    int wstride;       // floats per (plane, conv) filter slot = round_up(K*K + 1, 4) (bias at [K*K])
    int nslots;        // bwd: weight-gradient accumulation slots per CTA
    int raw_plane_bytes;
    int ws_partial_floats;  // bwd: floats of per-chunk partials in the workspace
};

struct GatherEntry {  // bwd: the (<= 4) destinations d0..d0+n-1 of level l-1 that read source i, with weights
    int d0, n;
    float w[4];
    int pad_[2];
};

RC_HD constexpr int rc_down_size(int n, int k) { return (n + 2 * (k / 2) - k) / 2 + 1; }
RC_HD constexpr int rc_round_up(int v, int m) { return (v + m - 1) / m * m; }
RC_HD constexpr int rc_div_up(int a, int b) { return (a + b - 1) / b; }
// x / n for 0 <= x < 2^16 by multiplication: magic = floor(2^32 / n) + 1 (n >= 2); magic 0 encodes n == 1
RC_HD unsigned rc_magic(int n) { return n <= 1 ? 0u : (unsigned)(0x100000000ull / (unsigned long long)n) + 1u; }
RC_HD int rc_fastdiv(int x, unsigned magic) {
    if (magic == 0u) return x;
#if defined(__CUDA_ARCH__)
    return (int)__umulhi((unsigned)x, magic);
#else
    return (int)(((unsigned long long)(unsigned)x * magic) >> 32);
#endif
}

// window floats loaded per input row by a 4-column strip
RC_HD int rc_win_s1(int k) { return rc_round_up(kStripW + 2 * (k / 2), 4); }
RC_HD int rc_win_s2(int k) { return rc_round_up(2 * (kStripW - 1) + k, 4); }

// rows-per-item so that the `g` lanes of a team are busy and few rounds are needed
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
        if (g.rows < pad + 4) g.rows = pad + 4;  // the backward gather reads 4 rows from the interior origin
        if (l < L) {  // stride-2 reader touches rows up to 2*(H_{l+1}-1) + K - 1
            const int r2 = 2 * (pl.lv[l + 1].H - 1) + K;
            if (r2 > g.rows) g.rows = r2;
        }
        g.offS = off; off += g.rows * g.pitch;
        g.offX = g.offGT = g.offGS = -1;
        if (l >= 1) {
            g.tpitch = rc_round_up(g.W + 2, 4);
            g.exact2x = (pl.lv[l - 1].H == 2 * g.H && pl.lv[l - 1].W == 2 * g.W) ? 1 : 0;
        }
    }
    int nT = 0;
    for (int l = 1; l <= L; ++l) { const int n = (pl.lv[l].H + 2) * pl.lv[l].tpitch; if (n > nT) nT = n; }
    pl.offT = off; off += rc_round_up(nT, 4) + 4;  // + slack: the 2x path may read 3 floats past the last row
    if (backward) {
        for (int l = 1; l <= L; ++l) {
            LevelGeo& g = pl.lv[l];
            if (l < L) { g.offX = off; off += g.rows * g.pitch; }
            g.offGT = off; off += g.rows * g.pitch;
            g.offGS = off; off += g.rows * g.pitch;
        }
        pl.offGY = off; off += pl.lv[0].rows * pl.lv[0].pitch;
        pl.pitchG0 = rc_round_up(W, 4);
        pl.offG0 = off; off += rc_round_up((H < 4 ? 4 : H) * pl.pitchG0, 4) + 4;
    }
    pl.plane_floats = rc_round_up(off, 4);
    pl.wstride = rc_round_up(K * K + 1, 4);
    pl.raw_plane_bytes = H * W * pl.esize;
    int vec = 16 / pl.esize;
    while (vec > 1 && (W % vec) != 0) vec >>= 1;
    pl.vec = vec;
    pl.magic_cpr = rc_magic(W / vec);

    // table region (shared by all planes of the CTA)
    int tb = 0;
    for (int l = 1; l <= L; ++l) {
        pl.lv[l].tabY = tb; tb += 8 * pl.lv[l - 1].H;
        pl.lv[l].tabX = tb; tb += 8 * pl.lv[l - 1].W;
        tb = rc_round_up(tb, 16);
        if (backward) {
            pl.lv[l].gatY = tb; tb += (int)sizeof(GatherEntry) * pl.lv[l].H;
            pl.lv[l].gatX = tb; tb += (int)sizeof(GatherEntry) * pl.lv[l].W;
        }
    }
    tb = rc_round_up(tb, 16);

    // ---- team size g: roughly one lane per 7 rows of a 4-column strip of level 0 ----
    const int strips0 = rc_div_up(W, kStripW);
    int g = 1;
    while (g < kMaxThreads && g * 2 <= strips0 * rc_div_up(H, 7)) g *= 2;
    if (g * 2 <= kMaxThreads && strips0 * rc_div_up(H, 7) > g + g / 2) g *= 2;  // e.g. 56x56: 112 items -> 128 lanes
    if (opt.force_g) g = opt.force_g;
    if (g > kMaxThreads) g = kMaxThreads;
    const int unit_lanes = g < 32 ? 32 : g;
    const int ppu = g < 32 ? 32 / g : 1;
    const int raw_unit = rc_round_up(ppu * pl.raw_plane_bytes, 128);
    const int n_raw = backward ? 3 : 2;
    auto smem_for = [&](int units, bool shared_raw) -> long {
        const int P = units * ppu;
        const int nslots = backward ? (g >= 32 ? units * (g / 32) : P) : 0;
        return 128 + tb + (long)P * (L + 2) * pl.wstride * 4 + (long)nslots * (L + 2) * pl.wstride * 4 + 256 +
               (long)units * raw_unit * (shared_raw ? n_raw - 1 : n_raw) + (long)P * pl.plane_floats * 4;
    };
    int units = kMaxThreads / unit_lanes;
    if (units < 1) units = 1;
    if (opt.force_P) {
        units = rc_div_up(opt.force_P, ppu);
    } else {
        while (units > 1 && units * ppu > rc_round_up(C, ppu)) --units;                   // not more planes than channels
        while (units > 1 && smem_for(units, false) * 2 > opt.smem_limit) --units;         // >= 2 CTAs per SM if possible
        // prefer a plane count that divides C (no ragged channel groups, TMA eligible)
        for (int cand = units; cand >= 1; --cand)
            if (C % (cand * ppu) == 0) { if (cand * 2 > units) units = cand; break; }
    }
    pl.g = g; pl.unit_lanes = unit_lanes; pl.ppu = ppu; pl.n_units = units;
    pl.P = units * ppu;
    pl.T = units * unit_lanes;
    if (pl.T > kMaxThreads) return 1;
    pl.nslots = backward ? (g >= 32 ? units * (g / 32) : pl.P) : 0;
    pl.n_cg = rc_div_up(C, pl.P);
    pl.use_tma = (!opt.force_no_tma && C % pl.P == 0 && ((long)ppu * pl.raw_plane_bytes) % 16 == 0) ? 1 : 0;
    pl.share_raw = smem_for(units, false) > opt.smem_limit ? 1 : 0;
    if (smem_for(units, pl.share_raw != 0) > opt.smem_limit) return 1;

    auto lanes_for = [&](int nitems) {
        if (g <= 32) return 32;  // warp-resident teams: the whole warp takes part in every stage
        int n = 32;
        while (n < g && n < nitems) n *= 2;
        return n;
    };
    auto grid = [&](int rows, int cols, int halo) {
        StripGrid sg;
        sg.nstrips = rc_div_up(cols, kStripW);
        sg.magic = rc_magic(sg.nstrips);
        sg.rpi = rc_pick_rpi(rows, cols, g, halo);
        sg.nitems = sg.nstrips * rc_div_up(rows, sg.rpi);
        sg.lanes = lanes_for(sg.nitems);
        return sg;
    };
    for (int l = 0; l <= L; ++l) {
        LevelGeo& lg = pl.lv[l];
        lg.g1 = grid(lg.H, lg.W, K - 1);
        lg.magic_W = rc_magic(lg.W);
        lg.gather_lanes = lanes_for(lg.H * lg.W);
        if (l >= 1) {
            lg.g2 = grid(lg.H, lg.W, K - 2);
            const LevelGeo& ld = pl.lv[l - 1];
            if (lg.exact2x && mode == 0) lg.gu = grid(lg.H, ld.W, 1);
            else lg.gu = grid(ld.H, ld.W, 0);
            lg.gt.nstrips = (ld.W + 1) / 2;
            lg.gt.magic = rc_magic(lg.gt.nstrips);
            lg.gt.rpi = 1;
            lg.gt.nitems = lg.gt.nstrips * ((ld.H + 1) / 2);
            lg.gt.lanes = lanes_for(lg.gt.nitems);
        }
    }

    // shared memory map
    int sm = 0;
    pl.smBar = sm; sm += 128;  // one mbarrier per unit (<= 8 units)
    pl.smTab = sm; sm += tb;
    pl.smW = sm; sm += pl.P * (L + 2) * pl.wstride * 4;
    pl.smWG = sm; sm += pl.nslots * (L + 2) * pl.wstride * 4;
    sm = rc_round_up(sm, 128);
    pl.raw_unit_bytes = raw_unit;
    pl.smRawX = sm; sm += units * raw_unit;
    pl.smRawG = sm; if (backward) sm += units * raw_unit;
    if (pl.share_raw) pl.smRawOut = backward ? pl.smRawG : pl.smRawX;
    else { pl.smRawOut = sm; sm += units * raw_unit; }
    pl.smPlanes = sm; sm += pl.P * pl.plane_floats * 4;
    pl.smem_bytes = sm;
    if (sm > opt.smem_limit) return 1;

    // grid: enough CTAs for a full wave of resident CTAs, never more chunks than images
    int ctas_per_sm = opt.smem_limit / (sm + 1024);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const int max_by_threads = 2048 / pl.T > 0 ? 2048 / pl.T : 1;
    if (ctas_per_sm > max_by_threads) ctas_per_sm = max_by_threads;
    if (ctas_per_sm > 32) ctas_per_sm = 32;
    const long resident = (long)opt.num_sms * ctas_per_sm;
    int n_chunk = (int)(resident / pl.n_cg);  // each CTA loops over its chunk of images
    if (n_chunk > B) n_chunk = B;
    if (n_chunk < 1) n_chunk = 1;
    if (opt.force_chunks) n_chunk = opt.force_chunks > B ? B : opt.force_chunks;
    pl.img_per_chunk = rc_div_up(B, n_chunk);
    pl.n_chunk = rc_div_up(B, pl.img_per_chunk);
    pl.ws_partial_floats = backward ? pl.n_chunk * (L + 2) * C * pl.wstride : 0;
    return 0;
}

}  // namespace recnext
