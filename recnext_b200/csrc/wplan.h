// wplan.h — launch plan of the TEAM-RESIDENT RecConv kernels (the fast path for planes whose whole pyramid
// fits in a slice of one SM's shared memory; recconv_plan.h keeps the big-plane path).
//
// Why this shape.  On B200 the fused block is bound by the FP32 FMA pipe, not by HBM (k = 5: ~48 FMA per
// element forward, ~105 backward, for 4 bytes of bf16 traffic; tools/fma_probe.cu measures 36 TFMA/s), so the
// schedule is built to keep the FMA pipe issuing:
//   * a TEAM (1, 2, 4 or 8 warps) owns a BATCH of G consecutive (n, c) planes for their whole life: every level
//     of the pyramid stays in the team's private slice of shared memory; teams never talk to each other, a
//     one-warp team synchronises with __syncwarp() only;
//   * every stage is a flat list of ITEMS = (plane, 4-column strip, block of rows) dealt round-robin to the
//     team's lanes; G and the row blocks are chosen so that the item count of every level is close to a
//     multiple of the lane count (small planes are batched: 14x14 -> 8 planes per warp, 7x7 -> 16);
//   * an item runs a rolling K-row register window down its strip: per output row 2 LDS.128 + 4*K*K FFMA;
//   * raw planes arrive by TMA bulk copy (cp.async.bulk, one per batch) into a buffer that aliases the
//     interpolation scratch T, and the next batch is prefetched while the last (largest) conv runs;
//     results go straight from registers to global memory.
#pragma once
#include "recconv_plan.h"

namespace recnext {

struct WGrid {            // items of one stage, per plane: j = rb * spr + strip (strip < strips), j < ipp = nrb * spr.
    int strips, nrb, rpb; //   spr >= strips pads a row block to a multiple of 8 lanes, so that the 8 lanes of a 128-bit
    int spr, ipp, rounds; //   shared-memory wavefront read ONE row (no bank conflicts between row blocks).  Plane g of the
    unsigned m_spr;       //   batch owns the aligned lane group [g * LPP, (g + 1) * LPP) of its team; lane jl of the
};                        //   group takes j = jl + round * LPP, round < rounds = ceil(ipp / LPP)

struct WLevel {
    int H, W, pitch, rows;
    int offS;             // float offset of S_l inside a plane block (padded, interior at (+pad, +pad))
    int offX, offGT, offGS;  // bwd: x_l copy (1..L-1), grad wrt t_l, grad wrt s_l / total grad of x_l (1..L); -1 if absent
    int tp, toff;         // T buffer of this level: floats per row, column offset of the interior (2 + replicate border
                          // on the aligned exact-2x bilinear path, else 0)
    int exact2x;          // level l-1 is exactly 2x this level (bilinear fast path)
    int up_fast;          // exact2x, bilinear, even pad: 128-bit aligned read-modify-write of level l-1
    int tabY, tabX;       // byte offsets (table region) of IdxLam[H_{l-1}] / IdxLam[W_{l-1}]
    int gatY, gatX;       // bwd: GatherEntry[H_l] / GatherEntry[W_l]
    WGrid g1;             // stride-1 stencils on this level
    WGrid g2;             // l >= 1: stride-2 stencil producing this level
    WGrid gu;             // l >= 1: upsample of this level into level l-1 (strips of level l-1; rows: source rows on the
                          //         exact-2x path, destination rows otherwise)
    WGrid gt;             // l >= 1, bwd: 2x2 blocks of level l-1 (transpose of the stride-2 conv)
    unsigned magic_W, magic_HW;
};

struct WPlan {
    int B, C, H, W, K, L, mode, dtype, wdtype, has_bias, backward;
    int G;                // planes per batch
    int TW;               // warps per team
    int NT;               // teams per CTA
    int threads;          // 32 * TW * NT
    int team_lanes;       // 32 * TW
    int LPP, lpp_shift;   // lanes per plane = team_lanes / G (power of two)
    int n_cg;             // channel groups per image = C / G
    long long n_batches;  // B * n_cg
    int grid;
    int esize, vec;
    unsigned magic_cpr, magic_H;
    WGrid gp;             // unpack: "strips" = column pairs of level 0 (W even), rows split into blocks
    int use_tma;
    WLevel lv[kMaxLevel + 1];
    int plane_floats;     // floats per plane block (all padded level buffers)
    int tplane_floats;    // floats of one plane's T buffer
    int raw_plane_bytes;
    int raw_bytes;        // G planes
    int wstride;          // floats per (plane, conv) filter slot (bias at [K*K])
    int offGY, offG0, pitchG0;   // bwd (float offsets inside the plane block)
    // byte offsets inside a team's slice
    int off_w, off_wg, off_planes, off_tr, off_raw2, team_bytes;
    // byte offsets inside the CTA's dynamic shared memory
    int smBar, smTab, smTeams, smem_bytes;
    // work split: team gt (global index) of n_teams_total.  n_teams_total >= n_cg: the team keeps ONE channel group
    // cg = gt % n_cg for its whole life and takes images r, r + tpc, ... (r = gt / n_cg < tpc); otherwise (tpc = 1)
    // it walks over channel groups gt, gt + n_teams_total, ... and takes every image of each.
    int n_teams_total, tpc;
    int dbg;              // timing experiments only: bit mask of stages to skip (RECNEXT_DBG), 0 in production
    int ws_partial_floats;  // bwd: [tpc][(L+2)][C][wstride] per-team filter-gradient partials
};

struct WPlanOptions {
    int force_G = 0, force_TW = 0, force_NT = 0, force_no_tma = 0, max_warps = 0, dbg = 0;
    int num_sms = 148;
    int smem_limit = 227 * 1024;
};

RC_H WGrid w_grid(int LPP, int rows, int strips) {
    WGrid g;
    g.strips = strips;
    int spr = 1;
    if (strips <= 4) { while (spr < strips) spr *= 2; }
    else spr = rc_round_up(strips, 8);
    g.spr = spr;
    int nrb = LPP / spr;  // row blocks so that one round keeps the plane's lanes busy
    if (nrb < 1) nrb = 1;
    if (nrb > rows) nrb = rows;
    g.rpb = rc_div_up(rows, nrb);
    g.nrb = rc_div_up(rows, g.rpb);
    g.ipp = g.nrb * g.spr;
    g.rounds = rc_div_up(g.ipp, LPP);
    g.m_spr = rc_magic(g.spr);
    return g;
}

// 0 ok; 1 does not fit / not eligible (caller falls back to the big-plane path); 2 bad arguments
RC_H int w_make_plan(WPlan& pl, int B, int C, int H, int W, int K, int L, int mode, int dtype, int wdtype, int has_bias, int backward,
                     const WPlanOptions& opt) {
    if (B < 1 || C < 1 || H < 1 || W < 1 || !(K == 3 || K == 5 || K == 7) || L < 0 || L > kMaxLevel) return 2;
    if (H > 1023 || W > 1023) return 1;
    pl = WPlan();
    pl.B = B; pl.C = C; pl.H = H; pl.W = W; pl.K = K; pl.L = L; pl.mode = mode; pl.dtype = dtype; pl.wdtype = wdtype;
    pl.has_bias = has_bias; pl.backward = backward; pl.dbg = opt.dbg;
    pl.esize = dtype == 0 ? 4 : 2;
    const int pad = K / 2;
    pl.lv[0].H = H; pl.lv[0].W = W;
    for (int l = 1; l <= L; ++l) { pl.lv[l].H = rc_down_size(pl.lv[l - 1].H, K); pl.lv[l].W = rc_down_size(pl.lv[l - 1].W, K); }

    // ---- per-plane buffers ----
    int off = 0;
    for (int l = 0; l <= L; ++l) {
        WLevel& g = pl.lv[l];
        int need = rc_round_up(g.W, kStripW) - kStripW + rc_win_s1(K);
        if (l < L) {
            const int n2 = 2 * (rc_round_up(pl.lv[l + 1].W, kStripW) - kStripW) + rc_win_s2(K);
            if (n2 > need) need = n2;
        }
        if (need < g.W + 2 * pad) need = g.W + 2 * pad;
        g.pitch = rc_round_up(need, 4);
        g.rows = g.H + 2 * pad;
        if (g.rows < pad + 4) g.rows = pad + 4;
        if (l < L) { const int r2 = 2 * (pl.lv[l + 1].H - 1) + K; if (r2 > g.rows) g.rows = r2; }
        g.offS = off; off += g.rows * g.pitch;
        g.offX = g.offGT = g.offGS = -1;
        g.exact2x = (l >= 1 && pl.lv[l - 1].H == 2 * g.H && pl.lv[l - 1].W == 2 * g.W) ? 1 : 0;
        g.up_fast = (g.exact2x && mode == 0 && (pad & 1) == 0) ? 1 : 0;
        g.toff = g.up_fast ? 2 : 0;
        g.tp = rc_round_up(g.W + 2 * g.toff, 4);
        g.magic_W = rc_magic(g.W);
        g.magic_HW = rc_magic(g.H * g.W);
    }
    if (backward) {
        for (int l = 1; l <= L; ++l) {
            WLevel& g = pl.lv[l];
            if (l < L) { g.offX = off; off += g.rows * g.pitch; }
            g.offGT = off; off += g.rows * g.pitch;
            g.offGS = off; off += g.rows * g.pitch;
        }
        pl.offGY = off; off += pl.lv[0].rows * pl.lv[0].pitch;
        // G0 (gradient w.r.t. s_0) is written over S_0 once the final conv's filter gradient has consumed it
        pl.offG0 = pl.lv[0].offS + pad * pl.lv[0].pitch + pad;
        pl.pitchG0 = pl.lv[0].pitch;
    }
    pl.plane_floats = rc_round_up(off, 4);
    int nT = 0;
    for (int l = 1; l <= L; ++l) { const int n = pl.lv[l].H * pl.lv[l].tp; if (n > nT) nT = n; }
    pl.tplane_floats = rc_round_up(nT, 4);
    pl.wstride = rc_round_up(K * K + 1, 4);
    pl.raw_plane_bytes = H * W * pl.esize;
    int vec = 16 / pl.esize;
    while (vec > 1 && (W % vec) != 0) vec >>= 1;
    pl.vec = vec;
    pl.magic_cpr = rc_magic(W / vec);
    pl.magic_H = rc_magic(H);

    // ---- tables (shared by the CTA) ----
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
    tb = rc_round_up(tb, 128);

    // ---- batch size G, team size TW, teams per CTA NT ----
    const int strips0 = rc_div_up(W, kStripW);
    auto team_bytes_for = [&](int G, int TW) -> long {
        const long wbytes = (long)G * (L + 2) * pl.wstride * 4;
        const int spp = (32 * TW / G) > 32 ? (32 * TW / G) / 32 : 1;  // gradient slots per plane (one per warp of a plane)
        const long wgbytes = backward ? wbytes * spp : 0;
        const long planes = (long)G * pl.plane_floats * 4;
        const long raw = rc_round_up(G * pl.raw_plane_bytes, 128);
        long tr = (long)G * pl.tplane_floats * 4;
        if (raw > tr) tr = raw;
        const long raw2 = backward ? raw : 0;
        return rc_round_up((int)(wbytes + wgbytes), 128) + rc_round_up((int)planes, 128) + rc_round_up((int)tr, 128) + raw2;
    };
    const long avail = opt.smem_limit - 256 - tb;
    // As many resident warps as the register budget allows (latency hiding): batch G planes per warp only while
    // that still leaves max_warps teams per SM; planes too big for that get several warps each (TW).
    const int max_warps = opt.max_warps ? opt.max_warps : ((K >= 7 || backward) ? 8 : 16);  // register budget (wdevice.cuh)
    int G = 1;
    if (opt.force_G) G = opt.force_G;
    else {
        int gmax = 32 / strips0;
        if (gmax < 1) gmax = 1;
        for (int cand = 2; cand <= gmax && cand <= 32; cand *= 2) {
            if (C % cand != 0) break;
            if (avail / team_bytes_for(cand, 1) < (max_warps >= 16 ? 12 : max_warps)) break;  // measured: 12 teams of 2 planes beat 16 of 1
            G = cand;
        }
    }
    if (C % G != 0 || G > 32 || (G & (G - 1)) != 0) return 1;
    int TW = 1;
    if (opt.force_TW) TW = opt.force_TW;
    else {
        while (TW < 8 && (avail / team_bytes_for(G, TW)) * TW < 6 && TW * 2 <= max_warps) TW *= 2;  // big planes: >= 6 warps per SM
    }
    if (G > 32 * TW) return 1;
    const long tbz = team_bytes_for(G, TW);
    if (tbz > avail) return 1;
    int fit = (int)(avail / tbz);  // teams that fit in one SM
    int NT = fit;
    if (NT * TW > max_warps) NT = max_warps / TW;
    if (TW > 1 && NT > 15) NT = 15;  // one named barrier per team
    if (opt.force_NT) NT = opt.force_NT;
    if (NT < 1) return 1;
    pl.G = G; pl.TW = TW; pl.NT = NT; pl.team_lanes = 32 * TW; pl.threads = 32 * TW * NT;
    pl.n_cg = C / G;
    pl.n_batches = (long long)B * pl.n_cg;
    pl.raw_bytes = G * pl.raw_plane_bytes;

    // team slice layout
    int o = 0;
    pl.off_w = o; o += G * (L + 2) * pl.wstride * 4;
    pl.off_wg = o;
    if (backward) { const int lpp = 32 * TW / G; o += G * (lpp > 32 ? lpp / 32 : 1) * (L + 2) * pl.wstride * 4; }
    o = rc_round_up(o, 128);
    pl.off_planes = o; o += rc_round_up(G * pl.plane_floats * 4, 128);
    pl.off_tr = o;
    { int tr = G * pl.tplane_floats * 4; const int raw = rc_round_up(pl.raw_bytes, 128); if (raw > tr) tr = raw; o += rc_round_up(tr, 128); }
    pl.off_raw2 = o; if (backward) o += rc_round_up(pl.raw_bytes, 128);
    pl.team_bytes = o;
    if (pl.team_bytes != (int)tbz) return 2;  // layout and estimate must agree

    pl.smBar = 0;
    pl.smTab = 256;  // 2 mbarriers per team (<= 16 teams)
    pl.smTeams = 256 + tb;
    pl.smem_bytes = pl.smTeams + NT * pl.team_bytes;
    if (pl.smem_bytes > opt.smem_limit) return 1;

    // alignment needed by cp.async.bulk: 16-byte sizes and addresses for every batch
    pl.use_tma = (!opt.force_no_tma && (pl.raw_bytes % 16) == 0) ? 1 : 0;

    // ---- item grids ----
    pl.LPP = pl.team_lanes / G;
    pl.gp = w_grid(pl.LPP, H, (W + 1) / 2);
    pl.lpp_shift = 0;
    while ((1 << pl.lpp_shift) < pl.LPP) ++pl.lpp_shift;
    for (int l = 0; l <= L; ++l) {
        WLevel& lg = pl.lv[l];
        lg.g1 = w_grid(pl.LPP, lg.H, rc_div_up(lg.W, kStripW));
        if (l >= 1) {
            lg.g2 = lg.g1;
            const WLevel& ld = pl.lv[l - 1];
            if (lg.up_fast) lg.gu = w_grid(pl.LPP, lg.H, rc_div_up(ld.W + pad, kStripW));  // strips of PADDED columns
            else if (lg.exact2x && mode == 0) lg.gu = w_grid(pl.LPP, lg.H, rc_div_up(ld.W, kStripW));
            else lg.gu = w_grid(pl.LPP, ld.H, rc_div_up(ld.W, kStripW));
            lg.gt = w_grid(pl.LPP, (ld.H + 1) / 2, (ld.W + 1) / 2);  // "strips" = column pairs, rows = row pairs
        }
    }

    // grid: one CTA per SM (persistent teams walk over the batches)
    long total_teams = (long)opt.num_sms * NT;
    int grid = opt.num_sms;
    if (pl.n_batches < total_teams) grid = (int)rc_div_up((int)pl.n_batches, NT);
    if (grid < 1) grid = 1;
    pl.grid = grid;
    pl.n_teams_total = grid * NT;
    pl.tpc = pl.n_teams_total >= pl.n_cg ? pl.n_teams_total / pl.n_cg : 1;
    if (pl.tpc > B) pl.tpc = B;
    pl.ws_partial_floats = backward ? pl.tpc * (L + 2) * C * pl.wstride : 0;
    return 0;
}

}  // namespace recnext
