// wbody.cuh — per-team schedules of the team-resident RecConv kernels (forward and backward).
//
// The schedule is a template over an execution context `Ctx` so that the identical stage code runs under CUDA
// (wdevice.cuh) and on the CPU for the schedule tests (tests/emu):
//   ctx.stage(f)               f(tl) for every lane tl of the team, then a team barrier
//   ctx.load(dst, src, bytes, q)  start moving one batch of raw planes global -> shared on queue q = 0 | 1 (TMA bulk copy
//                              issued by lane 0; the caller has just passed a team barrier); ctx.wait(q) blocks until
//                              it has landed
//   ctx.reduce<N>(acc, ...)    sums per-lane filter-gradient partials over the lanes of a plane
//
// Forward, per batch of G planes (reference model/recnext.py:24-34):
//   unpack x -> S_0 | S_l = down(S_{l-1}) l=1..L | for l=L..1: T = convs[L-l](S_l); S_{l-1} += up(T) |
//   y = convs[L](S_0) -> global.
// Backward recomputes S_l on chip and then runs the autograd chain of SURVEY.md §3.1:
//   dconvs[L] = corr(S_0, gy); G_0 = convs[L]^T gy (written over S_0) |
//   for l=1..L: GT_l = up^T(G_{l-1}); dconvs[L-l] = corr(S_l, GT_l); G_l = convs[L-l]^T GT_l |
//   for l=L..1: ddown += corr_s2(x_{l-1}, G_l); G_{l-1} += down^T(G_l) | gx = G_0 -> global.
#pragma once
#include "recconv_body.cuh"  // KernelArgs
#include "wstages.cuh"

namespace recnext {

struct WWork {  // the (channel group, image) pairs of one team, in order
    int cg, n, cg_step, n0, n_step, n_cg, B;
    RC_HD bool valid() const { return cg < n_cg; }
    RC_HD void next() {
        if (n + n_step < B) n += n_step;
        else { n = n0; cg += cg_step; }
    }
};

RC_HD WWork w_work(const WPlan& pl, int gteam) {
    WWork w;
    w.n_cg = pl.n_cg; w.B = pl.B;
    if (pl.n_teams_total >= pl.n_cg) {
        w.cg = gteam % pl.n_cg; w.cg_step = pl.n_cg; w.n0 = gteam / pl.n_cg; w.n_step = pl.tpc;
        if (w.n0 >= pl.tpc) w.cg = pl.n_cg;  // spare team
    } else {
        w.cg = gteam; w.cg_step = pl.n_teams_total; w.n0 = 0; w.n_step = 1;
    }
    w.n = w.n0;
    return w;
}

// CTA-wide initialisation: zero every team slice (borders of the padded buffers stay zero for the whole kernel)
// and build the interpolation tables.  Call with every thread, then synchronise the CTA.
RC_HD void w_cta_init(const WPlan& pl, unsigned char* smem, int tid, int nthreads) {
    float4* z = reinterpret_cast<float4*>(smem + pl.smTeams);
    const int n4 = pl.NT * pl.team_bytes / 16;
    const float4 zero = {0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < n4; i += nthreads) z[i] = zero;
    for (int l = 1; l <= pl.L; ++l) {
        rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + pl.smTab + pl.lv[l].tabY), pl.lv[l].H, pl.lv[l - 1].H, pl.mode, tid, nthreads);
        rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + pl.smTab + pl.lv[l].tabX), pl.lv[l].W, pl.lv[l - 1].W, pl.mode, tid, nthreads);
    }
}
RC_HD void w_cta_init_bwd(const WPlan& pl, unsigned char* smem, int tid, int nthreads) {  // after a CTA barrier
    for (int l = 1; l <= pl.L; ++l) {
        rc_build_gather_table(reinterpret_cast<GatherEntry*>(smem + pl.smTab + pl.lv[l].gatY),
                              reinterpret_cast<const IdxLam*>(smem + pl.smTab + pl.lv[l].tabY), pl.lv[l].H, pl.lv[l - 1].H, pl.mode, tid, nthreads);
        rc_build_gather_table(reinterpret_cast<GatherEntry*>(smem + pl.smTab + pl.lv[l].gatX),
                              reinterpret_cast<const IdxLam*>(smem + pl.smTab + pl.lv[l].tabX), pl.lv[l].W, pl.lv[l - 1].W, pl.mode, tid, nthreads);
    }
}

// filters (and biases) of channel group cg -> the team's shared-memory slots (fp32), bias at [K*K]
RC_HD void w_load_filters(const WPlan& pl, const KernelArgs& a, float* wsm, int cg, int tl) {
    const int KK = pl.K * pl.K;
    const int per_plane = (pl.L + 2) * pl.wstride;
    for (int i = tl; i < pl.G * per_plane; i += pl.team_lanes) {
        const int p = i / per_plane, r = i - p * per_plane;
        const int slot = r / pl.wstride, e = r - slot * pl.wstride;
        const long ch = (long)cg * pl.G + p;
        float v = 0.f;
        if (!(slot == 0 && pl.L == 0)) {  // `down` exists in the state_dict but is unused at level 0
            if (e < KK) v = rc_load_param(a.w[slot], pl.wdtype, ch * KK + e);
            else if (e == KK && pl.has_bias && a.b[slot]) v = rc_load_param(a.b[slot], pl.wdtype, ch);
        }
        wsm[i] = v;
    }
}

struct WLanePos {
    int tl, g, jl;
    float* pb;         // this lane's plane block
    const float* wp;   // this lane's plane's filter slots
    float* Tg;         // this lane's plane of the T buffer
};

RC_HD WLanePos w_lane_pos(const WPlan& pl, unsigned char* tsm, int tl) {
    WLanePos p;
    p.tl = tl; p.g = tl >> pl.lpp_shift; p.jl = tl & (pl.LPP - 1);
    p.pb = reinterpret_cast<float*>(tsm + pl.off_planes) + (long)p.g * pl.plane_floats;
    p.wp = reinterpret_cast<const float*>(tsm + pl.off_w) + p.g * (pl.L + 2) * pl.wstride;
    p.Tg = reinterpret_cast<float*>(tsm + pl.off_tr) + (long)p.g * pl.tplane_floats;
    return p;
}

// the recompute shared by forward and backward: S_l for all levels (s_l = x_l + u_l), optionally keeping x_l
template <int K, class Ctx>
RC_HD void w_pyramid(Ctx& ctx, const WPlan& pl, unsigned char* smem, unsigned char* tsm, bool keep_x) {
    constexpr int PAD = K / 2;
    const int LPP = pl.LPP;
    const bool use_bias = pl.has_bias != 0;
    for (int l = 1; l <= pl.L; ++l) {  // model/recnext.py:27-29
        if (pl.dbg & 2) break;
        ctx.stage([&](int tl) {
            const WLanePos p = w_lane_pos(pl, tsm, tl);
            const WLevel& gi = pl.lv[l - 1];
            const WLevel& go = pl.lv[l];
            float* dst = p.pb + go.offS + PAD * go.pitch + PAD;
            float* dstx = (keep_x && go.offX >= 0) ? p.pb + go.offX + PAD * go.pitch + PAD : nullptr;
            const int pitch = go.pitch, Wo = go.W;
            w_conv_s2<K>(p.pb + gi.offS, gi.pitch, go.H, go.g2, p.wp, use_bias, p.jl, LPP,
                         [&](int row, int c0, const float (&acc)[kStripW]) {
                             w_store_level(dst, pitch, Wo, row, c0, acc);
                             if (dstx) w_store_level(dstx, pitch, Wo, row, c0, acc);
                         });
        });
    }
    for (int l = pl.L; l >= 1; --l) {  // model/recnext.py:31-33 ; convs[L-l] acts on level l
        if (!(pl.dbg & 4)) ctx.stage([&](int tl) {
            const WLanePos p = w_lane_pos(pl, tsm, tl);
            const WLevel& gl = pl.lv[l];
            float* T = p.Tg;
            const int tp = gl.tp, Wl = gl.W;
            const bool fast = gl.up_fast != 0;
            w_conv_s1<K, false>(p.pb + gl.offS, gl.pitch, gl.H, gl.g1, p.wp + (1 + (pl.L - l)) * pl.wstride, use_bias, p.jl, LPP,
                                [&](int row, int c0, const float (&acc)[kStripW]) {
                                    if (fast) w_store_T_fast(T, tp, Wl, row, c0, acc);
                                    else *reinterpret_cast<float4*>(T + row * tp + c0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                                });
        });
        if (!(pl.dbg & 8)) ctx.stage([&](int tl) {
            const WLanePos p = w_lane_pos(pl, tsm, tl);
            const WLevel& gl = pl.lv[l];
            const WLevel& gd = pl.lv[l - 1];
            float* dstS = p.pb + gd.offS + PAD * gd.pitch + PAD;
            if (gl.up_fast)
                w_up2x_add_fast(p.pb + gd.offS + PAD * gd.pitch, gd.pitch, PAD, gd.W, p.Tg, gl.tp, gl.H, gl.gu, p.jl, LPP);
            else if (gl.exact2x && pl.mode == 0)
                w_up2x_add(dstS, gd.pitch, gd.W, p.Tg, gl.tp, gl.H, gl.W, gl.gu, p.jl, LPP);
            else
                w_up_add(dstS, gd.pitch, gd.H, gd.W, p.Tg, gl.tp, gl.H, gl.W, reinterpret_cast<const IdxLam*>(smem + pl.smTab + gl.tabY),
                         reinterpret_cast<const IdxLam*>(smem + pl.smTab + gl.tabX), pl.mode, gl.gu, p.jl, LPP);
        });
    }
}

template <typename T, class Ctx>
RC_HD void w_unpack(Ctx& ctx, const WPlan& pl, unsigned char* tsm, const void* raw, int off_dst) {
    if (pl.dbg & 1) return;
    if ((pl.W & 1) == 0 && (pl.K / 2) % 2 == 0) {  // column pairs: conflict-free 32/64-bit loads and 64-bit stores
        ctx.stage([&](int tl) {
            const WLanePos p = w_lane_pos(pl, tsm, tl);
            w_unpack_pairs<T>(reinterpret_cast<const T*>(raw) + (long)p.g * pl.H * pl.W, pl.H, pl.W,
                              p.pb + off_dst + (pl.K / 2) * pl.lv[0].pitch + pl.K / 2, pl.lv[0].pitch, pl.gp, p.jl, pl.LPP);
        });
        return;
    }
    ctx.stage([&](int tl) {
        rc_unpack_unit<T>(reinterpret_cast<const T*>(raw), pl.G, pl.H, pl.W, pl.vec, pl.magic_cpr, pl.magic_H,
                          reinterpret_cast<float*>(tsm + pl.off_planes), pl.plane_floats, off_dst, pl.lv[0].pitch, pl.K / 2, tl,
                          pl.team_lanes);
    });
}

template <int K, typename T, class Ctx>
RC_HD void w_forward_team(Ctx& ctx, const WPlan& pl, const KernelArgs& a, unsigned char* smem, int team, int gteam) {
    unsigned char* tsm = smem + pl.smTeams + (long)team * pl.team_bytes;
    float* wsm = reinterpret_cast<float*>(tsm + pl.off_w);
    unsigned char* raw = tsm + pl.off_tr;  // aliases T: a batch is loaded only after the last use of T
    const long plane_elems = (long)pl.H * pl.W;
    const T* gx = reinterpret_cast<const T*>(a.x);
    T* gy = reinterpret_cast<T*>(a.out);
    const bool use_bias = pl.has_bias != 0;

    WWork cur = w_work(pl, gteam);
    if (!cur.valid()) return;
    auto src_of = [&](const WWork& w) { return gx + ((long)w.n * pl.C + (long)w.cg * pl.G) * plane_elems; };
    ctx.load(raw, src_of(cur), pl.raw_bytes, 0);
    int cg_loaded = -1;
    while (cur.valid()) {
        if (cur.cg != cg_loaded) {
            ctx.stage([&](int tl) { w_load_filters(pl, a, wsm, cur.cg, tl); });
            cg_loaded = cur.cg;
        }
        ctx.wait(0);
        w_unpack<T>(ctx, pl, tsm, raw, pl.lv[0].offS);
        WWork nxt = cur;
        nxt.next();
        if (pl.L == 0 && nxt.valid()) ctx.load(raw, src_of(nxt), pl.raw_bytes, 0);
        w_pyramid<K>(ctx, pl, smem, tsm, false);
        if (pl.L > 0 && nxt.valid()) ctx.load(raw, src_of(nxt), pl.raw_bytes, 0);  // T is dead: prefetch under the last conv
        if (!(pl.dbg & 16)) ctx.stage([&](int tl) {  // model/recnext.py:34
            const WLanePos p = w_lane_pos(pl, tsm, tl);
            const WLevel& g0 = pl.lv[0];
            T* dst = gy + ((long)cur.n * pl.C + (long)cur.cg * pl.G + p.g) * plane_elems;
            const int W = pl.W;
            w_conv_s1<K, false>(p.pb + g0.offS, g0.pitch, g0.H, g0.g1, p.wp + (1 + pl.L) * pl.wstride, use_bias, p.jl, pl.LPP,
                                [&](int row, int c0, const float (&acc)[kStripW]) { w_store_global4<T>(dst, W, row, c0, acc); });
        });
        cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
// Filter-gradient accumulation slots of a team: [G * spp][(L+2)][wstride] floats, spp = max(1, LPP / 32) slots per
// plane (one per warp of a multi-warp plane).  They persist over all batches of one channel group.
RC_HD int w_slots_per_plane(const WPlan& pl) { return pl.LPP > 32 ? pl.LPP / 32 : 1; }

template <int K, typename T, class Ctx>
RC_HD void w_backward_team(Ctx& ctx, const WPlan& pl, const KernelArgs& a, unsigned char* smem, int team, int gteam) {
    constexpr int PAD = K / 2;
    constexpr int NA = K * K + 1;
    unsigned char* tsm = smem + pl.smTeams + (long)team * pl.team_bytes;
    float* wsm = reinterpret_cast<float*>(tsm + pl.off_w);
    float* wg = reinterpret_cast<float*>(tsm + pl.off_wg);
    unsigned char* raw1 = tsm + pl.off_tr;    // x (aliases T: reloaded for the `down` filter gradient)
    unsigned char* raw2 = tsm + pl.off_raw2;  // gy
    const long plane_elems = (long)pl.H * pl.W;
    const T* gx_in = reinterpret_cast<const T*>(a.x);
    const T* gg_in = reinterpret_cast<const T*>(a.gy);
    T* g_out = reinterpret_cast<T*>(a.out);
    const int L = pl.L, LPP = pl.LPP;
    const WLevel& g0 = pl.lv[0];
    const int spp = w_slots_per_plane(pl);
    const int nslot_floats = pl.G * spp * (L + 2) * pl.wstride;

    WWork cur = w_work(pl, gteam);
    if (!cur.valid()) return;
    const int rank = pl.n_teams_total >= pl.n_cg ? cur.n0 : 0;  // which of the tpc partial sets this team fills
    auto off_of = [&](const WWork& w) { return ((long)w.n * pl.C + (long)w.cg * pl.G) * plane_elems; };
    auto slot_of = [&](const WLanePos& p, int stage) {
        const int s = spp > 1 ? (p.g * spp + (p.jl >> 5)) : p.g;
        return wg + ((long)s * (L + 2) + stage) * pl.wstride;
    };
    auto flush = [&](int cg, int rank) {  // slots -> workspace [rank][(L+2)][C][wstride], warps of a plane summed in order
        ctx.stage([&](int tl) {
            const int per_plane = (L + 2) * pl.wstride;
            for (int i = tl; i < pl.G * per_plane; i += pl.team_lanes) {
                const int p = i / per_plane, r = i - p * per_plane;
                const int stage = r / pl.wstride, e = r - stage * pl.wstride;
                float sum = 0.f;
                for (int q = 0; q < spp; ++q) sum += wg[((long)(p * spp + q) * (L + 2) + stage) * pl.wstride + e];
                const long ch = (long)cg * pl.G + p;
                a.partial[(((long)rank * (L + 2) + stage) * pl.C + ch) * pl.wstride + e] = sum;
            }
        });
    };

    ctx.load(raw1, gx_in + off_of(cur), pl.raw_bytes, 0);
    ctx.load(raw2, gg_in + off_of(cur), pl.raw_bytes, 1);
    int cg_loaded = -1;
    while (cur.valid()) {
        if (cur.cg != cg_loaded) {
            if (cg_loaded >= 0) flush(cg_loaded, rank);
            ctx.stage([&](int tl) {
                w_load_filters(pl, a, wsm, cur.cg, tl);
                for (int i = tl; i < nslot_floats; i += pl.team_lanes) wg[i] = 0.f;
            });
            cg_loaded = cur.cg;
        }
        WWork nxt = cur;
        nxt.next();
        ctx.wait(0);
        w_unpack<T>(ctx, pl, tsm, raw1, g0.offS);
        w_pyramid<K>(ctx, pl, smem, tsm, true);
        if (L > 0) ctx.load(raw1, gx_in + off_of(cur), pl.raw_bytes, 0);  // T is dead: x again, for the `down` filter gradient
        else if (nxt.valid()) ctx.load(raw1, gx_in + off_of(nxt), pl.raw_bytes, 0);
        ctx.wait(1);
        w_unpack<T>(ctx, pl, tsm, raw2, pl.offGY);
        if (nxt.valid()) ctx.load(raw2, gg_in + off_of(nxt), pl.raw_bytes, 1);

        // y = convs[L](s_0): filter gradient, then (S_0 is dead) the input gradient G_0 written over it
        if (!(pl.dbg & 32)) ctx.stage([&](int tl) {
            const WLanePos p = w_lane_pos(pl, tsm, tl);
            float acc[NA];
#pragma unroll
            for (int i = 0; i < NA; ++i) acc[i] = 0.f;
            w_wgrad_s1<K>(p.pb + g0.offS, p.pb + pl.offGY, g0.pitch, g0.H, g0.g1, p.jl, LPP, acc);
            ctx.template reduce<NA>(pl, tl, acc, slot_of(p, 1 + L));
        });
        if (!(pl.dbg & 64)) ctx.stage([&](int tl) {
            const WLanePos p = w_lane_pos(pl, tsm, tl);
            float* G0 = p.pb + pl.offG0;
            T* dsto = g_out + off_of(cur) + (long)p.g * plane_elems;
            const int W = pl.W, gp = pl.pitchG0;
            const bool direct = (L == 0);
            w_conv_s1<K, true>(p.pb + pl.offGY, g0.pitch, g0.H, g0.g1, p.wp + (1 + L) * pl.wstride, false, p.jl, LPP,
                               [&](int row, int c0, const float (&v)[kStripW]) {
                                   if (direct) { w_store_global4<T>(dsto, W, row, c0, v); return; }
                                   w_store_level(G0, gp, W, row, c0, v);
                               });
        });
        if (L > 0) {
            ctx.wait(0);
            w_unpack<T>(ctx, pl, tsm, raw1, pl.offGY);  // gy is dead: x_0 (padded) for the `down` filter gradient
            if (nxt.valid()) ctx.load(raw1, gx_in + off_of(nxt), pl.raw_bytes, 0);
        }

        for (int l = 1; l <= L; ++l) {
            if (!(pl.dbg & 128)) ctx.stage([&](int tl) {  // GT_l = up^T(G_{l-1})
                const WLanePos p = w_lane_pos(pl, tsm, tl);
                const WLevel& gl = pl.lv[l];
                const WLevel& gd = pl.lv[l - 1];
                const float* gsrc = (l == 1) ? p.pb + pl.offG0 : p.pb + gd.offGS + PAD * gd.pitch + PAD;
                const int gpitch = (l == 1) ? pl.pitchG0 : gd.pitch;
                w_up_bwd(p.pb + gl.offGT, gl.pitch, PAD, gl.H, gl.W, gsrc, gpitch,
                         reinterpret_cast<const GatherEntry*>(smem + pl.smTab + gl.gatY),
                         reinterpret_cast<const GatherEntry*>(smem + pl.smTab + gl.gatX), gl.magic_W, p.jl, LPP);
            });
            if (!(pl.dbg & 256)) ctx.stage([&](int tl) {  // dconvs[L-l] = corr(S_l, GT_l);  G_l = convs[L-l]^T GT_l
                const WLanePos p = w_lane_pos(pl, tsm, tl);
                const WLevel& gl = pl.lv[l];
                float acc[NA];
#pragma unroll
                for (int i = 0; i < NA; ++i) acc[i] = 0.f;
                w_wgrad_s1<K>(p.pb + gl.offS, p.pb + gl.offGT, gl.pitch, gl.H, gl.g1, p.jl, LPP, acc);
                ctx.template reduce<NA>(pl, tl, acc, slot_of(p, 1 + (L - l)));
                float* dst = p.pb + gl.offGS + PAD * gl.pitch + PAD;
                const int pitch = gl.pitch, Wl = gl.W;
                w_conv_s1<K, true>(p.pb + gl.offGT, gl.pitch, gl.H, gl.g1, p.wp + (1 + (L - l)) * pl.wstride, false, p.jl, LPP,
                                   [&](int row, int c0, const float (&v)[kStripW]) { w_store_level(dst, pitch, Wl, row, c0, v); });
            });
        }
        for (int l = L; l >= 1; --l) {  // x_l = down(x_{l-1}): filter gradient (summed over levels), input gradient into G_{l-1}
            if (!(pl.dbg & 512)) ctx.stage([&](int tl) {
                const WLanePos p = w_lane_pos(pl, tsm, tl);
                const WLevel& gl = pl.lv[l];
                const WLevel& gd = pl.lv[l - 1];
                float acc[NA];
#pragma unroll
                for (int i = 0; i < NA; ++i) acc[i] = 0.f;
                const float* X = (l == 1) ? p.pb + pl.offGY : p.pb + gd.offX;
                w_wgrad_s2<K>(X, gd.pitch, p.pb + gl.offGS, gl.pitch, gl.H, gl.g2, p.jl, LPP, acc);
                ctx.template reduce<NA>(pl, tl, acc, slot_of(p, 0));
                if (l == 1) {
                    const float* G0 = p.pb + pl.offG0;
                    T* dsto = g_out + off_of(cur) + (long)p.g * plane_elems;
                    const int W = pl.W, gp = pl.pitchG0;
                    w_convT_s2<K>(p.pb + gl.offGS, gl.pitch, p.wp, gd.H, gd.W, gl.gt, p.jl, LPP,
                                  [&](int i, int j, float v) { dsto[(long)i * W + j] = Elem<T>::from_f(G0[i * gp + j] + v); });
                } else {
                    float* dst = p.pb + gd.offGS + PAD * gd.pitch + PAD;
                    const int pitch = gd.pitch;
                    w_convT_s2<K>(p.pb + gl.offGS, gl.pitch, p.wp, gd.H, gd.W, gl.gt, p.jl, LPP,
                                  [&](int i, int j, float v) { dst[i * pitch + j] += v; });
                }
            });
        }
        cur = nxt;
    }
    flush(cg_loaded, rank);
}

}  // namespace recnext
