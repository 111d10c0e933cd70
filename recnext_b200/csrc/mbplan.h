// mbplan.h — launch plan of the TENSOR-CORE RecConv backward (16-bit activations, k = 5): autograd of model/recnext.py:24-34.
//
// Same formulation as the forward (mplan.h): every 5x5 depthwise stencil is five banded-Toeplitz products on the tensor cores
// (mma.sync m16n8k16; tools/tc_probe.cu shows why not tcgen05: its operand fetch costs 39 clocks per 128x16x16 MMA).  The backward
// needs four kinds of products per plane:
//   dgrad    gs = K^T(gt)            the forward conv tile with the filter flipped
//   wgrad    dK[r][s] = sum S[i+r][j+s] gt[i][j] = diagonals of S_r^T . gt      (A = S^T, B = gt, K dimension = image rows)
//   down^T   G_{l-1} += D^T(G_l)     polyphase: output rows of one parity take filter rows of that parity, columns by a
//                                    stride-2 Toeplitz band (two band phases, for even / odd output column tiles)
//   wgrad_s2 dD[r][s] = sum x_{l-1}[2i+r][2j+s] G_l[i][j]  = slope-2 diagonals of X_r^T . G_l
// and the transposed interpolation as a table-driven gather on the CUDA cores.
//
// Shared-memory state of one plane: four pyramid SETS of identical geometry (padded 16-bit level buffers, rows split by parity as
// in the forward):   X  (x_0 .. x_L)        S  (s_l = x_l + u_l; its level 0 later holds gs_0)
//                     GA (gy, gt_1 .. gt_L; interior at column 8)     GB (gs_1 .. gs_L; hosts the forward's T scratch first)
// Gradient buffers that are the B operand of a weight gradient keep their interior at column 8 instead of 2: the 12-column window
// of S that meets an 8-column block of gt is then two aligned 16-byte chunks.
#pragma once
#include "mplan.h"

namespace recnext {

struct MBPlan {
    MPlan f;             // geometry + team shape; l0_bytes / upper_bytes / off_upper / lv[].off describe ONE set, lv[].offT points into GB
    int set_l0, set_up;  // bytes of one plane's level-0 buffer / levels >= 1 of one set
    int offS, offGA, offGB;   // byte offsets of the sets inside a team slice (X at 0; GB holds levels >= 1 only)
    int team_bytes;
    int nregs;           // fragment registers per channel: forward (20 + 10 (L+1)), dgrad (10 (L+1)), down^T (20)
    int regDgrad, regDT;
    int smTabF, smTabG, smFrag, smBias, smSlots, smTeams, smem_bytes;
    int slot_floats;     // floats of one team's weight-gradient slots: TW x G x (L + 2) x 28
    int kmax;            // CTAs that can share one channel group (partials per group in the workspace)
    long ws_floats;      // workspace: [n_cg][kmax][NTEAM][TW][G][(L+2)][28]
    int use_tma;         // raw planes prefetched by TMA bulk copies into dead regions (needs 16-byte multiples and L >= 1)
    int grid;
};

struct MBPlanOptions {
    int force_G = 0, force_TW = 0, force_NT = 0, num_sms = 148, smem_limit = 227 * 1024;
};

// 0 ok; 1 not eligible / does not fit (caller uses the FMA or streamed kernels); 2 bad arguments
RC_H int mb_make_plan(MBPlan& bp, int B, int C, int H, int W, int K, int L, int mode, int dtype, int wdtype, int has_bias, const MBPlanOptions& opt) {
    if (B < 1 || C < 1 || H < 1 || W < 1 || L < 0 || L > kMaxLevel) return 2;
    if (K != 5 || !(dtype == 1 || dtype == 2)) return 1;
    if (H > 1023 || W > 1023 || H * W < 64) return 1;
    bp = MBPlan{};
    MPlan& pl = bp.f;
    pl.B = B; pl.C = C; pl.H = H; pl.W = W; pl.L = L; pl.mode = mode; pl.dtype = dtype; pl.wdtype = wdtype; pl.has_bias = has_bias;
    pl.lv[0].H = H; pl.lv[0].W = W;
    for (int l = 1; l <= L; ++l) { pl.lv[l].H = rc_down_size(pl.lv[l - 1].H, 5); pl.lv[l].W = rc_down_size(pl.lv[l - 1].W, 5); }
    int off = 0;
    for (int l = 0; l <= L; ++l) {
        MLevel& g = pl.lv[l];
        g.NT = rc_div_up(g.W, 8); g.MT = rc_div_up(g.H, 16);
        g.ntc = g.NT <= 1 ? 1 : (g.NT == 2 ? 2 : (g.NT <= 4 ? 4 : 7));
        int kb = g.NT + 2;                                         // interior at column 8 + one chunk of right padding (gradient buffers)
        if (l < L && 2 * rc_div_up(pl.lv[l + 1].W, 8) + 1 > kb) kb = 2 * rc_div_up(pl.lv[l + 1].W, 8) + 1;   // input of `down`
        if (l >= 1 && 2 * g.NT + 3 > 0 && l <= L) { /* wgrad_s2 windows are clamped to the last chunk */ }
        if ((kb & 1) == 0) ++kb;
        g.pitchB = 16 * kb;
        const int rows = g.H + 4, nE = (rows + 1) / 2, nO = rows / 2;
        g.off = off;
        int d = nE * g.pitchB + 16;
        while (((d / 16) & 7) != 4) d += 16;
        g.parDelta = d;
        off += d + nO * g.pitchB + 16;
        off = rc_round_up(off, 128);
        if (l == 0) { bp.set_l0 = off; off = 0; }
        g.exact2x = (l >= 1 && pl.lv[l - 1].H == 2 * g.H && pl.lv[l - 1].W == 2 * g.W) ? 1 : 0;
        if (l >= 1) {
            g.tpB = rc_round_up((g.W + 4) * 2, 16);
            if (((g.tpB / 16) & 1) == 0) g.tpB += 16;
        }
    }
    bp.set_up = off;
    // T_l (forward scratch of the up pass) must fit inside the level-l buffer of GB
    for (int l = 1; l <= L; ++l) {
        const int lvl_bytes = (l < L ? pl.lv[l + 1].off : bp.set_up) - pl.lv[l].off;
        if (pl.lv[l].H * pl.lv[l].tpB + 48 > lvl_bytes) return 1;
    }
    // interpolation tables: forward (IdxLam, 8 bytes per destination) and gather (GatherEntry, 32 bytes per source), generic path for every level
    int tf = 0, tg = 0;
    for (int l = 1; l <= L; ++l) {
        MLevel& g = pl.lv[l];
        const bool fast = g.exact2x && mode == 0;
        if (fast) { g.tabY = -1; g.tabX = -1; }
        else { g.tabY = tf; tf += 8 * pl.lv[l - 1].H; g.tabX = tf; tf += 8 * pl.lv[l - 1].W; }
    }
    tf = rc_round_up(tf, 128);
    for (int l = 1; l <= L; ++l) tg += 32 * (pl.lv[l].H + pl.lv[l].W) + 8 * (pl.lv[l - 1].H + pl.lv[l - 1].W);
    tg = rc_round_up(tg, 128);

    bp.nregs = 20 + 10 * (L + 1) + 10 * (L + 1) + 20;
    bp.regDgrad = 20 + 10 * (L + 1);
    bp.regDT = bp.regDgrad + 10 * (L + 1);
    pl.nregs = bp.nregs;
    int G = 1;
    if (opt.force_G) G = opt.force_G;
    else while (G < 8 && C % (2 * G) == 0 && G * H * W < 128 && 2 * G * bp.nregs * 128 <= 40 * 1024) G *= 2;   // (measured: more teams beat bigger batches)
    if (G < 1 || C % G != 0) return 1;
    pl.G = G; pl.n_cg = C / G;
    pl.l0_bytes = bp.set_l0; pl.upper_bytes = bp.set_up; pl.off_upper = G * bp.set_l0;
    const int set_bytes = G * (bp.set_l0 + bp.set_up);
    bp.offS = set_bytes; bp.offGA = 2 * set_bytes; bp.offGB = 3 * set_bytes;
    bp.team_bytes = 3 * set_bytes + G * bp.set_up;
    // T of the S view lives in GB: tbuf(g, l) = base_S + off_upper + g * upper_bytes + offT
    for (int l = 1; l <= L; ++l) pl.lv[l].offT = (bp.offGB - bp.offS - pl.off_upper) + pl.lv[l].off;

    pl.smBar = 0;
    bp.use_tma = (L >= 1 && ((G * H * W * 2) % 16) == 0 && G * bp.set_up >= G * H * W * 2 && G * H * W * 2 <= 64 * 1024) ? 1 : 0;
    bp.smTabF = 256;
    bp.smTabG = bp.smTabF + tf;
    bp.smFrag = bp.smTabG + tg;
    bp.smBias = bp.smFrag + G * bp.nregs * 128;
    bp.smSlots = rc_round_up(bp.smBias + G * (L + 2) * 4, 128);
    pl.smTab = bp.smTabF; pl.smFrag = bp.smFrag; pl.smBias = bp.smBias;
    // team shape: as many teams as fit; warps per team so that the SM holds ~12-16 warps
    int TW = opt.force_TW ? opt.force_TW : 1;
    int NTEAM = 0;
    for (;;) {
        const int slot_bytes = TW * G * (L + 2) * 28 * 4;
        const long avail = (long)opt.smem_limit - bp.smSlots - 512;
        const int fit = (int)(avail / (bp.team_bytes + slot_bytes));
        if (fit < 1) return 1;
        NTEAM = fit;
        if (NTEAM * TW > 16) NTEAM = 16 / TW;
        if (TW > 1 && NTEAM > 15) NTEAM = 15;
        if (opt.force_TW || NTEAM * TW >= 12 || TW >= 4 || fit * TW * 2 > 16) break;
        TW *= 2;
    }
    if (opt.force_NT && opt.force_NT <= NTEAM) NTEAM = opt.force_NT;
    if (NTEAM < 1) return 1;
    pl.TW = TW; pl.NTEAM = NTEAM; pl.team_lanes = 32 * TW; pl.threads = 32 * TW * NTEAM;
    bp.slot_floats = TW * G * (L + 2) * 28;
    bp.smTeams = rc_round_up(bp.smSlots + NTEAM * bp.slot_floats * 4, 128);
    pl.smTeams = bp.smTeams; pl.team_bytes = bp.team_bytes;
    bp.smem_bytes = bp.smTeams + NTEAM * bp.team_bytes + 256;   // + slack: ragged conv tiles over-read a few chunks past a row
    pl.smem_bytes = bp.smem_bytes;
    if (bp.smem_bytes > opt.smem_limit) return 1;

    pl.rp_shift = m_lane_shift((W & 1) ? W : W / 2, pl.team_lanes);
    for (int l = 1; l <= L; ++l) {
        pl.lv[l].up_shift = m_lane_shift((pl.lv[l - 1].W + 1) / 2, pl.team_lanes);
        pl.lv[l].up_rpg = rc_div_up(pl.lv[l].H, pl.team_lanes >> pl.lv[l].up_shift);
        pl.lv[l].up2_shift = m_lane_shift((pl.lv[l].W + 1) / 2, pl.team_lanes);
        pl.lv[l].up2_rpg = rc_div_up(pl.lv[l].H, pl.team_lanes >> pl.lv[l].up2_shift);
    }
    const long total = (long)pl.n_cg * B;
    long grid = opt.num_sms;
    if (total < grid) grid = total;
    if (grid < 1) grid = 1;
    pl.grid = bp.grid = (int)grid;
    // a CTA takes the items [total * b / grid, total * (b + 1) / grid): a channel group (B items) meets at most kmax CTAs
    const long per = total / grid;   // >= 1
    bp.kmax = (int)((B + per - 1) / per) + 1;
    bp.ws_floats = (long)pl.n_cg * bp.kmax * NTEAM * bp.slot_floats;
    return 0;
}

}  // namespace recnext
