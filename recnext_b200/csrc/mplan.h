// mplan.h — launch plan of the TENSOR-CORE RecConv forward (16-bit activations, k = 5).
//
// Why.  Measured on B200 (tools/mma_probe.cu): scalar FP32 FMA tops out at 31.6 TFMA/s (FFMA2 gives nothing more),
// while the warp-level HMMA path sustains 278 TMAC/s.  The fused block needs ~48 FMA per element, so on the FMA
// pipe a bf16 forward cannot pass ~33 % of the HBM roofline.  Here every depthwise 5x5 stencil is evaluated as five
// banded-Toeplitz matrix products on the tensor cores:
//     out[i, 8q + n] = sum_r  sum_k  S[i + r, 8q + k] * Bt_r[k, n],     Bt_r[k, n] = w[r, k - n]  (0 <= k - n <= 4)
// i.e. one m16n8k16 MMA per (16 output rows, 8 output columns, filter row r): A = 16 image rows x 16 columns read
// with ldmatrix straight out of the padded level buffer (a filter row is just another row address), B = the
// channel's Toeplitz band (two registers per filter row, precomputed per channel), fp32 accumulators.  The stride-2
// `down` filter is the same with Bt_r[k, n] = w[r, k - 2n] over 24 columns (m16n8k16 + m16n8k8).  3.7x more MACs
// than the direct form, on a pipe that is 8.8x faster.  Products of 16-bit operands are exact in fp32 and every
// intermediate is rounded to the activation dtype exactly where the reference's autocast graph rounds it
// (conv outputs, f + x, interpolate), so this path tracks the reference's own bf16 numerics (model/recnext.py:24-34).
//
// Data layout.  A CTA serves one CHANNEL GROUP (G consecutive channels) at a time and walks over images; its
// per-channel Toeplitz fragments live in one table shared by all warps.  A TEAM (TW warps) owns a batch of G planes:
// every pyramid level is a padded 16-bit buffer in the team's slice (interior at (+2, +2), zero borders written
// once), rows split by PARITY into two arrays so that both the stride-1 (8 consecutive rows) and the stride-2
// (8 alternate rows) ldmatrix row sets are bank-conflict free: row pitch = odd number of 16-byte chunks, odd array
// offset = 4 chunks mod 8.  Raw planes arrive by one TMA bulk copy per batch (prefetched a whole batch ahead).
#pragma once
#include "recconv_plan.h"

namespace recnext {

struct MLevel {
    int H, W;
    int NT, MT;          // n-tiles (8 columns) and m-tiles (16 rows) of this level
    int ntc;             // n-tiles handled per pass of the tile routine (1, 2, 4, 7)
    int pitchB;          // row pitch in bytes (odd multiple of 16)
    int off, parDelta;   // byte offset of the even-row array inside a plane block; distance to the odd-row array
    int tpB;             // l >= 1: row pitch (bytes) of the T buffer of this level; interior at element 2
    int offT;            // l >= 1: byte offset (upper block) of T_l.  T_l is written when every level deeper than l is dead,
                         // so it starts where level l+1 starts (the zero borders it tramples are restored per batch)
    int tabY, tabX;      // l >= 1: byte offsets (table region) of IdxLam[H_{l-1}] / IdxLam[W_{l-1}]; -1 if the level takes
                         // the fixed-stencil exact-2x bilinear path and needs no table
    int exact2x;         // level l-1 is exactly 2x this level
    int up_shift;        // l >= 1: up-add lane mapping over level l-1: lanes per row group = 1 << up_shift
    int up_rpg;          // l >= 1: exact-2x path: source rows per row group
    int up2_shift, up2_rpg;  // l >= 1: exact-2x path with two source columns per lane (the default)
};

struct MPlan {
    int B, C, H, W, L, mode, dtype, wdtype, has_bias;
    int variant;         // 0: RecConv2d forward; 1: `down` only (x -> x_1 in global memory: RecAttn2d.down[0]);
                         // 2: y = conv(x + interpolate(z)) with an external low-resolution operand z (RecAttn2d tail)
    int z_shift;         // variant 2: lane mapping of the z -> T copy (lanes per row group = 1 << z_shift)
    int G, TW, NTEAM, threads, team_lanes;
    int n_cg;
    int use_tma;
    int rp_shift;        // repack lane mapping (column pairs of level 0, or columns if W is odd)
    MLevel lv[kMaxLevel + 1];
    int l0_bytes;        // level-0 buffer of one plane (the team slice starts with G of them)
    int upper_bytes;     // levels 1..L + T of one plane (MLevel::off of l >= 1 and offT are relative to this block)
    int zero_bytes;      // leading part of an upper block that holds padded level buffers (re-zeroed per batch: the raw
                         // batch and the T buffers of shallower levels land on top of it)
    int raw_bytes;       // G raw planes
    int off_upper;       // byte offset (team slice) of the upper region = the TMA landing buffer: a batch is loaded
                         // only after the last use of levels >= 1 and T (under the final conv)
    int team_bytes;
    int nregs;           // Toeplitz fragment registers per channel: 20 (down) + 10 per conv
    int smBar, smTab, smFrag, smBias, smTeams, smem_bytes;
    int grid;
    int dbg;
};

struct MPlanOptions {
    int force_G = 0, force_TW = 0, force_NT = 0, force_no_tma = 0, max_warps = 16, dbg = 0;
    int variant = 0, zH = 0, zW = 0;   // variant 2: size of the external operand z (becomes level 1)
    int num_sms = 148;
    int smem_limit = 227 * 1024;
};

RC_HD constexpr int m_pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }
RC_HD constexpr int m_log2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

RC_HD constexpr int m_lane_shift(int pairs, int team_lanes) {
    int p = m_pow2_ceil(pairs);
    if (p > team_lanes) p = team_lanes;
    return m_log2(p);
}

// 0 ok; 1 not eligible / does not fit (caller uses the FMA kernels); 2 bad arguments
RC_HD constexpr int m_make_plan(MPlan& pl, int B, int C, int H, int W, int K, int L, int mode, int dtype, int wdtype, int has_bias,
                     const MPlanOptions& opt) {
    if (B < 1 || C < 1 || H < 1 || W < 1 || L < 0 || L > kMaxLevel) return 2;
    if (K != 5 || !(dtype == 1 || dtype == 2)) return 1;
    if (H > 1023 || W > 1023) return 1;
    pl = MPlan{};
    pl.B = B; pl.C = C; pl.H = H; pl.W = W; pl.L = L; pl.mode = mode; pl.dtype = dtype; pl.wdtype = wdtype; pl.has_bias = has_bias;
    pl.dbg = opt.dbg;
    pl.variant = opt.variant;
    if (opt.variant != 0 && L != 1) return 2;
    pl.lv[0].H = H; pl.lv[0].W = W;
    for (int l = 1; l <= L; ++l) { pl.lv[l].H = rc_down_size(pl.lv[l - 1].H, 5); pl.lv[l].W = rc_down_size(pl.lv[l - 1].W, 5); }
    if (opt.variant == 2) {
        if (opt.zH < 1 || opt.zW < 1 || opt.zH > 1023 || opt.zW > 1023) return 2;
        pl.lv[1].H = opt.zH; pl.lv[1].W = opt.zW;
    }
    for (int l = 0; l <= L; ++l) {
        MLevel& g = pl.lv[l];
        g.NT = rc_div_up(g.W, 8); g.MT = rc_div_up(g.H, 16);
        g.ntc = g.NT <= 1 ? 1 : (g.NT == 2 ? 2 : (g.NT <= 4 ? 4 : 7));
    }
    // level buffers: level 0 on its own, levels >= 1 + T in the "upper" block
    int off = 0;
    for (int l = 0; l <= L; ++l) {
        MLevel& g = pl.lv[l];
        int kb = g.NT + 1;                                        // as the input of a stride-1 conv
        if (l < L && 2 * pl.lv[l + 1].NT + 1 > kb) kb = 2 * pl.lv[l + 1].NT + 1;  // as the input of `down`
        if ((kb & 1) == 0) ++kb;
        g.pitchB = 16 * kb;
        const int rows = g.H + 4, nE = (rows + 1) / 2, nO = rows / 2;
        g.off = off;
        int d = nE * g.pitchB + 16;                                // + slack for the one-chunk over-read of a paired load
        while (((d / 16) & 7) != 4) d += 16;
        g.parDelta = d;
        off += d + nO * g.pitchB + 16;
        off = rc_round_up(off, 128);
        if (l == 0) { pl.l0_bytes = off; off = 0; }
        g.exact2x = (l >= 1 && pl.lv[l - 1].H == 2 * g.H && pl.lv[l - 1].W == 2 * g.W) ? 1 : 0;
        if (l >= 1) {
            // T rows are written from MMA fragments (8 rows two apart x 4 words per store): a pitch of 8 (mod 64) bytes puts
            // those 32 words in 32 different banks
            // T rows are ldmatrix operands of the tensor-core upsample (16-byte aligned, odd number of chunks)
            g.tpB = rc_round_up((g.W + 4) * 2, 16);
            if (((g.tpB / 16) & 1) == 0) g.tpB += 16;
        }
    }
    pl.zero_bytes = off;
    {
        int top = off;
        for (int l = 1; l <= L; ++l) {
            MLevel& g = pl.lv[l];
            g.offT = l < L ? pl.lv[l + 1].off : off;
            const int e = g.offT + rc_round_up(g.H * g.tpB + 32, 128);   // + slack: the last window of a T row may over-read one chunk
            if (e > top) top = e;
        }
        pl.upper_bytes = top;
    }

    // tables shared by the CTA (levels on the generic interpolation path only)
    int tb = 0;
    for (int l = 1; l <= L; ++l) {
        MLevel& g = pl.lv[l];
        if (g.exact2x && mode == 0 && !(opt.dbg & 1)) { g.tabY = -1; g.tabX = -1; continue; }
        g.tabY = tb; tb += 8 * pl.lv[l - 1].H;
        g.tabX = tb; tb += 8 * pl.lv[l - 1].W;
    }
    tb = rc_round_up(tb, 128);

    // planes per batch: the raw batch must be a whole number of 16-byte chunks for the bulk copy, and tiny planes
    // are batched to amortise the per-batch overhead
    const int plane_raw = H * W * 2;
    int G = 1;
    if (opt.force_G) G = opt.force_G;
    else {
        const int frag_bytes = (20 + 10 * (L + 1)) * 128;          // per channel
        while (G < 16 && C % (2 * G) == 0 && (((G * plane_raw) % 16) != 0 || (G * H * W < 512 && 2 * G * frag_bytes <= 48 * 1024))) G *= 2;
    }
    if (G < 1 || C % G != 0) return 1;
    pl.G = G; pl.n_cg = C / G;
    pl.raw_bytes = G * plane_raw;
    pl.use_tma = (!opt.force_no_tma && (pl.raw_bytes % 16) == 0 && pl.raw_bytes <= 64 * 1024) ? 1 : 0;
    pl.nregs = 20 + 10 * (L + 1);

    pl.off_upper = G * pl.l0_bytes;
    {
        int up = G * pl.upper_bytes;
        if (pl.use_tma && rc_round_up(pl.raw_bytes, 128) > up) up = rc_round_up(pl.raw_bytes, 128);
        pl.team_bytes = pl.off_upper + up;
    }
    pl.smBar = 0;
    pl.smTab = 256;
    pl.smFrag = pl.smTab + tb;
    pl.smBias = pl.smFrag + G * pl.nregs * 128;
    pl.smTeams = rc_round_up(pl.smBias + G * (L + 2) * 4, 128);
    const long avail = (long)opt.smem_limit - pl.smTeams;
    if (avail < pl.team_bytes) return 1;
    const int fit = (int)(avail / pl.team_bytes);
    int max_warps = opt.max_warps > 24 ? 24 : opt.max_warps;   // 16: any kernel (128 registers); 24: the small-plane specialised kernels
    if (max_warps < 1) max_warps = 1;
    int TW = 1;
    if (opt.force_TW) TW = opt.force_TW;
    else { while (TW < 8 && fit * TW < 8 && TW * 2 <= max_warps) TW *= 2; }
    int NTEAM = fit;
    if (NTEAM * TW > max_warps) NTEAM = max_warps / TW;
    if (TW > 1 && NTEAM > 15) NTEAM = 15;   // one named barrier per team
    if (opt.force_NT) NTEAM = opt.force_NT;
    if (NTEAM < 1 || NTEAM > fit || NTEAM * TW > max_warps || NTEAM > 24) return 1;
    pl.TW = TW; pl.NTEAM = NTEAM; pl.team_lanes = 32 * TW; pl.threads = 32 * TW * NTEAM;
    pl.smem_bytes = pl.smTeams + NTEAM * pl.team_bytes;
    if (pl.smem_bytes > opt.smem_limit) return 1;

    // lane mappings of the element-wise stages: (row group, column pair)
    pl.rp_shift = m_lane_shift((W & 1) ? W : W / 2, pl.team_lanes);
    pl.z_shift = opt.variant == 2 ? m_lane_shift(pl.lv[1].W, pl.team_lanes) : 0;
    for (int l = 1; l <= L; ++l) {
        pl.lv[l].up_shift = m_lane_shift((pl.lv[l - 1].W + 1) / 2, pl.team_lanes);
        pl.lv[l].up_rpg = rc_div_up(pl.lv[l].H, pl.team_lanes >> pl.lv[l].up_shift);
        pl.lv[l].up2_shift = m_lane_shift((pl.lv[l].W + 1) / 2, pl.team_lanes);
        pl.lv[l].up2_rpg = rc_div_up(pl.lv[l].H, pl.team_lanes >> pl.lv[l].up2_shift);
    }

    // persistent grid: one CTA per SM; with less work than team slots, spread it over as many SMs as possible
    // (a batch is one warp-serial chain of stages: an idle SM is worth more than a full one)
    const long total = (long)pl.n_cg * B;
    long grid = opt.num_sms;
    if (total < grid) grid = total;
    if (grid < 1) grid = 1;
    pl.grid = (int)grid;
    return 0;
}

// warps per SM of the specialised kernels: planes up to 28x28 need <= 80 registers per thread (ptxas), so 24 warps fit
// (measured: 24 warps buy nothing — the kernels are bound by shared-memory bandwidth, not latency — so this stays at 16)
RC_HD constexpr int m_static_max_warps(int H, int W) { return (H <= 0 && W <= 0) ? 24 : 16; }

// The layout-defining part of a plan for fixed plane geometry, evaluated at COMPILE time by the specialised kernels
// (mfwd.cuh): every level size, pitch and offset folds into immediates.  The host compares it field by field with
// the run-time plan before choosing a specialised kernel; B, C, grid, bias, parameter dtype and TMA eligibility
// stay run-time values.
RC_HD constexpr MPlan m_static_plan(int H, int W, int L, int G, int dtype, int variant = 0, int mode = 0) {
    MPlan pl{};
    MPlanOptions o{};
    o.force_G = G;
    o.max_warps = m_static_max_warps(H, W);
    o.variant = variant;
    if (variant == 2) { o.zH = rc_down_size(H, 5); o.zW = rc_down_size(W, 5); }   // RecAttn2d: z has the size of down(x)
    m_make_plan(pl, 1, G, H, W, 5, L, mode, dtype, dtype, 0, o);
    return pl;
}
// copies the run-time fields of `rt` into a static plan so that the two can be compared with memcmp
RC_H MPlan m_static_patched(MPlan st, const MPlan& rt) {
    st.B = rt.B; st.C = rt.C; st.n_cg = rt.n_cg; st.has_bias = rt.has_bias; st.wdtype = rt.wdtype; st.use_tma = rt.use_tma;
    st.grid = rt.grid; st.dbg = rt.dbg;
    return st;
}

}  // namespace recnext
