// mbwd.cuh — TENSOR-CORE fused RecConv backward for 16-bit activations, k = 5 (plan and formulation: mbplan.h).
// Reference semantics: the autograd graph of model/recnext.py:24-34 under autocast: gradients between
// ops are 16-bit tensors, every product is accumulated in fp32, filter gradients are fp32.  sm_100a only.
#pragma once
#include "mfwd.cuh"
#include "mbplan.h"

namespace recnext {

__device__ __forceinline__ void m_ldsm4t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void m_ldsm2t(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// views of the four pyramid sets of a team (same geometry, different base)
template <typename T>
struct MBTeam {
    const MBPlan& bp;
    MTeam<T> X, S, A, Bv;      // Bv: GB (levels >= 1 only)
    uint32_t smem32;
    int team, wt, lane, tl;
    __device__ __forceinline__ void sync() const { X.sync(); }
};

// ---- fragment tables: forward (down, convs), flipped convs (dgrad), down^T band phases ---------------------------
// register (channel p, reg, lane) at tab[(p * nregs + reg) * 32 + lane]; a lane holds elements (k = 2 t4 + 8 h + e, n = gq) of a band
template <typename T>
__device__ __forceinline__ void mb_build_frags(const MBPlan& bp, const KernelArgs& a, unsigned char* smem, int cg, int tid, int nthreads) {
    const MPlan& pl = bp.f;
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + bp.smFrag);
    const int L = pl.L;
    const int nf = 20 + 10 * (L + 1);
    for (int idx = tid; idx < pl.G * bp.nregs * 32; idx += nthreads) {
        const int p = idx / (bp.nregs * 32), rem = idx - p * bp.nregs * 32;
        const int reg = rem >> 5, ln = rem & 31;
        const int gq = ln >> 2, t4 = ln & 3;
        const long ch = (long)cg * pl.G + p;
        float v[2] = {0.f, 0.f};
        if (reg < 20) {
            // forward `down` (stride 2): k - 2n in 0..4 over 24 window columns (registers j = 0, 1: m16n8k16; j = 2: m16n8k8)
            if (L > 0 && a.w[0] != nullptr && (reg & 3) != 3) {
                const int r = reg >> 2, kbase = 2 * t4 + 8 * (reg & 3);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int s = kbase + e - 2 * gq;
                    if (s >= 0 && s <= 4) v[e] = rc_load_param(a.w[0], pl.wdtype, ch * 25 + r * 5 + s);
                }
            }
        } else if (reg < nf + 10 * (L + 1)) {
            // convs[j]: forward band (k - n in 0..4); the dgrad registers hold the same band of the FLIPPED filter w'[r][s] = w[4-r][4-s]
            const bool flip = reg >= nf;
            const int rg = flip ? reg - nf : reg - 20;
            const int j = rg / 10, r = (rg % 10) >> 1, kbase = 2 * t4 + 8 * (rg & 1);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int s = kbase + e - gq;
                if (s >= 0 && s <= 4 && a.w[1 + j] != nullptr)
                    v[e] = rc_load_param(a.w[1 + j], pl.wdtype, ch * 25 + (flip ? (4 - r) * 5 + (4 - s) : r * 5 + s));
            }
        } else {
            // down^T: register (r, phase, h): window element k = 2 t4 + 8 h + e, output column n = gq of an even (phase 0) or odd
            // (phase 1) output tile: s = n + 6 - 2k (+ 8 for odd tiles)
            const int rg = reg - nf - 10 * (L + 1), r = rg >> 2, ph = (rg >> 1) & 1, h = rg & 1;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 2 * t4 + 8 * h + e;
                const int s = gq + 6 - 2 * k + 8 * ph;
                if (s >= 0 && s <= 4 && L > 0 && a.w[0] != nullptr) v[e] = rc_load_param(a.w[0], pl.wdtype, ch * 25 + r * 5 + s);
            }
        }
        tab[idx] = MmaT<T>::pack(v[0], v[1]);
    }
    if (pl.has_bias) {
        float* bt = reinterpret_cast<float*>(smem + pl.smBias);
        for (int idx = tid; idx < pl.G * (L + 2); idx += nthreads) {
            const int p = idx / (L + 2), slot = idx - p * (L + 2);
            float v = 0.f;
            if (a.b[slot] && !(slot == 0 && L == 0)) v = rc_load_param(a.b[slot], pl.wdtype, (long)cg * pl.G + p);
            bt[idx] = v;
        }
    }
}

// raw planes (global memory, dense) -> padded level 0 of set `tm` with the interior at column `coff` (2: activations, 8: gradients)
template <typename T>
__device__ __forceinline__ void mb_repack(const MTeam<T>& tm, const T* __restrict__ src, int coff) {
    const MPlan& pl = tm.pl;
    const int H = pl.H, W = pl.W;
    const unsigned short* s16 = reinterpret_cast<const unsigned short*>(src);
    if ((W & 1) == 0 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
        const int CP = W >> 1;
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
        for (int g = 0; g < pl.G; ++g) {
            const MBuf b = tm.buf(g, 0);
            for (int idx = tm.tl; idx < H * CP; idx += pl.team_lanes) {
                const int i = idx / CP, jj = idx - i * CP;
                m_sts32(b.row(i + 2) + 2u * coff + 4u * jj, s32[(g * H + i) * CP + jj]);
            }
        }
    } else {
        for (int g = 0; g < pl.G; ++g) {
            const MBuf b = tm.buf(g, 0);
            for (int idx = tm.tl; idx < H * W; idx += pl.team_lanes) {
                const int i = idx / W, j = idx - i * W;
                m_sts16(b.row(i + 2) + 2u * (coff + j), s16[(g * H + i) * W + j]);
            }
        }
    }
}

// ---- weight gradient of a stride-1 conv: dK[r][s] += sum_{i,j} S[i+r][j+s] gt[i][j] -------------------------------
// S: activation buffer (interior at column 2), Gt: gradient buffer (interior at column 8).  Items (16 S rows, 8 gt columns) are
// dealt to the team's warps; acc[r] is the 16 x 8 tile D_r[m][n] = sum_k S[16 ks + k][8 (c-1) + m] gt[16 ks + k - r][8 c + n]
// (buffer coordinates), whose diagonals m - n = s carry dK[r][s].
template <typename T>
__device__ __forceinline__ void mb_wgrad_s1(const MBuf& S, const MBuf& Gt, int wt, int TW, int lane, float (&acc)[5][4], float& bsum, bool want_bias) {
    const int KS = (S.H + 4 + 15) / 16, NT = (S.W + 7) / 8;
    const int rowMaxS = S.H + 3, rowMaxG = Gt.H + 3;
    const int lr = lane & 7, lm = lane >> 3;
    for (int item = wt; item < KS * NT; item += TW) {
        const int ks = item / NT, c = item - ks * NT + 1;
        // A = S^T: matrices (rows lo, chunk c-1), (rows lo, chunk c), (rows hi, chunk c-1), (rows hi, chunk c)
        int ra = 16 * ks + lr + 8 * (lm >> 1);
        ra = ra < rowMaxS ? ra : rowMaxS;
        uint32_t a0, a1, a2, a3;
        m_ldsm4t(a0, a1, a2, a3, S.row(ra) + (uint32_t)(c - 1 + (lm & 1)) * 16u);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            int rb = 16 * ks + lr + 8 * (lm & 1) - r + 2;
            rb = rb < 0 ? 0 : (rb < rowMaxG ? rb : rowMaxG);
            uint32_t b0, b1;
            m_ldsm2t(b0, b1, Gt.row(rb) + (uint32_t)c * 16u);
            MmaT<T>::mma16(acc[r], a0, a1, a2, a3, b0, b1);
            if (r == 0 && want_bias) {
                const float2 u0 = MmaT<T>::unpack(b0), u1 = MmaT<T>::unpack(b1);
                bsum += (u0.x + u0.y) + (u1.x + u1.y);
            }
        }
    }
}

// ---- weight gradient of the stride-2 `down`: dD[r][s] += sum_{i,j} X[2i+r][2j+s] G[i][j] ------------------------------
// X: activation buffer of level l-1, G: gs buffer of level l (interior at column 2).  acc[r][tt] = tile tt of the 32 x 8 product
// D_r[m][n] = sum_k X[2 (16 ks + k) + r - 2 (r/2) ..][8 (2c-1) + m] G[16 ks + k - r/2][8 c + n]; dD[r][s] sits at m = 2 n + s + 4.
template <typename T>
__device__ __forceinline__ void mb_wgrad_s2(const MBuf& X, const MBuf& G, int wt, int TW, int lane, float (&acc)[5][2][4], float& bsum, bool want_bias) {
    const int KS = (G.H + 2 + 15) / 16, NTg = (G.W + 2 + 7) / 8;
    const int rowMaxX = X.H + 3, rowMaxG = G.H + 3;
    const int xchunks = X.pitchB / 16 - 1;   // last addressable chunk of an X row
    const int lr = lane & 7, lm = lane >> 3;
    for (int item = wt; item < KS * NTg; item += TW) {
        const int ks = item / NTg, c = item - ks * NTg;
        uint32_t A[2][2][4];   // [row parity of X][tile][a0..a3]
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            int rx = 2 * (16 * ks + lr + 8 * (lm >> 1)) + par;
            rx = rx < rowMaxX ? rx : rowMaxX;
            const uint32_t rowa = X.row(rx);
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                int ch = 2 * c - 1 + 2 * tt + (lm & 1);
                ch = ch < 0 ? 0 : (ch < xchunks ? ch : xchunks);
                m_ldsm4t(A[par][tt][0], A[par][tt][1], A[par][tt][2], A[par][tt][3], rowa + (uint32_t)ch * 16u);
            }
        }
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            int rb = 16 * ks + lr + 8 * (lm & 1) - (r >> 1) + 2;
            rb = rb < 0 ? 0 : (rb < rowMaxG ? rb : rowMaxG);
            uint32_t b0, b1;
            m_ldsm2t(b0, b1, G.row(rb) + (uint32_t)c * 16u);
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) MmaT<T>::mma16(acc[r][tt], A[r & 1][tt][0], A[r & 1][tt][1], A[r & 1][tt][2], A[r & 1][tt][3], b0, b1);
            if (r == 0 && want_bias) {
                const float2 u0 = MmaT<T>::unpack(b0), u1 = MmaT<T>::unpack(b1);
                bsum += (u0.x + u0.y) + (u1.x + u1.y);
            }
        }
    }
}

// warp-reduces the 25 filter-gradient values (+ bias) a wgrad pass left in its accumulator tiles and adds them to slot[0..25]
__device__ __forceinline__ void mb_flush25(const float (&pw)[25], float bsum, float* slot, int lane, bool want_bias) {
    float v[32];
    rc_group_reduce<32, 25>(v, pw, lane);
    if (lane < 25) slot[lane] += v[0];
    if (want_bias) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
        if (lane == 0) slot[25] += bsum;
    }
    __syncwarp();
}
__device__ __forceinline__ void mb_extract_s1(const float (&acc)[5][4], float (&pw)[25], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < 25; ++i) pw[i] = 0.f;
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int m = g + 8 * (e >> 1), n = 2 * t + (e & 1), s = m - n;
#pragma unroll
            for (int ss = 0; ss < 5; ++ss) pw[r * 5 + ss] += (s == ss) ? acc[r][e] : 0.f;
        }
}
__device__ __forceinline__ void mb_extract_s2(const float (&acc)[5][2][4], float (&pw)[25], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < 25; ++i) pw[i] = 0.f;
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int tt = 0; tt < 2; ++tt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = 16 * tt + g + 8 * (e >> 1), n = 2 * t + (e & 1), s = m - 2 * n - 4;
#pragma unroll
                for (int ss = 0; ss < 5; ++ss) pw[r * 5 + ss] += (s == ss) ? acc[r][tt][e] : 0.f;
            }
}

// ---- down^T: dst (level l-1) += D^T(src = G of level l).  Output rows of parity py take the filter rows of parity py -----------
// epi(y, x, v0, v1): output image row y, columns x, x + 1
template <typename T, class Epi>
__device__ __forceinline__ void mb_down_t(const MBuf& src, int Hd, int Wd, uint32_t bfr /* lane's column of the down^T registers */, int wt, int TW, int lane, Epi epi) {
    const int NTd = (Wd + 7) / 8, QP = (NTd + 1) / 2;
    const int rowMax = src.H + 3;
    const int lg = lane >> 2, lt = lane & 3;
    const int lr = lane & 7, lm = lane >> 3;
    int item = wt;
    for (int py = 0; py < 2; ++py) {
        const int na = (Hd - py + 1) / 2;               // output rows of this parity
        const int MT = (na + 15) / 16;
        for (; item < MT * QP; item += TW) {
            const int mt = item / QP, qp = item - mt * QP;
            const int a0 = 16 * mt;
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const int r = py + 2 * rr;
                if (r > 4) continue;
                const int delta = (py + 2 - r) / 2;     // input row = a + delta   (py + 2 - r is even)
                int ri = a0 + lr + 8 * (lm & 1) + delta + 2;
                ri = ri < 0 ? 0 : (ri < rowMax ? ri : rowMax);
                uint32_t x0, x1, x2, x3;
                m_ldsm4(x0, x1, x2, x3, src.row(ri) + (uint32_t)(qp + (lm >> 1)) * 16u);
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                    const uint32_t b0 = m_lds32(bfr + (uint32_t)(r * 4 + ph * 2) * 128u), b1 = m_lds32(bfr + (uint32_t)(r * 4 + ph * 2 + 1) * 128u);
                    MmaT<T>::mma16(acc[ph], x0, x1, x2, x3, b0, b1);
                }
            }
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
                const int x = 8 * (2 * qp + ph) + 2 * lt;
                if (x < Wd) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int a = a0 + lg + 8 * h;
                        if (a < na) epi(2 * a + py, x, acc[ph][2 * h], acc[ph][2 * h + 1]);
                    }
                }
            }
        }
        item -= MT * QP;
    }
}

// ---- transposed interpolation: gt (level l, interior at column 8) = up^T(gs of level l-1, interior at column 2) ------------------
template <typename T>
__device__ __forceinline__ void mb_gather(const MBuf& gs, const MBuf& gt, const GatherEntry* __restrict__ gy, const GatherEntry* __restrict__ gx, int tl, int team_lanes) {
    const int Hl = gt.H, Wl = gt.W;
    for (int idx = tl; idx < Hl * Wl; idx += team_lanes) {
        const int iy = idx / Wl, ix = idx - iy * Wl;
        const GatherEntry ey = gy[iy], ex = gx[ix];
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float wy = u == 0 ? ey.w[0] : (u == 1 ? ey.w[1] : (u == 2 ? ey.w[2] : ey.w[3]));
            if (wy == 0.f) continue;
            const uint32_t row = gs.row(ey.d0 + u + 2) + 2u * (ex.d0 + 2);
            float rs = 0.f;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float wx = v == 0 ? ex.w[0] : (v == 1 ? ex.w[1] : (v == 2 ? ex.w[2] : ex.w[3]));
                if (wx != 0.f) rs = fmaf(wx, MmaT<T>::one(m_lds16(row + 2u * v)), rs);
            }
            acc = fmaf(wy, rs, acc);
        }
        const uint32_t w = MmaT<T>::pack(acc, 0.f);
        m_sts16(gt.row(iy + 2) + 2u * (ix + 8), w);
    }
}

template <typename T>
__device__ __forceinline__ void mb_zero(uint32_t saddr, int bytes, int tl, int team_lanes) {
    for (int i = tl; i < bytes / 16; i += team_lanes) asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(saddr + 16u * i), "r"(0u) : "memory");
}
__device__ __forceinline__ void mb_copy(uint32_t dst, uint32_t src, int bytes, int tl, int team_lanes) {
    for (int i = tl; i < bytes / 16; i += team_lanes) {
        uint32_t a, b, c, d;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(src + 16u * i));
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst + 16u * i), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512, 1) recconv_mbwd_kernel(const __grid_constant__ MBPlan bp, const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const MPlan& pl = bp.f;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int team = warp / pl.TW, wt = warp - team * pl.TW;
    const uint32_t smem32 = rc_smem_u32(smem);
    const uint32_t tbase = smem32 + (uint32_t)(bp.smTeams + team * bp.team_bytes);
    const int tl = wt * 32 + lane;
    MTeam<T> tmX{pl, pl, smem32, tbase, team, wt, lane, tl};
    MTeam<T> tmS{pl, pl, smem32, tbase + (uint32_t)bp.offS, team, wt, lane, tl};
    MTeam<T> tmA{pl, pl, smem32, tbase + (uint32_t)bp.offGA, team, wt, lane, tl};
    MTeam<T> tmB{pl, pl, smem32, tbase + (uint32_t)(bp.offGB - pl.off_upper), team, wt, lane, tl};
    const int L = pl.L, G = pl.G, TW = pl.TW;
    const int lg = lane >> 2, lt = lane & 3;
    float* slots = reinterpret_cast<float*>(smem + bp.smSlots) + team * bp.slot_floats;   // [TW][G][(L+2)][28]
    auto slot = [&](int g, int s) { return slots + ((wt * G + g) * (L + 2) + s) * 28; };

    // ---- CTA init: zero every team slice (borders stay zero), interpolation tables
    {
        uint4* z = reinterpret_cast<uint4*>(smem + bp.smTeams);
        const int n16 = pl.NTEAM * bp.team_bytes / 16;
        const uint4 zero = {0u, 0u, 0u, 0u};
        for (int i = tid; i < n16; i += blockDim.x) z[i] = zero;
        int tg = 0;
        for (int l = 1; l <= L; ++l) {
            const MLevel& ls = pl.lv[l];
            const MLevel& ld = pl.lv[l - 1];
            if (ls.tabY >= 0) {
                rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + bp.smTabF + ls.tabY), ls.H, ld.H, pl.mode, tid, blockDim.x);
                rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + bp.smTabF + ls.tabX), ls.W, ld.W, pl.mode, tid, blockDim.x);
            }
            // gather tables of level l: [IdxLam Y][IdxLam X][Gather Y][Gather X]
            IdxLam* ty = reinterpret_cast<IdxLam*>(smem + bp.smTabG + tg);
            IdxLam* tx = ty + ld.H;
            rc_build_fwd_table(ty, ls.H, ld.H, pl.mode, tid, blockDim.x);
            rc_build_fwd_table(tx, ls.W, ld.W, pl.mode, tid, blockDim.x);
            tg += 8 * (ld.H + ld.W) + 32 * (ls.H + ls.W);
        }
    }
    __syncthreads();
    {
        int tg = 0;
        for (int l = 1; l <= L; ++l) {
            const MLevel& ls = pl.lv[l];
            const MLevel& ld = pl.lv[l - 1];
            IdxLam* ty = reinterpret_cast<IdxLam*>(smem + bp.smTabG + tg);
            IdxLam* tx = ty + ld.H;
            GatherEntry* gy = reinterpret_cast<GatherEntry*>(tx + ld.W);
            GatherEntry* gx = gy + ls.H;
            rc_build_gather_table(gy, ty, ls.H, ld.H, pl.mode, tid, blockDim.x);
            rc_build_gather_table(gx, tx, ls.W, ld.W, pl.mode, tid, blockDim.x);
            tg += 8 * (ld.H + ld.W) + 32 * (ls.H + ls.W);
        }
    }
    __syncthreads();

    // RECNEXT_PROF=1: team 0 of CTA 0 records clock64() after every stage of its first planes (timing experiments only)
    int pidx = 0;
    auto stamp = [&]() {
        if (a.prof != nullptr && blockIdx.x == 0 && tid == 0 && pidx < 4000) a.prof[pidx++] = clock64();
    };
    // raw planes arrive by TMA bulk copies one plane ahead: x into GA level 0 (gy is dead after the final conv's dgrad), gy into the
    // S set's upper block (s_l is dead once the last level's weight gradient is done); both are unpacked at the start of the plane
    const uint32_t bar_x = smem32 + (uint32_t)(16 * team), bar_g = bar_x + 8u;
    uint32_t ph_x = 0, ph_g = 0;
    const bool tma = bp.use_tma != 0;
    if (tma && tl == 0) { rc_mbar_init(bar_x, 1); rc_mbar_init(bar_g, 1); }
    rc_fence_proxy_async();
    __syncthreads();
    const int plane_elems = pl.H * pl.W;
    const T* gxin = reinterpret_cast<const T*>(a.x);
    const T* ggy = reinterpret_cast<const T*>(a.gy);
    T* gout = reinterpret_cast<T*>(a.out);
    const long total = (long)pl.n_cg * pl.B;
    const long start = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1) / gridDim.x;
    if (start >= end) return;
    const int cg_first = (int)(start / pl.B), cg_last = (int)((end - 1) / pl.B);
    const bool wb = pl.has_bias != 0;
    for (int cg = cg_first; cg <= cg_last; ++cg) {
        __syncthreads();
        mb_build_frags<T>(bp, a, smem, cg, tid, blockDim.x);
        for (int i = tid; i < pl.NTEAM * bp.slot_floats; i += blockDim.x) reinterpret_cast<float*>(smem + bp.smSlots)[i] = 0.f;
        __syncthreads();
        const long cg_lo = (long)cg * pl.B > start ? (long)cg * pl.B : start;
        const long cg_end = ((long)(cg + 1) * pl.B) < end ? ((long)(cg + 1) * pl.B) : end;
        for (long my = cg_lo + team; my < cg_end; my += pl.NTEAM) {
            const int n = (int)(my - (long)cg * pl.B);
            const long p0 = ((long)n * pl.C + (long)cg * G) * (long)plane_elems;
            const bool has_next = my + pl.NTEAM < cg_end;
            const long p0n = p0 + (long)pl.NTEAM * pl.C * (long)plane_elems;   // same channel group, image n + NTEAM
            unsigned char* rawx = smem + bp.smTeams + team * bp.team_bytes + bp.offGA;
            unsigned char* rawg = smem + bp.smTeams + team * bp.team_bytes + bp.offS + pl.off_upper;
            const uint32_t raw_bytes = (uint32_t)(G * plane_elems * 2);
            // ================= recompute the pyramid (model/recnext.py:27-33) =================
            stamp();
            if (tma) {
                if (my == cg_lo + team) {   // first plane of this team in the group: nothing was prefetched
                    rc_fence_proxy_async();
                    tmX.sync();
                    if (tl == 0) {
                        rc_mbar_expect_tx(bar_x, raw_bytes); rc_bulk_g2s(rawx, gxin + p0, raw_bytes, bar_x);
                        rc_mbar_expect_tx(bar_g, raw_bytes); rc_bulk_g2s(rawg, ggy + p0, raw_bytes, bar_g);
                    }
                }
                while (!rc_mbar_try_wait(bar_x, ph_x)) {}
                ph_x ^= 1u;
                mb_repack<T>(tmX, reinterpret_cast<const T*>(rawx), 2);
                tmX.sync();
                while (!rc_mbar_try_wait(bar_g, ph_g)) {}
                ph_g ^= 1u;
                // the raw x planes sat on top of GA level 0: clear, then gy -> GA level 0 (interior at column 8)
                mb_zero<T>(tbase + (uint32_t)bp.offGA, G * bp.set_l0, tl, pl.team_lanes);
                tmX.sync();
                mb_repack<T>(tmA, reinterpret_cast<const T*>(rawg), 8);
                tmX.sync();
            } else {
                mb_repack<T>(tmX, gxin + p0, 2);
                tmX.sync();
            }
            stamp();
#pragma unroll 1
            for (int l = 1; l <= L; ++l) {
                const MLevel& lo = pl.lv[l];
                for (int g = 0, mt = wt; g < G; mt += TW) {
                    if (mt >= lo.MT) { mt -= lo.MT + TW; ++g; continue; }
                    const MBuf in = tmX.buf(g, l - 1);
                    const MBuf out = tmX.buf(g, l);
                    const int i0 = mt * 16, Wo = lo.W;
                    const int ia = i0 + lg, ib = ia + 8;
                    const bool va = ia < lo.H, vb = ib < lo.H;
                    const uint32_t da = out.row(ia + 2) + 4u + 4u * lt, db = out.row(ib + 2) + 4u + 4u * lt;
                    m_conv_rows<T, true>(in, lo.ntc, i0, lo.NT, tmX.frag(g, 0), tmX.bias(g, 0), lane, [&](int q, const float (&acc)[4]) {
                        const int c = 8 * q + 2 * lt;
                        if (c < Wo) {
                            const bool pair = c + 1 < Wo;
                            if (va) m_sts32(da + 16u * q, MmaT<T>::pack(acc[0], pair ? acc[1] : 0.f));
                            if (vb) m_sts32(db + 16u * q, MmaT<T>::pack(acc[2], pair ? acc[3] : 0.f));
                        }
                    });
                }
                tmX.sync();
                stamp();
            }
            // S := X (s_l starts as x_l), then the up pass on S with T in GB
            mb_copy(tbase + (uint32_t)bp.offS, tbase, G * (bp.set_l0 + bp.set_up), tl, pl.team_lanes);
            tmX.sync();
            stamp();
#pragma unroll 1
            for (int l = L; l >= 1; --l) {
                const MLevel& lv = pl.lv[l];
                for (int g = 0, mt = wt; g < G; mt += TW) {
                    if (mt >= lv.MT) { mt -= lv.MT + TW; ++g; continue; }
                    const MBuf in = tmS.buf(g, l);
                    const int i0 = mt * 16, Wl = lv.W;
                    const int ia = i0 + 2 * lg;
                    const bool va = ia < lv.H, vb = ia + 1 < lv.H;
                    const uint32_t ta = tmS.tbuf(g, l) + (uint32_t)(ia * lv.tpB) + 4u + 4u * lt, tpB = (uint32_t)lv.tpB;
                    m_conv_rows<T, false>(in, lv.ntc, i0, lv.NT, tmS.frag(g, 20 + 10 * (L - l)), tmS.bias(g, 1 + (L - l)), lane,
                                          [&](int q, const float (&acc)[4]) {
                        const int c = 8 * q + 2 * lt;
                        if (c < Wl) {
                            const bool pair = c + 1 < Wl;
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (h ? vb : va) {
                                    const uint32_t w = MmaT<T>::pack(acc[2 * h], pair ? acc[2 * h + 1] : acc[2 * h]);
                                    const uint32_t ad = ta + h * tpB + 16u * q;
                                    m_sts32(ad, w);
                                    if (c == 0) m_sts16(ad - 2u, w);
                                    if (c + 1 == Wl - 1) m_sts16(ad + 4u, w >> 16);
                                }
                            }
                        }
                    });
                }
                tmS.sync();
                stamp();
                if (lv.tabY < 0) m_up2x_add<T>(tmS, l);
                else m_up_add<T>(tmS, smem, l);
                tmS.sync();
                stamp();
            }
            // GB hosted T: restore its zeros; gy -> GA level 0 (interior at column 8) unless the TMA path already did it
            if (L > 0) mb_zero<T>(tbase + (uint32_t)bp.offGB, G * bp.set_up, tl, pl.team_lanes);
            if (!tma) mb_repack<T>(tmA, ggy + p0, 8);
            tmX.sync();
            stamp();
            // ================= final conv: dK_L = corr(s_0, gy); gs_0 = K_L^T(gy) -> S level 0 =================
            for (int g = 0; g < G; ++g) {
                float acc[5][4] = {};
                float bsum = 0.f;
                mb_wgrad_s1<T>(tmS.buf(g, 0), tmA.buf(g, 0), wt, TW, lane, acc, bsum, wb);
                float pw[25];
                mb_extract_s1(acc, pw, lane);
                mb_flush25(pw, bsum, slot(g, 1 + L), lane, wb);
            }
            tmX.sync();   // every warp is done reading s_0 before gs_0 overwrites it
            stamp();
            {
                const MLevel& lv = pl.lv[0];
                const int H = pl.H, W = pl.W;
                for (int g = 0, mt = wt; g < G; mt += TW) {
                    if (mt >= lv.MT) { mt -= lv.MT + TW; ++g; continue; }
                    const MBuf in = tmA.buf(g, 0);
                    const MBuf out = tmS.buf(g, 0);
                    const int i0 = mt * 16;
                    const int ia = i0 + 2 * lg;
                    const bool va = ia < H, vb = ia + 1 < H;
                    const uint32_t ra = out.row(ia + 2), rb = out.row(ia + 3);
                    m_conv_rows<T, false>(in, (lv.NT + 1 <= 4) ? 4 : 7, i0, lv.NT + 1, tmX.frag(g, bp.regDgrad + 10 * L), 0.f, lane, [&](int q, const float (&acc)[4]) {
                        const int c = 8 * q - 6 + 2 * lt;   // gradient buffers keep their interior at column 8: the tile is shifted by 6 columns
                        if (c >= 0 && c < W) {
                            const bool pair = c + 1 < W;
                            if (va) m_sts32(ra + 4u + 2u * c, MmaT<T>::pack(acc[0], pair ? acc[1] : 0.f));
                            if (vb) m_sts32(rb + 4u + 2u * c, MmaT<T>::pack(acc[2], pair ? acc[3] : 0.f));
                        }
                    });
                }
            }
            if (tma && has_next) {   // gy is dead: the next plane's x lands on GA level 0
                rc_fence_proxy_async();
                tmX.sync();
                if (tl == 0) { rc_mbar_expect_tx(bar_x, raw_bytes); rc_bulk_g2s(rawx, gxin + p0n, raw_bytes, bar_x); }
            } else tmX.sync();
            stamp();
            // ================= per level: gt_l = up^T(gs_{l-1}); dK_{L-l} = corr(s_l, gt_l); gs_l = K_{L-l}^T(gt_l) =================
            {
                int tg = 0;
#pragma unroll 1
                for (int l = 1; l <= L; ++l) {
                    const MLevel& ls = pl.lv[l];
                    const MLevel& ld = pl.lv[l - 1];
                    const GatherEntry* gy = reinterpret_cast<const GatherEntry*>(smem + bp.smTabG + tg + 8 * (ld.H + ld.W));
                    const GatherEntry* gx = gy + ls.H;
                    tg += 8 * (ld.H + ld.W) + 32 * (ls.H + ls.W);
                    for (int g = 0; g < G; ++g) mb_gather<T>(l == 1 ? tmS.buf(g, 0) : tmB.buf(g, l - 1), tmA.buf(g, l), gy, gx, tl, pl.team_lanes);
                    tmX.sync();
                    stamp();
                    for (int g = 0; g < G; ++g) {
                        float acc[5][4] = {};
                        float bsum = 0.f;
                        mb_wgrad_s1<T>(tmS.buf(g, l), tmA.buf(g, l), wt, TW, lane, acc, bsum, wb);
                        float pw[25];
                        mb_extract_s1(acc, pw, lane);
                        mb_flush25(pw, bsum, slot(g, 1 + (L - l)), lane, wb);
                    }
                    stamp();
                    for (int g = 0, mt = wt; g < G; mt += TW) {
                        if (mt >= ls.MT) { mt -= ls.MT + TW; ++g; continue; }
                        const MBuf in = tmA.buf(g, l);
                        const MBuf out = tmB.buf(g, l);
                        const int i0 = mt * 16, Wl = ls.W;
                        const int ia = i0 + 2 * lg;
                        const bool va = ia < ls.H, vb = ia + 1 < ls.H;
                        const uint32_t ra = out.row(ia + 2), rb = out.row(ia + 3);
                        m_conv_rows<T, false>(in, (ls.NT + 1 <= 4) ? 4 : 7, i0, ls.NT + 1, tmX.frag(g, bp.regDgrad + 10 * (L - l)), 0.f, lane, [&](int q, const float (&acc)[4]) {
                            const int c = 8 * q - 6 + 2 * lt;
                            if (c >= 0 && c < Wl) {
                                const bool pair = c + 1 < Wl;
                                if (va) m_sts32(ra + 4u + 2u * c, MmaT<T>::pack(acc[0], pair ? acc[1] : 0.f));
                                if (vb) m_sts32(rb + 4u + 2u * c, MmaT<T>::pack(acc[2], pair ? acc[3] : 0.f));
                            }
                        });
                    }
                    tmX.sync();
                    stamp();
                }
            }
            if (tma && has_next) {   // every s_l (l >= 1) is dead: the next plane's gy lands on the S set's upper block
                rc_fence_proxy_async();
                tmX.sync();
                if (tl == 0) { rc_mbar_expect_tx(bar_g, raw_bytes); rc_bulk_g2s(rawg, ggy + p0n, raw_bytes, bar_g); }
            }
            // ================= down chain: dD += corr_s2(x_{l-1}, G_l); G_{l-1} = gs_{l-1} + D^T(G_l); gx = G_0 =================
#pragma unroll 1
            for (int l = L; l >= 1; --l) {
                const MLevel& ld = pl.lv[l - 1];
                for (int g = 0; g < G; ++g) {
                    float acc[5][2][4] = {};
                    float bsum = 0.f;
                    mb_wgrad_s2<T>(tmX.buf(g, l - 1), tmB.buf(g, l), wt, TW, lane, acc, bsum, wb);
                    float pw[25];
                    mb_extract_s2(acc, pw, lane);
                    mb_flush25(pw, bsum, slot(g, 0), lane, wb);
                }
                stamp();
                for (int g = 0; g < G; ++g) {
                    const MBuf src = tmB.buf(g, l);
                    const uint32_t bfr = tmX.frag(g, bp.regDT);
                    if (l > 1) {
                        const MBuf dst = tmB.buf(g, l - 1);
                        mb_down_t<T>(src, ld.H, ld.W, bfr, wt, TW, lane, [&](int y, int x, float v0, float v1) {
                            const uint32_t ad = dst.row(y + 2) + 4u + 2u * x;
                            const float2 old = MmaT<T>::unpack(m_lds32(ad));
                            m_sts32(ad, MmaT<T>::pack(old.x + v0, (x + 1 < ld.W) ? old.y + v1 : 0.f));
                        });
                    } else {
                        const MBuf gs0 = tmS.buf(g, 0);
                        T* dstp = gout + p0 + (long)g * plane_elems;
                        const bool even = (pl.W & 1) == 0;
                        mb_down_t<T>(src, pl.H, pl.W, bfr, wt, TW, lane, [&](int y, int x, float v0, float v1) {
                            const float2 old = MmaT<T>::unpack(m_lds32(gs0.row(y + 2) + 4u + 2u * x));
                            const uint32_t w = MmaT<T>::pack(old.x + v0, old.y + v1);
                            T* d = dstp + y * pl.W + x;
                            if (even) *reinterpret_cast<uint32_t*>(d) = w;
                            else {
                                *reinterpret_cast<unsigned short*>(d) = (unsigned short)(w & 0xffffu);
                                if (x + 1 < pl.W) *reinterpret_cast<unsigned short*>(d + 1) = (unsigned short)(w >> 16);
                            }
                        });
                    }
                }
                tmX.sync();
                stamp();
            }
            if (L == 0) {   // plain depthwise conv: gx = gs_0
                for (int g = 0; g < G; ++g) {
                    const MBuf gs0 = tmS.buf(g, 0);
                    const unsigned short* dummy = nullptr; (void)dummy;
                    T* dstp = gout + p0 + (long)g * plane_elems;
                    for (int idx = tl; idx < plane_elems; idx += pl.team_lanes) {
                        const int y = idx / pl.W, x = idx - y * pl.W;
                        reinterpret_cast<unsigned short*>(dstp)[idx] = (unsigned short)m_lds16(gs0.row(y + 2) + 4u + 2u * x);
                    }
                }
                tmX.sync();
            }
        }
        // ---- flush the channel group's filter-gradient slots: partial[cg][k][team][wt][g][slot][28], k = this CTA's index inside the group
        __syncthreads();
        {
            const long first_item = (long)cg * pl.B;
            // first CTA that holds an item of this group: smallest b with total * (b + 1) / grid > first_item
            long b0 = first_item * gridDim.x / total;
            while (total * (b0 + 1) / gridDim.x <= first_item) ++b0;
            while (b0 > 0 && total * b0 / gridDim.x > first_item) --b0;
            const int k = (int)(blockIdx.x - b0);
            float* dst = a.partial + ((long)cg * bp.kmax + k) * (long)(pl.NTEAM * bp.slot_floats);
            const float* src = reinterpret_cast<const float*>(smem + bp.smSlots);
            for (int i = tid; i < pl.NTEAM * bp.slot_floats; i += blockDim.x) dst[i] = src[i];
        }
    }
}

// gw[slot][c][e] = sum over (k, team, warp) of the partial slots of channel c's group, fixed order; gb likewise (e == 25)
__global__ void recconv_mbwd_finalize(const float* __restrict__ partial, float* __restrict__ gw, float* __restrict__ gb, int n_cg, int kmax, int nteam, int TW,
                                      int G, int L) {
    const long total = (long)n_cg * G * (L + 2) * 26;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int e = (int)(i % 26);
        long r = i / 26;
        const int s = (int)(r % (L + 2)); r /= (L + 2);
        const int g = (int)(r % G);
        const int cg = (int)(r / G);
        float acc = 0.f;
        for (int k = 0; k < kmax; ++k)
            for (int t = 0; t < nteam; ++t)
                for (int w = 0; w < TW; ++w)
                    acc += partial[((((long)cg * kmax + k) * nteam + t) * TW + w) * (long)(G * (L + 2) * 28) + (g * (L + 2) + s) * 28 + e];
        const long c = (long)cg * G + g;
        const long C = (long)n_cg * G;
        if (e < 25) gw[((long)s * C + c) * 25 + e] = acc;
        else if (gb) gb[(long)s * C + c] = acc;
    }
}

inline cudaError_t mb_launch_bwd(const MBPlan& bp, const KernelArgs& a, float* gw, float* gb, cudaStream_t stream) {
    cudaError_t err = cudaMemsetAsync(a.partial, 0, (size_t)bp.ws_floats * sizeof(float), stream);
    if (err != cudaSuccess) return err;
    if (bp.f.dtype == 1) {
        static DeviceOnce configured = {};
        err = rc_once_per_device(configured, [] { return cudaFuncSetAttribute(recconv_mbwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
        if (err != cudaSuccess) return err;
        recconv_mbwd_kernel<__nv_bfloat16><<<bp.grid, bp.f.threads, bp.smem_bytes, stream>>>(bp, a);
    } else {
        static DeviceOnce configured = {};
        err = rc_once_per_device(configured, [] { return cudaFuncSetAttribute(recconv_mbwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
        if (err != cudaSuccess) return err;
        recconv_mbwd_kernel<__half><<<bp.grid, bp.f.threads, bp.smem_bytes, stream>>>(bp, a);
    }
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    const long total = (long)bp.f.n_cg * bp.f.G * (bp.f.L + 2) * 26;
    recconv_mbwd_finalize<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a.partial, gw, gb, bp.f.n_cg, bp.kmax, bp.f.NTEAM, bp.f.TW, bp.f.G, bp.f.L);
    return cudaGetLastError();
}

}  // namespace recnext
