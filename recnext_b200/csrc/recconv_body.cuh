// recconv_body.cuh — stage schedules of the fused RecConv forward / backward kernels.
//
// The schedule is a template over an execution context `Ctx`:
//   ctx.run(n, f)         f(pos) for every thread whose lane-in-unit is < n (device: this thread; host: a loop)
//   ctx.sync(n)           barrier among the first n lanes of each unit (device: __syncwarp or a named barrier;
//                         units never interact).  A stage that needs few lanes is skipped by the other warps.
//   ctx.cta_sync()        whole-CTA barrier (prologue and epilogue only)
//   ctx.load_begin / load_wait / store / store_drain   movement of one unit's raw planes (TMA bulk copies)
//   ctx.wgrad_commit      reduction of per-lane weight-gradient partials into the CTA's accumulation slots
// so the identical stage code runs under CUDA (recconv_device.cuh) and on the CPU for logic tests (tests/emu).
//
// Forward, per image of the CTA's channel group (reference model/recnext.py:24-34):
//   unpack x -> S_0 | S_l = down(S_{l-1}) l=1..L | for l=L..1: T = convs[L-l](S_l); S_{l-1} += up(T) |
//   y = convs[L](S_0) -> raw out.
// Backward recomputes S_l on chip and then runs the autograd chain of SURVEY.md §3.1:
//   dconvs[L] = corr(S_0, gy); G_0 = convs[L]^T gy |
//   for l=1..L: GT_l = up^T(G_{l-1}); dconvs[L-l] = corr(S_l, GT_l); G_l = convs[L-l]^T GT_l |
//   for l=L..1: ddown += corr_s2(x_{l-1}, G_l); G_{l-1} += down^T(G_l) | gx = G_0.
#pragma once
#include "recconv_stages.cuh"

namespace recnext {

struct KernelArgs {
    const void* x;
    const void* gy;   // bwd
    void* out;        // fwd: y; bwd: gx
    float* partial;   // bwd: [n_chunk][(L+2)][C][wstride]
    const void* w[kMaxLevel + 2];  // slot 0 = down, 1+j = convs[j]
    const void* b[kMaxLevel + 2];  // may be null
    long long* prof;  // timing experiments only (RECNEXT_PROF): per-stage clock64() of team 0 of CTA 0, else null
};

struct ThreadPos {
    int tid, u, ul, p, lane, c;  // thread, unit, lane within unit, plane slot, lane within team, channel
    bool active;                 // owns a real plane
};

RC_HD ThreadPos rc_thread_pos(const Plan& pl, int tid, int cg) {
    ThreadPos t;
    t.tid = tid;
    t.u = tid / pl.unit_lanes;
    t.ul = tid - t.u * pl.unit_lanes;
    t.p = tid / pl.g;
    t.lane = tid - t.p * pl.g;
    t.c = cg * pl.P + t.p;
    t.active = t.p < pl.P && t.c < pl.C;
    return t;
}

struct UnitIO {  // what one unit moves per image
    void* sm_x; void* sm_g; void* sm_out;
    long goff;    // element offset of the unit's first plane inside image 0
    long bytes;   // bytes of the unit's active planes
    int nact;     // active planes of the unit
};

template <typename T>
RC_HD UnitIO rc_unit_io(const Plan& pl, unsigned char* smem, int cg, int u) {
    UnitIO io;
    const int p0 = u * pl.ppu;
    int nact = pl.C - (cg * pl.P + p0);
    nact = nact < 0 ? 0 : (nact > pl.ppu ? pl.ppu : nact);
    io.nact = nact;
    io.sm_x = smem + pl.smRawX + (long)u * pl.raw_unit_bytes;
    io.sm_g = smem + pl.smRawG + (long)u * pl.raw_unit_bytes;
    io.sm_out = smem + pl.smRawOut + (long)u * pl.raw_unit_bytes;
    io.goff = ((long)cg * pl.P + p0) * pl.H * pl.W;
    io.bytes = (long)nact * pl.H * pl.W * (long)sizeof(T);
    return io;
}

template <class Ctx>
RC_HD void rc_prologue(Ctx& ctx, const Plan& pl, const KernelArgs& a, unsigned char* smem, int cg) {
    const int nact = (pl.C - cg * pl.P) < pl.P ? (pl.C - cg * pl.P) : pl.P;
    ctx.run_all([&](int tid) {
        // zero every plane block (borders must stay zero for the whole kernel) and the accumulation slots
        float4* z = reinterpret_cast<float4*>(smem + pl.smPlanes);
        const int n4 = pl.P * pl.plane_floats / 4;
        const float4 zero = {0.f, 0.f, 0.f, 0.f};
        for (int i = tid; i < n4; i += pl.T) z[i] = zero;
        float* wg = reinterpret_cast<float*>(smem + pl.smWG);
        const int nwg = pl.nslots * (pl.L + 2) * pl.wstride;
        for (int i = tid; i < nwg; i += pl.T) wg[i] = 0.f;
        // filters of this channel group -> shared (fp32), bias at [K*K] of each slot
        float* wsm = reinterpret_cast<float*>(smem + pl.smW);
        const int KK = pl.K * pl.K;
        const int per_plane = (pl.L + 2) * pl.wstride;
        for (int i = tid; i < pl.P * per_plane; i += pl.T) {
            const int p = i / per_plane, r = i - p * per_plane;
            const int slot = r / pl.wstride, e = r - slot * pl.wstride;
            const long ch = (long)cg * pl.P + p;
            float v = 0.f;
            if (p < nact && !(slot == 0 && pl.L == 0)) {  // `down` exists in the state_dict but is unused at level 0
                if (e < KK) v = rc_load_param(a.w[slot], pl.wdtype, ch * KK + e);
                else if (e == KK && pl.has_bias && a.b[slot]) v = rc_load_param(a.b[slot], pl.wdtype, ch);
            }
            wsm[i] = v;
        }
        for (int l = 1; l <= pl.L; ++l) {
            rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + pl.smTab + pl.lv[l].tabY), pl.lv[l].H, pl.lv[l - 1].H,
                               pl.mode, tid, pl.T);
            rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + pl.smTab + pl.lv[l].tabX), pl.lv[l].W, pl.lv[l - 1].W,
                               pl.mode, tid, pl.T);
        }
    });
    ctx.cta_sync();
    if (pl.backward) {
        ctx.run_all([&](int tid) {
            for (int l = 1; l <= pl.L; ++l) {
                rc_build_gather_table(reinterpret_cast<GatherEntry*>(smem + pl.smTab + pl.lv[l].gatY),
                                      reinterpret_cast<const IdxLam*>(smem + pl.smTab + pl.lv[l].tabY), pl.lv[l].H,
                                      pl.lv[l - 1].H, pl.mode, tid, pl.T);
                rc_build_gather_table(reinterpret_cast<GatherEntry*>(smem + pl.smTab + pl.lv[l].gatX),
                                      reinterpret_cast<const IdxLam*>(smem + pl.smTab + pl.lv[l].tabX), pl.lv[l].W,
                                      pl.lv[l - 1].W, pl.mode, tid, pl.T);
            }
        });
        ctx.cta_sync();
    }
}

template <typename T, class Ctx>
RC_HD void rc_unpack_stage(Ctx& ctx, const Plan& pl, unsigned char* smem, int cg, bool from_g, int off_dst) {
    const int PAD = pl.K / 2;
    const unsigned magic_H = rc_magic(pl.H);
    ctx.run(pl.unit_lanes, [&](const ThreadPos& t) {
        const UnitIO io = rc_unit_io<T>(pl, smem, cg, t.u);
        if (io.nact == 0) return;
        float* planes = reinterpret_cast<float*>(smem + pl.smPlanes) + (long)t.u * pl.ppu * pl.plane_floats;
        rc_unpack_unit<T>(reinterpret_cast<const T*>(from_g ? io.sm_g : io.sm_x), io.nact, pl.H, pl.W, pl.vec, pl.magic_cpr, magic_H,
                          planes, pl.plane_floats, off_dst, pl.lv[0].pitch, PAD, t.ul, pl.unit_lanes);
    });
}

// Stage sequencing inside a unit: before a stage that uses the first n lanes, the first max(previous n, n) lanes
// synchronise, so warps that are not needed skip both the stage and its barriers.
struct StageSeq {
    int prev;
    template <class Ctx, class F> RC_HD void stage(Ctx& ctx, int n, F f) {
        ctx.sync(prev > n ? prev : n);
        ctx.run(n, f);
        prev = n;
    }
    template <class Ctx> RC_HD void join(Ctx& ctx, int n) {  // barrier only (before a unit-wide operation)
        ctx.sync(prev > n ? prev : n);
        prev = n;
    }
};

// The recompute shared by forward and backward: S_l for all levels (s_l = x_l + u_l), optionally keeping x_l.
template <int K, class Ctx>
RC_HD void rc_pyramid(Ctx& ctx, StageSeq& seq, const Plan& pl, unsigned char* smem, bool keep_x) {
    constexpr int PAD = K / 2;
    float* planes = reinterpret_cast<float*>(smem + pl.smPlanes);
    const float* wsm = reinterpret_cast<const float*>(smem + pl.smW);
    for (int l = 1; l <= pl.L; ++l) {  // model/recnext.py:27-29
        seq.stage(ctx, pl.lv[l].g2.lanes, [&](const ThreadPos& t) {
            if (!t.active) return;
            float* pb = planes + (long)t.p * pl.plane_floats;
            const LevelGeo& gi = pl.lv[l - 1];
            const LevelGeo& go = pl.lv[l];
            float* dst = pb + go.offS + PAD * go.pitch + PAD;
            float* dstx = (keep_x && go.offX >= 0) ? pb + go.offX + PAD * go.pitch + PAD : nullptr;
            const int pitch = go.pitch, Wo = go.W;
            rc_conv_s2<K>(pb + gi.offS, gi.pitch, wsm + (t.p * (pl.L + 2) + 0) * pl.wstride, pl.has_bias != 0, go.H, go.g2, t.lane, pl.g,
                          [&](int row, int c0, const float (&acc)[kStripW]) {
#pragma unroll
                              for (int c = 0; c < kStripW; ++c)
                                  if (c0 + c < Wo) {
                                      dst[row * pitch + c0 + c] = acc[c];
                                      if (dstx) dstx[row * pitch + c0 + c] = acc[c];
                                  }
                          });
        });
    }
    for (int l = pl.L; l >= 1; --l) {  // model/recnext.py:31-33 ; convs[L-l] acts on level l
        seq.stage(ctx, pl.lv[l].g1.lanes, [&](const ThreadPos& t) {
            if (!t.active) return;
            float* pb = planes + (long)t.p * pl.plane_floats;
            const LevelGeo& gl = pl.lv[l];
            float* T = pb + pl.offT;
            const int Hl = gl.H, Wl = gl.W, tp = gl.tpitch;
            rc_conv_s1<K, false>(pb + gl.offS, gl.pitch, wsm + (t.p * (pl.L + 2) + 1 + (pl.L - l)) * pl.wstride, pl.has_bias != 0,
                                 gl.H, gl.g1, t.lane, pl.g,
                                 [&](int row, int c0, const float (&acc)[kStripW]) { rc_store_T(T, tp, Hl, Wl, row, c0, acc); });
        });
        seq.stage(ctx, pl.lv[l].gu.lanes, [&](const ThreadPos& t) {
            if (!t.active) return;
            float* pb = planes + (long)t.p * pl.plane_floats;
            const LevelGeo& gl = pl.lv[l];
            const LevelGeo& gd = pl.lv[l - 1];
            if (gl.exact2x && pl.mode == 0)
                rc_upsample2x_add(pb + gd.offS, gd.pitch, PAD, gd.H, gd.W, pb + pl.offT, gl.tpitch, gl.H, gl.gu, t.lane, pl.g);
            else
                rc_upsample_add(pb + gd.offS, gd.pitch, PAD, gd.H, gd.W, pb + pl.offT, gl.tpitch, gl.H, gl.W,
                                reinterpret_cast<const IdxLam*>(smem + pl.smTab + gl.tabY),
                                reinterpret_cast<const IdxLam*>(smem + pl.smTab + gl.tabX), pl.mode, gl.gu, t.lane, pl.g);
        });
    }
}

// stores 4 results of one output row into the unit's raw output buffer (element type T)
template <typename T>
RC_HD void rc_store_raw4(T* __restrict__ dst, int W, int row, int c0, const float (&v)[kStripW]) {
    T* d = dst + row * W + c0;
    if (c0 + kStripW <= W && (W & 3) == 0) {
        alignas(16) T tmp[kStripW];
#pragma unroll
        for (int c = 0; c < kStripW; ++c) tmp[c] = Elem<T>::from_f(v[c]);
        if (sizeof(T) == 4) *reinterpret_cast<float4*>(d) = *reinterpret_cast<const float4*>(tmp);
        else *reinterpret_cast<float2*>(d) = *reinterpret_cast<const float2*>(tmp);
    } else {
#pragma unroll
        for (int c = 0; c < kStripW; ++c)
            if (c0 + c < W) d[c] = Elem<T>::from_f(v[c]);
    }
}

template <int K, typename T, class Ctx>
RC_HD void rc_forward_body(Ctx& ctx, const Plan& pl, const KernelArgs& a, unsigned char* smem, int cg, int chunk) {
    const int HW = pl.H * pl.W;
    float* planes = reinterpret_cast<float*>(smem + pl.smPlanes);
    const float* wsm = reinterpret_cast<const float*>(smem + pl.smW);
    const T* gx = reinterpret_cast<const T*>(a.x);
    T* gy = reinterpret_cast<T*>(a.out);

    rc_prologue(ctx, pl, a, smem, cg);
    const int first = chunk * pl.img_per_chunk;
    const int last = (first + pl.img_per_chunk) < pl.B ? (first + pl.img_per_chunk) : pl.B;
    if (first >= last) return;
    const long img_stride = (long)pl.C * HW;
    auto load = [&](int n) {
        ctx.load_begin([&](int u, void*& d0, const void*& s0, long& b0, void*& d1, const void*& s1, long& b1) {
            const UnitIO io = rc_unit_io<T>(pl, smem, cg, u);
            d0 = io.sm_x; s0 = gx + n * img_stride + io.goff; b0 = io.bytes; d1 = nullptr; s1 = nullptr; b1 = 0;
        });
    };
    load(first);
    StageSeq seq{pl.unit_lanes};
    for (int n = first; n < last; ++n) {
        ctx.store_drain();
        ctx.load_wait();
        rc_unpack_stage<T>(ctx, pl, smem, cg, false, pl.lv[0].offS);  // (the previous store synchronised the unit)
        if (n + 1 < last && !pl.share_raw) {
            seq.join(ctx, pl.unit_lanes);  // every lane has consumed the raw input
            load(n + 1);
        }
        rc_pyramid<K>(ctx, seq, pl, smem, false);
        seq.stage(ctx, pl.lv[0].g1.lanes, [&](const ThreadPos& t) {  // model/recnext.py:34
            if (!t.active) return;
            float* pb = planes + (long)t.p * pl.plane_floats;
            const LevelGeo& g0 = pl.lv[0];
            const UnitIO io = rc_unit_io<T>(pl, smem, cg, t.u);
            T* dst = reinterpret_cast<T*>(io.sm_out) + (long)(t.p - t.u * pl.ppu) * HW;
            const int W = pl.W;
            rc_conv_s1<K, false>(pb + g0.offS, g0.pitch, wsm + (t.p * (pl.L + 2) + 1 + pl.L) * pl.wstride, pl.has_bias != 0, g0.H, g0.g1,
                                 t.lane, pl.g, [&](int row, int c0, const float (&acc)[kStripW]) { rc_store_raw4<T>(dst, W, row, c0, acc); });
        });
        seq.prev = pl.unit_lanes;  // ctx.store synchronises the whole unit
        ctx.store([&](int u, void*& dst, const void*& src, long& bytes) {
            const UnitIO io = rc_unit_io<T>(pl, smem, cg, u);
            dst = gy + n * img_stride + io.goff; src = io.sm_out; bytes = io.bytes;
        });
        if (n + 1 < last && pl.share_raw) {  // raw out aliases raw in: the store must have read it first
            ctx.store_drain();
            load(n + 1);
        }
    }
    ctx.store_drain();
}

template <int K, typename T, class Ctx>
RC_HD void rc_backward_body(Ctx& ctx, const Plan& pl, const KernelArgs& a, unsigned char* smem, int cg, int chunk) {
    constexpr int PAD = K / 2;
    constexpr int NA = K * K + 1;
    const int nact = (pl.C - cg * pl.P) < pl.P ? (pl.C - cg * pl.P) : pl.P;
    const int HW = pl.H * pl.W;
    const int L = pl.L;
    float* planes = reinterpret_cast<float*>(smem + pl.smPlanes);
    const float* wsm = reinterpret_cast<const float*>(smem + pl.smW);
    float* wg = reinterpret_cast<float*>(smem + pl.smWG);
    const T* gx_in = reinterpret_cast<const T*>(a.x);
    const T* gg_in = reinterpret_cast<const T*>(a.gy);
    T* g_out = reinterpret_cast<T*>(a.out);
    const LevelGeo& g0 = pl.lv[0];

    rc_prologue(ctx, pl, a, smem, cg);
    const int first = chunk * pl.img_per_chunk;
    const int last = (first + pl.img_per_chunk) < pl.B ? (first + pl.img_per_chunk) : pl.B;
    const long img_stride = (long)pl.C * HW;
    auto slot_of = [&](const ThreadPos& t, int stage) {
        const int slot = pl.g >= 32 ? (t.tid >> 5) : t.p;
        return wg + ((long)slot * (L + 2) + stage) * pl.wstride;
    };
    auto load = [&](int n) {
        ctx.load_begin([&](int u, void*& d0, const void*& s0, long& b0, void*& d1, const void*& s1, long& b1) {
            const UnitIO io = rc_unit_io<T>(pl, smem, cg, u);
            d0 = io.sm_x; s0 = gx_in + n * img_stride + io.goff; b0 = io.bytes;
            d1 = io.sm_g; s1 = gg_in + n * img_stride + io.goff; b1 = io.bytes;
        });
    };
    if (first < last) load(first);
    StageSeq seq{pl.unit_lanes};
    for (int n = first; n < last; ++n) {
        ctx.store_drain();
        ctx.load_wait();
        rc_unpack_stage<T>(ctx, pl, smem, cg, false, g0.offS);  // (the previous store synchronised the unit)
        rc_unpack_stage<T>(ctx, pl, smem, cg, true, pl.offGY);
        rc_pyramid<K>(ctx, seq, pl, smem, true);

        // y = convs[L](s_0): weight gradient and input gradient
        seq.stage(ctx, g0.g1.lanes, [&](const ThreadPos& t) {
            float acc[NA];
#pragma unroll
            for (int i = 0; i < NA; ++i) acc[i] = 0.f;
            if (t.active) {
                float* pb = planes + (long)t.p * pl.plane_floats;
                rc_wgrad_s1<K>(pb + g0.offS, pb + pl.offGY, g0.pitch, g0.H, g0.g1, t.lane, pl.g, acc);
                float* G0 = pb + pl.offG0;
                const UnitIO io = rc_unit_io<T>(pl, smem, cg, t.u);
                T* dsto = reinterpret_cast<T*>(io.sm_out) + (long)(t.p - t.u * pl.ppu) * HW;
                const int W = pl.W, gp = pl.pitchG0;
                const bool direct = (L == 0);
                rc_conv_s1<K, true>(pb + pl.offGY, g0.pitch, wsm + (t.p * (L + 2) + 1 + L) * pl.wstride, false, g0.H, g0.g1, t.lane, pl.g,
                                    [&](int row, int c0, const float (&v)[kStripW]) {
                                        if (direct) { rc_store_raw4<T>(dsto, W, row, c0, v); return; }
#pragma unroll
                                        for (int c = 0; c < kStripW; ++c)
                                            if (c0 + c < W) G0[row * gp + c0 + c] = v[c];
                                    });
            }
            ctx.template wgrad_commit<NA>(t, pl, acc, slot_of(t, 1 + L));
        });

        for (int l = 1; l <= L; ++l) {
            if (l == 1) {  // S_0 is dead after the stage above: refill it with x_0 for the `down` filter gradient
                seq.join(ctx, pl.unit_lanes);
                rc_unpack_stage<T>(ctx, pl, smem, cg, false, g0.offS);
                if (n + 1 < last && !pl.share_raw) {  // both raw inputs are free from here on: prefetch the next image
                    seq.join(ctx, pl.unit_lanes);
                    load(n + 1);
                }
            }
            // GT_l = up^T(G_{l-1})
            seq.stage(ctx, pl.lv[l].gather_lanes, [&](const ThreadPos& t) {
                if (!t.active) return;
                float* pb = planes + (long)t.p * pl.plane_floats;
                const LevelGeo& gl = pl.lv[l];
                const LevelGeo& gd = pl.lv[l - 1];
                const float* gsrc = (l == 1) ? pb + pl.offG0 : pb + gd.offGS + PAD * gd.pitch + PAD;
                const int gpitch = (l == 1) ? pl.pitchG0 : gd.pitch;
                rc_upsample_bwd(pb + gl.offGT, gl.pitch, PAD, gl.H, gl.W, gsrc, gpitch,
                                reinterpret_cast<const GatherEntry*>(smem + pl.smTab + gl.gatY),
                                reinterpret_cast<const GatherEntry*>(smem + pl.smTab + gl.gatX), gl.magic_W, t.lane, pl.g);
            });
            seq.stage(ctx, pl.lv[l].g1.lanes, [&](const ThreadPos& t) {
                float acc[NA];
#pragma unroll
                for (int i = 0; i < NA; ++i) acc[i] = 0.f;
                if (t.active) {
                    float* pb = planes + (long)t.p * pl.plane_floats;
                    const LevelGeo& gl = pl.lv[l];
                    rc_wgrad_s1<K>(pb + gl.offS, pb + gl.offGT, gl.pitch, gl.H, gl.g1, t.lane, pl.g, acc);
                    float* dst = pb + gl.offGS + PAD * gl.pitch + PAD;
                    const int pitch = gl.pitch, Wl = gl.W;
                    rc_conv_s1<K, true>(pb + gl.offGT, gl.pitch, wsm + (t.p * (L + 2) + 1 + (L - l)) * pl.wstride, false, gl.H, gl.g1,
                                        t.lane, pl.g, [&](int row, int c0, const float (&v)[kStripW]) {
#pragma unroll
                                            for (int c = 0; c < kStripW; ++c)
                                                if (c0 + c < Wl) dst[row * pitch + c0 + c] = v[c];
                                        });
                }
                ctx.template wgrad_commit<NA>(t, pl, acc, slot_of(t, 1 + (L - l)));
            });
        }

        for (int l = L; l >= 1; --l) {
            // x_l = down(x_{l-1}): filter gradient (summed over levels) and input gradient added to G_{l-1}
            const int lanes = pl.lv[l].g2.lanes > pl.lv[l].gt.lanes ? pl.lv[l].g2.lanes : pl.lv[l].gt.lanes;
            seq.stage(ctx, lanes, [&](const ThreadPos& t) {
                float acc[NA];
#pragma unroll
                for (int i = 0; i < NA; ++i) acc[i] = 0.f;
                if (t.active) {
                    float* pb = planes + (long)t.p * pl.plane_floats;
                    const LevelGeo& gl = pl.lv[l];
                    const LevelGeo& gd = pl.lv[l - 1];
                    const float* X = (l - 1 == 0) ? pb + g0.offS : pb + gd.offX;
                    rc_wgrad_s2<K>(X, gd.pitch, pb + gl.offGS, gl.pitch, gl.H, gl.g2, t.lane, pl.g, acc);
                    if (l - 1 == 0) {
                        const float* G0 = pb + pl.offG0;
                        const UnitIO io = rc_unit_io<T>(pl, smem, cg, t.u);
                        T* dsto = reinterpret_cast<T*>(io.sm_out) + (long)(t.p - t.u * pl.ppu) * HW;
                        const int W = pl.W, gp = pl.pitchG0;
                        rc_convT_s2<K>(pb + gl.offGS, gl.pitch, wsm + (t.p * (L + 2) + 0) * pl.wstride, gd.H, gd.W, gl.gt, t.lane, pl.g,
                                       [&](int i, int j, float v) { dsto[i * W + j] = Elem<T>::from_f(G0[i * gp + j] + v); });
                    } else {
                        float* dst = pb + gd.offGS + PAD * gd.pitch + PAD;
                        const int pitch = gd.pitch;
                        rc_convT_s2<K>(pb + gl.offGS, gl.pitch, wsm + (t.p * (L + 2) + 0) * pl.wstride, gd.H, gd.W, gl.gt, t.lane, pl.g,
                                       [&](int i, int j, float v) { dst[i * pitch + j] += v; });
                    }
                }
                ctx.template wgrad_commit<NA>(t, pl, acc, slot_of(t, 0));
            });
        }
        seq.prev = pl.unit_lanes;  // ctx.store synchronises the whole unit
        ctx.store([&](int u, void*& dst, const void*& src, long& bytes) {
            const UnitIO io = rc_unit_io<T>(pl, smem, cg, u);
            dst = g_out + n * img_stride + io.goff; src = io.sm_out; bytes = io.bytes;
        });
        if ((L == 0 || pl.share_raw) && n + 1 < last) {
            if (pl.share_raw) ctx.store_drain();
            load(n + 1);
        }
    }
    ctx.store_drain();

    // per-CTA partials -> workspace [chunk][(L+2)][C][wstride]; slots of one plane are summed in fixed order
    ctx.cta_sync();
    ctx.run_all([&](int tid) {
        const int per_plane = (L + 2) * pl.wstride;
        const int spp = pl.g >= 32 ? pl.g / 32 : 1;  // slots per plane
        for (int i = tid; i < nact * per_plane; i += pl.T) {
            const int p = i / per_plane, r = i - p * per_plane;
            const int stage = r / pl.wstride, e = r - stage * pl.wstride;
            float s = 0.f;
            for (int q = 0; q < spp; ++q) s += wg[((long)(p * spp + q) * (L + 2) + stage) * pl.wstride + e];
            const long ch = (long)cg * pl.P + p;
            a.partial[(((long)chunk * (L + 2) + stage) * pl.C + ch) * pl.wstride + e] = s;
        }
    });
}

}  // namespace recnext
