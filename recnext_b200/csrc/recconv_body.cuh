// recconv_body.cuh — stage schedules of the fused RecConv forward / backward kernels.
//
// The schedule is a template over an execution context `Ctx`:
//   ctx.run(f)            f(tid) for every thread of the CTA (device: this thread; host emulation: a loop)
//   ctx.sync()            CTA barrier
//   ctx.load_begin / load_wait / store / store_drain   movement of a raw plane group between global and shared
//   ctx.wgrad_commit      reduction of per-lane weight-gradient partials into the CTA's accumulation slots
// so the identical stage code runs under CUDA (recconv_kernels.cu) and on the CPU for logic tests (tests/emu).
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
    float* partial;   // bwd: [n_chunk][(L+2)][C][K*K+1]
    const void* w[kMaxLevel + 2];  // slot 0 = down, 1+j = convs[j]
    const void* b[kMaxLevel + 2];  // may be null
};

struct ThreadPos {
    int tid, p, lane, c;  // thread, plane slot, lane within plane, channel
    bool active;          // owns a real plane
};

RC_HD ThreadPos rc_thread_pos(const Plan& pl, int tid, int cg) {
    ThreadPos t;
    t.tid = tid; t.p = tid / pl.g; t.lane = tid - t.p * pl.g; t.c = cg * pl.P + t.p;
    t.active = t.p < pl.P && t.c < pl.C;
    return t;
}

template <class Ctx>
RC_HD void rc_prologue(Ctx& ctx, const Plan& pl, const KernelArgs& a, unsigned char* smem, int cg) {
    const int nact = (pl.C - cg * pl.P) < pl.P ? (pl.C - cg * pl.P) : pl.P;
    // zero every plane block (borders must stay zero for the whole kernel) and the accumulation slots
    ctx.run([&](int tid) {
        float4* z = reinterpret_cast<float4*>(smem + pl.smPlanes);
        const int n4 = pl.P * pl.plane_floats / 4;
        const float4 zero = {0.f, 0.f, 0.f, 0.f};
        for (int i = tid; i < n4; i += pl.T) z[i] = zero;
        float* wg = reinterpret_cast<float*>(smem + pl.smWG);
        const int nwg = pl.nslots * (pl.L + 2) * pl.wstride;
        for (int i = tid; i < nwg; i += pl.T) wg[i] = 0.f;
        // filters of this channel group -> shared (fp32), bias in the last element of each slot
        float* wsm = reinterpret_cast<float*>(smem + pl.smW);
        const int KK = pl.K * pl.K;
        const int per_plane = (pl.L + 2) * pl.wstride;
        for (int i = tid; i < nact * per_plane; i += pl.T) {
            const int p = i / per_plane, r = i - p * per_plane;
            const int slot = r / pl.wstride, e = r - slot * pl.wstride;
            const long ch = (long)cg * pl.P + p;
            float v = 0.f;
            if (slot == 0 && pl.L == 0) v = 0.f;  // `down` exists in the state_dict but is unused at level 0
            else if (e < KK) v = rc_load_param(a.w[slot], pl.wdtype, ch * KK + e);
            else if (pl.has_bias && a.b[slot]) v = rc_load_param(a.b[slot], pl.wdtype, ch);
            wsm[i] = v;
        }
        // forward interpolation tables
        for (int l = 1; l <= pl.L; ++l) {
            rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + pl.smTab + pl.lv[l].tabY), pl.lv[l].H, pl.lv[l - 1].H,
                               pl.mode, tid, pl.T);
            rc_build_fwd_table(reinterpret_cast<IdxLam*>(smem + pl.smTab + pl.lv[l].tabX), pl.lv[l].W, pl.lv[l - 1].W,
                               pl.mode, tid, pl.T);
        }
    });
    ctx.sync();
    if (pl.backward) {
        ctx.run([&](int tid) {
            for (int l = 1; l <= pl.L; ++l) {
                rc_build_range_table(reinterpret_cast<Range*>(smem + pl.smTab + pl.lv[l].rngY),
                                     reinterpret_cast<const IdxLam*>(smem + pl.smTab + pl.lv[l].tabY), pl.lv[l].H,
                                     pl.lv[l - 1].H, pl.mode, tid, pl.T);
                rc_build_range_table(reinterpret_cast<Range*>(smem + pl.smTab + pl.lv[l].rngX),
                                     reinterpret_cast<const IdxLam*>(smem + pl.smTab + pl.lv[l].tabX), pl.lv[l].W,
                                     pl.lv[l - 1].W, pl.mode, tid, pl.T);
            }
        });
        ctx.sync();
    }
}

// The recompute shared by forward and backward: S_l for all levels (s_l = x_l + u_l), optionally keeping x_l.
template <int K, class Ctx>
RC_HD void rc_pyramid(Ctx& ctx, const Plan& pl, unsigned char* smem, int cg, bool keep_x) {
    constexpr int PAD = K / 2;
    float* planes = reinterpret_cast<float*>(smem + pl.smPlanes);
    const float* wsm = reinterpret_cast<const float*>(smem + pl.smW);
    for (int l = 1; l <= pl.L; ++l) {  // model/recnext.py:27-29
        ctx.run([&](int tid) {
            const ThreadPos t = rc_thread_pos(pl, tid, cg);
            if (!t.active) return;
            float* pb = planes + (long)t.p * pl.plane_floats;
            const LevelGeo& gi = pl.lv[l - 1];
            const LevelGeo& go = pl.lv[l];
            float* dst = pb + go.offS + PAD * go.pitch + PAD;
            float* dstx = (keep_x && go.offX >= 0) ? pb + go.offX + PAD * go.pitch + PAD : nullptr;
            const int pitch = go.pitch, Wo = go.W;
            rc_conv_s2<K>(pb + gi.offS, gi.pitch, wsm + (t.p * (pl.L + 2) + 0) * pl.wstride, pl.has_bias != 0, go.H,
                          go.W, go.rpi_down, t.lane, pl.g, [&](int row, int c0, const float (&acc)[kStripW]) {
#pragma unroll
                              for (int c = 0; c < kStripW; ++c)
                                  if (c0 + c < Wo) {
                                      dst[row * pitch + c0 + c] = acc[c];
                                      if (dstx) dstx[row * pitch + c0 + c] = acc[c];
                                  }
                          });
        });
        ctx.sync();
    }
    for (int l = pl.L; l >= 1; --l) {  // model/recnext.py:31-33 ; convs[L-l] acts on level l
        ctx.run([&](int tid) {
            const ThreadPos t = rc_thread_pos(pl, tid, cg);
            if (!t.active) return;
            float* pb = planes + (long)t.p * pl.plane_floats;
            const LevelGeo& gl = pl.lv[l];
            float* T = pb + pl.offT;
            const int Wl = gl.W;
            rc_conv_s1<K, false>(pb + gl.offS, gl.pitch, wsm + (t.p * (pl.L + 2) + 1 + (pl.L - l)) * pl.wstride,
                                 pl.has_bias != 0, gl.H, gl.W, gl.rpi, t.lane, pl.g,
                                 [&](int row, int c0, const float (&acc)[kStripW]) {
#pragma unroll
                                     for (int c = 0; c < kStripW; ++c)
                                         if (c0 + c < Wl) T[row * Wl + c0 + c] = acc[c];
                                 });
        });
        ctx.sync();
        ctx.run([&](int tid) {
            const ThreadPos t = rc_thread_pos(pl, tid, cg);
            if (!t.active) return;
            float* pb = planes + (long)t.p * pl.plane_floats;
            const LevelGeo& gl = pl.lv[l];
            const LevelGeo& gd = pl.lv[l - 1];
            rc_upsample_add(pb + gd.offS, gd.pitch, PAD, gd.H, gd.W, pb + pl.offT, gl.H, gl.W,
                            reinterpret_cast<const IdxLam*>(smem + pl.smTab + gl.tabY),
                            reinterpret_cast<const IdxLam*>(smem + pl.smTab + gl.tabX), pl.mode, t.lane, pl.g);
        });
        ctx.sync();
    }
}

template <int K, typename T, class Ctx>
RC_HD void rc_forward_body(Ctx& ctx, const Plan& pl, const KernelArgs& a, unsigned char* smem, int cg, int chunk) {
    constexpr int PAD = K / 2;
    const int nact = (pl.C - cg * pl.P) < pl.P ? (pl.C - cg * pl.P) : pl.P;
    const int HW = pl.H * pl.W;
    const long group_bytes = (long)nact * HW * sizeof(T);
    float* planes = reinterpret_cast<float*>(smem + pl.smPlanes);
    const float* wsm = reinterpret_cast<const float*>(smem + pl.smW);
    T* rawx = reinterpret_cast<T*>(smem + pl.smRawX);
    T* rawo = reinterpret_cast<T*>(smem + pl.smRawOut);
    const T* gx = reinterpret_cast<const T*>(a.x);
    T* gy = reinterpret_cast<T*>(a.out);

    rc_prologue(ctx, pl, a, smem, cg);
    const int first = chunk * pl.img_per_chunk;
    const int last = (first + pl.img_per_chunk) < pl.B ? (first + pl.img_per_chunk) : pl.B;
    if (first >= last) return;
    auto goff = [&](int n) { return ((long)n * pl.C + (long)cg * pl.P) * HW; };
    ctx.load_begin(rawx, gx + goff(first), group_bytes, nullptr, nullptr, 0);
    for (int n = first; n < last; ++n) {
        ctx.store_drain();
        ctx.load_wait();
        ctx.run([&](int tid) {
            rc_unpack_group<T>(rawx, nact, pl.H, pl.W, planes, pl.plane_floats, pl.lv[0].offS, pl.lv[0].pitch, PAD, tid, pl.T);
        });
        ctx.sync();
        if (n + 1 < last && !pl.share_raw) ctx.load_begin(rawx, gx + goff(n + 1), group_bytes, nullptr, nullptr, 0);
        rc_pyramid<K>(ctx, pl, smem, cg, false);
        ctx.run([&](int tid) {  // model/recnext.py:34
            const ThreadPos t = rc_thread_pos(pl, tid, cg);
            if (!t.active) return;
            float* pb = planes + (long)t.p * pl.plane_floats;
            const LevelGeo& g0 = pl.lv[0];
            T* dst = rawo + (long)t.p * HW;
            const int W = pl.W;
            rc_conv_s1<K, false>(pb + g0.offS, g0.pitch, wsm + (t.p * (pl.L + 2) + 1 + pl.L) * pl.wstride, pl.has_bias != 0,
                                 g0.H, g0.W, g0.rpi, t.lane, pl.g, [&](int row, int c0, const float (&acc)[kStripW]) {
#pragma unroll
                                     for (int c = 0; c < kStripW; ++c)
                                         if (c0 + c < W) dst[row * W + c0 + c] = Elem<T>::from_f(acc[c]);
                                 });
        });
        ctx.store(gy + goff(n), rawo, group_bytes);
        if (n + 1 < last && pl.share_raw) {  // raw out aliases raw in: the store must have read it first
            ctx.store_drain();
            ctx.load_begin(rawx, gx + goff(n + 1), group_bytes, nullptr, nullptr, 0);
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
    const long group_bytes = (long)nact * HW * sizeof(T);
    float* planes = reinterpret_cast<float*>(smem + pl.smPlanes);
    const float* wsm = reinterpret_cast<const float*>(smem + pl.smW);
    float* wg = reinterpret_cast<float*>(smem + pl.smWG);
    T* rawx = reinterpret_cast<T*>(smem + pl.smRawX);
    T* rawg = reinterpret_cast<T*>(smem + pl.smRawG);
    T* rawo = reinterpret_cast<T*>(smem + pl.smRawOut);
    const T* gx_in = reinterpret_cast<const T*>(a.x);
    const T* gg_in = reinterpret_cast<const T*>(a.gy);
    T* g_out = reinterpret_cast<T*>(a.out);
    const LevelGeo& g0 = pl.lv[0];

    rc_prologue(ctx, pl, a, smem, cg);
    const int first = chunk * pl.img_per_chunk;
    const int last = (first + pl.img_per_chunk) < pl.B ? (first + pl.img_per_chunk) : pl.B;
    auto goff = [&](int n) { return ((long)n * pl.C + (long)cg * pl.P) * HW; };
    auto slot_of = [&](const ThreadPos& t, int stage) {
        const int slot = pl.g >= 32 ? (t.tid >> 5) : t.p;
        return wg + ((long)slot * (L + 2) + stage) * pl.wstride;
    };
    if (first < last) ctx.load_begin(rawx, gx_in + goff(first), group_bytes, rawg, gg_in + goff(first), group_bytes);
    for (int n = first; n < last; ++n) {
        ctx.store_drain();
        ctx.load_wait();
        ctx.run([&](int tid) {
            rc_unpack_group<T>(rawx, nact, pl.H, pl.W, planes, pl.plane_floats, g0.offS, g0.pitch, PAD, tid, pl.T);
            rc_unpack_group<T>(rawg, nact, pl.H, pl.W, planes, pl.plane_floats, pl.offGY, g0.pitch, PAD, tid, pl.T);
        });
        ctx.sync();
        rc_pyramid<K>(ctx, pl, smem, cg, true);

        // y = convs[L](s_0): weight gradient and input gradient
        ctx.run([&](int tid) {
            const ThreadPos t = rc_thread_pos(pl, tid, cg);
            float acc[NA];
#pragma unroll
            for (int i = 0; i < NA; ++i) acc[i] = 0.f;
            if (t.active) {
                float* pb = planes + (long)t.p * pl.plane_floats;
                rc_wgrad_s1<K>(pb + g0.offS, pb + pl.offGY, g0.pitch, g0.H, g0.W, g0.rpi, t.lane, pl.g, acc);
                float* G0 = pb + pl.offG0;
                T* dsto = rawo + (long)t.p * HW;
                const int W = pl.W, gp = pl.pitchG0;
                const bool direct = (L == 0);
                rc_conv_s1<K, true>(pb + pl.offGY, g0.pitch, wsm + (t.p * (L + 2) + 1 + L) * pl.wstride, false, g0.H, g0.W,
                                    g0.rpi, t.lane, pl.g, [&](int row, int c0, const float (&v)[kStripW]) {
#pragma unroll
                                        for (int c = 0; c < kStripW; ++c)
                                            if (c0 + c < W) {
                                                if (direct) dsto[row * W + c0 + c] = Elem<T>::from_f(v[c]);
                                                else G0[row * gp + c0 + c] = v[c];
                                            }
                                    });
            }
            ctx.template wgrad_commit<NA>(t, pl, acc, slot_of(t, 1 + L));
        });
        ctx.sync();

        for (int l = 1; l <= L; ++l) {
            // GT_l = up^T(G_{l-1}); S_0 is dead after the stage above, so level 1 also refills it with x_0
            ctx.run([&](int tid) {
                if (l == 1)
                    rc_unpack_group<T>(rawx, nact, pl.H, pl.W, planes, pl.plane_floats, g0.offS, g0.pitch, PAD, tid, pl.T);
                const ThreadPos t = rc_thread_pos(pl, tid, cg);
                if (!t.active) return;
                float* pb = planes + (long)t.p * pl.plane_floats;
                const LevelGeo& gl = pl.lv[l];
                const LevelGeo& gd = pl.lv[l - 1];
                const float* gsrc = (l == 1) ? pb + pl.offG0 : pb + gd.offGS + PAD * gd.pitch + PAD;
                const int gpitch = (l == 1) ? pl.pitchG0 : gd.pitch;
                rc_upsample_bwd(pb + gl.offGT, gl.pitch, PAD, gl.H, gl.W, gsrc, gpitch,
                                reinterpret_cast<const IdxLam*>(smem + pl.smTab + gl.tabY),
                                reinterpret_cast<const IdxLam*>(smem + pl.smTab + gl.tabX),
                                reinterpret_cast<const Range*>(smem + pl.smTab + gl.rngY),
                                reinterpret_cast<const Range*>(smem + pl.smTab + gl.rngX), pl.mode, t.lane, pl.g);
            });
            ctx.sync();
            if (l == 1 && n + 1 < last && !pl.share_raw)  // both raw input buffers are free from here on: prefetch the next image
                ctx.load_begin(rawx, gx_in + goff(n + 1), group_bytes, rawg, gg_in + goff(n + 1), group_bytes);
            ctx.run([&](int tid) {
                const ThreadPos t = rc_thread_pos(pl, tid, cg);
                float acc[NA];
#pragma unroll
                for (int i = 0; i < NA; ++i) acc[i] = 0.f;
                if (t.active) {
                    float* pb = planes + (long)t.p * pl.plane_floats;
                    const LevelGeo& gl = pl.lv[l];
                    rc_wgrad_s1<K>(pb + gl.offS, pb + gl.offGT, gl.pitch, gl.H, gl.W, gl.rpi, t.lane, pl.g, acc);
                    float* dst = pb + gl.offGS + PAD * gl.pitch + PAD;
                    const int pitch = gl.pitch, Wl = gl.W;
                    rc_conv_s1<K, true>(pb + gl.offGT, gl.pitch, wsm + (t.p * (L + 2) + 1 + (L - l)) * pl.wstride, false,
                                        gl.H, gl.W, gl.rpi, t.lane, pl.g, [&](int row, int c0, const float (&v)[kStripW]) {
#pragma unroll
                                            for (int c = 0; c < kStripW; ++c)
                                                if (c0 + c < Wl) dst[row * pitch + c0 + c] = v[c];
                                        });
                }
                ctx.template wgrad_commit<NA>(t, pl, acc, slot_of(t, 1 + (L - l)));
            });
            ctx.sync();
        }

        for (int l = L; l >= 1; --l) {
            // x_l = down(x_{l-1}): filter gradient (summed over levels) and input gradient added to G_{l-1}
            ctx.run([&](int tid) {
                const ThreadPos t = rc_thread_pos(pl, tid, cg);
                float acc[NA];
#pragma unroll
                for (int i = 0; i < NA; ++i) acc[i] = 0.f;
                if (t.active) {
                    float* pb = planes + (long)t.p * pl.plane_floats;
                    const LevelGeo& gl = pl.lv[l];
                    const LevelGeo& gd = pl.lv[l - 1];
                    const float* X = (l - 1 == 0) ? pb + g0.offS : pb + gd.offX;
                    rc_wgrad_s2<K>(X, gd.pitch, pb + gl.offGS, gl.pitch, gl.H, gl.W, gl.rpi_down, t.lane, pl.g, acc);
                    if (l - 1 == 0) {
                        const float* G0 = pb + pl.offG0;
                        T* dsto = rawo + (long)t.p * HW;
                        const int W = pl.W, gp = pl.pitchG0;
                        rc_convT_s2<K>(pb + gl.offGS, gl.pitch, wsm + (t.p * (L + 2) + 0) * pl.wstride, gd.H, gd.W, t.lane,
                                       pl.g, [&](int i, int j, float v) { dsto[i * W + j] = Elem<T>::from_f(G0[i * gp + j] + v); });
                    } else {
                        float* dst = pb + gd.offGS + PAD * gd.pitch + PAD;
                        const int pitch = gd.pitch;
                        rc_convT_s2<K>(pb + gl.offGS, gl.pitch, wsm + (t.p * (L + 2) + 0) * pl.wstride, gd.H, gd.W, t.lane,
                                       pl.g, [&](int i, int j, float v) { dst[i * pitch + j] += v; });
                    }
                }
                ctx.template wgrad_commit<NA>(t, pl, acc, slot_of(t, 0));
            });
            if (l > 1) ctx.sync();
        }
        ctx.store(g_out + goff(n), rawo, group_bytes);
        if ((L == 0 || pl.share_raw) && n + 1 < last) {
            if (pl.share_raw) ctx.store_drain();
            ctx.load_begin(rawx, gx_in + goff(n + 1), group_bytes, rawg, gg_in + goff(n + 1), group_bytes);
        }
    }
    ctx.store_drain();

    // per-CTA partials -> workspace [chunk][(L+2)][C][K*K+1]; slots of one plane are summed in fixed order
    ctx.sync();
    ctx.run([&](int tid) {
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
