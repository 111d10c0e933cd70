// tests/emu/wemu.cu — TEST INFRASTRUCTURE ONLY.
//
// Runs the team schedules of recnext_b200/csrc/wbody.cuh on the CPU: every lane of a team is a loop iteration,
// team barriers are the ends of those loops, TMA bulk copies are memcpy.  Lets `pytest -m "not gpu"` check the
// tiling, index arithmetic and gradient chain of the CUDA source against the oracle without a GPU.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/recnext_b200.h"
#include "../../recnext_b200/csrc/wbody.cuh"

using namespace recnext;

struct HostWCtx {
    int team_lanes;
    template <class F> void stage(F f) { for (int tl = 0; tl < team_lanes; ++tl) f(tl); }
    void load(void* dst, const void* src, int bytes, int) { memcpy(dst, src, bytes); }
    void wait(int) {}
    template <int N> void reduce(const WPlan& pl, int tl, float (&acc)[N], float* slot) {
        (void)pl; (void)tl;
        for (int i = 0; i < N; ++i) slot[i] += acc[i];
    }
};

template <int K, typename T>
static void run_all(const WPlan& pl, const KernelArgs& a, bool bwd) {
    std::vector<unsigned char> smem_raw(pl.smem_bytes + 256);
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw.data() + 127) & ~(uintptr_t)127);
    for (int blk = 0; blk < pl.grid; ++blk) {
        memset(smem, 0xff, pl.smem_bytes);  // poison: reads of never-written cells show up as NaN
        for (int t = 0; t < pl.threads; ++t) w_cta_init(pl, smem, t, pl.threads);
        if (bwd) for (int t = 0; t < pl.threads; ++t) w_cta_init_bwd(pl, smem, t, pl.threads);
        for (int team = 0; team < pl.NT; ++team) {
            HostWCtx ctx{pl.team_lanes};
            if (bwd) w_backward_team<K, T>(ctx, pl, a, smem, team, blk * pl.NT + team);
            else
            w_forward_team<K, T>(ctx, pl, a, smem, team, blk * pl.NT + team);
        }
    }
}

template <typename T>
static void run_k(const WPlan& pl, const KernelArgs& a, bool bwd) {
    switch (pl.K) {
        case 3: run_all<3, T>(pl, a, bwd); break;
        case 5: run_all<5, T>(pl, a, bwd); break;
        case 7: run_all<7, T>(pl, a, bwd); break;
    }
}

extern "C" {

// opts: {force_G, force_TW, force_NT, force_no_tma, num_sms}; pointers are HOST pointers here.
__attribute__((visibility("default")))
int wemu_recconv(const recconv_desc* d, const recconv_params* p, const void* x, const void* gy, void* out, float* gw, float* gb,
                 int backward, const int* opts, int* plan_out /* G,TW,NT,grid,smem,use_tma,tpc,LPP */) {
    WPlanOptions opt;
    if (opts) { opt.force_G = opts[0]; opt.force_TW = opts[1]; opt.force_NT = opts[2]; opt.force_no_tma = opts[3]; if (opts[4]) opt.num_sms = opts[4]; }
    WPlan pl;
    int rc = w_make_plan(pl, d->B, d->C, d->H, d->W, d->k, d->level, d->mode, d->dtype, d->wdtype, d->has_bias, backward, opt);
    if (rc) return -rc;
    if (plan_out) { plan_out[0] = pl.G; plan_out[1] = pl.TW; plan_out[2] = pl.NT; plan_out[3] = pl.grid; plan_out[4] = pl.smem_bytes; plan_out[5] = pl.use_tma; plan_out[6] = pl.tpc; plan_out[7] = pl.LPP; }
    KernelArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.gy = gy; a.out = out;
    a.w[0] = p->w_down; a.b[0] = d->has_bias ? p->b_down : nullptr;
    for (int j = 0; j <= d->level; ++j) { a.w[1 + j] = p->w_convs[j]; a.b[1 + j] = d->has_bias ? p->b_convs[j] : nullptr; }
    std::vector<float> partial((size_t)pl.ws_partial_floats + 1, 0.f);
    a.partial = partial.data();
    switch (d->dtype) {
        case RECNEXT_F32: run_k<float>(pl, a, backward != 0); break;
        case RECNEXT_BF16: run_k<__nv_bfloat16>(pl, a, backward != 0); break;
        case RECNEXT_F16: run_k<__half>(pl, a, backward != 0); break;
        default: return -2;
    }
    if (backward) {
        const int KK = d->k * d->k, ws = pl.wstride;
        const long total = (long)(d->level + 2) * d->C * ws;
        for (long i = 0; i < total; ++i) {
            float s = 0.f;
            for (int ch = 0; ch < pl.tpc; ++ch) s += partial[(size_t)ch * total + i];
            const long sc = i / ws; const int e = (int)(i - sc * ws);
            if (e < KK) gw[sc * KK + e] = s; else if (e == KK && gb) gb[sc] = s;
        }
    }
    return 0;
}
}
