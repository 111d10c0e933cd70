// tests/emu/recconv_emu.cu — TEST INFRASTRUCTURE ONLY.
//
// Runs the stage schedules of recnext_b200/csrc/recconv_body.cuh on the CPU by replacing the CUDA execution
// context with a sequential one (every "thread" of a CTA is a loop iteration; barriers are no-ops because
// stages never communicate inside a stage).  This lets `pytest -m "not gpu"` check the kernels' index
// arithmetic, tiling and gradient chain against the oracle without a GPU.  It is NOT part of the product:
// recnext_b200 never loads it, and it is deliberately slow.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/recnext_b200.h"
#include "../../recnext_b200/csrc/recconv_body.cuh"

using namespace recnext;

struct HostCtx {
    int T, n_units, cg;
    const Plan* pl;
    template <class F> __host__ __device__ void run_all(F f) { for (int t = 0; t < T; ++t) f(t); }
    template <class F> __host__ __device__ void run(int n, F f) {
        for (int t = 0; t < T; ++t) {
            const ThreadPos pos = rc_thread_pos(*pl, t, cg);
            if (pos.ul < n) f(pos);
        }
    }
    __host__ __device__ void sync(int) {}
    __host__ __device__ void cta_sync() {}
    template <class F> __host__ __device__ void load_begin(F desc) {
        for (int u = 0; u < n_units; ++u) {
            void *d0, *d1; const void *s0, *s1; long b0, b1;
            desc(u, d0, s0, b0, d1, s1, b1);
            if (b0) memcpy(d0, s0, b0);
            if (b1) memcpy(d1, s1, b1);
        }
    }
    __host__ __device__ void load_wait() {}
    template <class F> __host__ __device__ void store(F desc) {
        for (int u = 0; u < n_units; ++u) {
            void* dst; const void* src; long bytes;
            desc(u, dst, src, bytes);
            if (bytes) memcpy(dst, src, bytes);
        }
    }
    __host__ __device__ void store_drain() {}
    template <int N> __host__ __device__ void wgrad_commit(const ThreadPos& t, const Plan& pl, float (&acc)[N], float* slot) {
        if (t.p < pl.P) for (int i = 0; i < N; ++i) slot[i] += acc[i];
    }
};

template <int K, typename T>
static void run_all(const Plan& pl, const KernelArgs& a, bool bwd) {
    std::vector<unsigned char> smem_raw(pl.smem_bytes + 256);
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw.data() + 127) & ~(uintptr_t)127);
    for (int blk = 0; blk < pl.n_cg * pl.n_chunk; ++blk) {
        // poison shared memory so that reads of never-written (non-zeroed) cells show up as NaN
        memset(smem, 0xff, pl.smem_bytes);
        const int cg = blk % pl.n_cg, chunk = blk / pl.n_cg;
        HostCtx ctx{pl.T, pl.n_units, cg, &pl};
        if (bwd) rc_backward_body<K, T>(ctx, pl, a, smem, cg, chunk);
        else rc_forward_body<K, T>(ctx, pl, a, smem, cg, chunk);
    }
}

template <typename T>
static void run_k(const Plan& pl, const KernelArgs& a, bool bwd) {
    switch (pl.K) {
        case 3: run_all<3, T>(pl, a, bwd); break;
        case 5: run_all<5, T>(pl, a, bwd); break;
        case 7: run_all<7, T>(pl, a, bwd); break;
    }
}

extern "C" {

// opts: {force_P, force_g, force_chunks, force_no_tma}; pointers are HOST pointers here.
__attribute__((visibility("default")))
int emu_recconv(const recconv_desc* d, const recconv_params* p, const void* x, const void* gy, void* out, float* gw, float* gb,
                int backward, const int* opts, int* plan_out /* P,g,T,n_cg,n_chunk,smem,use_tma */) {
    PlanOptions opt;
    if (opts) { opt.force_P = opts[0]; opt.force_g = opts[1]; opt.force_chunks = opts[2]; opt.force_no_tma = opts[3]; }
    Plan pl;
    int rc = rc_make_plan(pl, d->B, d->C, d->H, d->W, d->k, d->level, d->mode, d->dtype, d->wdtype, d->has_bias, backward, opt);
    if (rc) { fprintf(stderr, "emu plan rc=%d B=%d C=%d H=%d W=%d k=%d L=%d T=%d P=%d g=%d smem=%d\n", rc, d->B, d->C, d->H, d->W, d->k, d->level, pl.T, pl.P, pl.g, pl.smem_bytes); return -rc; }
    if (plan_out) { plan_out[0] = pl.P; plan_out[1] = pl.g; plan_out[2] = pl.T; plan_out[3] = pl.n_cg; plan_out[4] = pl.n_chunk; plan_out[5] = pl.smem_bytes; plan_out[6] = pl.use_tma; }
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
            for (int ch = 0; ch < pl.n_chunk; ++ch) s += partial[(size_t)ch * total + i];
            const long sc = i / ws; const int e = (int)(i - sc * ws);
            if (e < KK) gw[sc * KK + e] = s; else if (e == KK && gb) gb[sc] = s;
        }
    }
    return 0;
}
}
