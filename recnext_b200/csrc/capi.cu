// capi.cu — extern "C" entry points of librecnext_b200.so (see include/recnext_b200.h).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/recnext_b200.h"
#include "recconv_body.cuh"
#include "wplan.h"
#include "mplan.h"
#include "mbplan.h"
#include "gstream.h"
#include "ffn_tc.h"
#include "stem.h"
#include "devcfg.h"
#include <stdlib.h>

namespace recnext {
cudaError_t m_launch(const MPlan&, const KernelArgs&, cudaStream_t);  // recconv_m5.cu: tensor-core forward
cudaError_t mb_launch(const MBPlan&, const KernelArgs&, float* gw, float* gb, cudaStream_t);  // recconv_mb5.cu: tensor-core backward
bool m_static_geometry(const MPlan&);
struct FfnPlan { int B, C, HID, HW, NTN, tiles, chunkB, PB, offX, offH, smem_bytes, dtype, NQ, staged, offW, kc; };  // ffn_mma.cu
int ffn_make_plan(FfnPlan&, int B, int C, int HID, int HW, int dtype);
int linattn_launch(int B, int dim, int heads, int n, int dtype, const void* q, const void* k, const float* qb, const float* kb, const void* v, const void* pe,
                   const float* pew, const float* peb, int pw, void* out, cudaStream_t stream, cudaError_t* err);
int dwdown_launch(int B, int C, int H, int W, int dtype, const void* x, const float* w, const float* b, void* out, cudaStream_t stream, cudaError_t* err);
cudaError_t ffn_launch(const FfnPlan&, const void* y, const void* x, const void* w1, const float* b1, const void* w2, const float* b2, void* out,
                       cudaStream_t stream);
template <int K, typename T, bool BWD> cudaError_t w_launch(const WPlan&, const KernelArgs&, cudaStream_t);
typedef cudaError_t (*w_launch_fn)(const WPlan&, const KernelArgs&, cudaStream_t);
#define W_DECLARE_K(K)                                                                                           \
    extern template cudaError_t w_launch<K, float, false>(const WPlan&, const KernelArgs&, cudaStream_t);          \
    extern template cudaError_t w_launch<K, __nv_bfloat16, false>(const WPlan&, const KernelArgs&, cudaStream_t);  \
    extern template cudaError_t w_launch<K, __half, false>(const WPlan&, const KernelArgs&, cudaStream_t);         \
    extern template cudaError_t w_launch<K, float, true>(const WPlan&, const KernelArgs&, cudaStream_t);           \
    extern template cudaError_t w_launch<K, __nv_bfloat16, true>(const WPlan&, const KernelArgs&, cudaStream_t);   \
    extern template cudaError_t w_launch<K, __half, true>(const WPlan&, const KernelArgs&, cudaStream_t);
W_DECLARE_K(3)
W_DECLARE_K(5)
W_DECLARE_K(7)
typedef cudaError_t (*rc_launch_fn)(const Plan&, const KernelArgs&, cudaStream_t);
template <int K, typename T, bool BWD> cudaError_t rc_launch(const Plan&, const KernelArgs&, cudaStream_t);

#define RC_DECLARE_K(K)                                                                                       \
    extern template cudaError_t rc_launch<K, float, false>(const Plan&, const KernelArgs&, cudaStream_t);     \
    extern template cudaError_t rc_launch<K, float, true>(const Plan&, const KernelArgs&, cudaStream_t);      \
    extern template cudaError_t rc_launch<K, __nv_bfloat16, false>(const Plan&, const KernelArgs&, cudaStream_t); \
    extern template cudaError_t rc_launch<K, __nv_bfloat16, true>(const Plan&, const KernelArgs&, cudaStream_t);  \
    extern template cudaError_t rc_launch<K, __half, false>(const Plan&, const KernelArgs&, cudaStream_t);    \
    extern template cudaError_t rc_launch<K, __half, true>(const Plan&, const KernelArgs&, cudaStream_t);
RC_DECLARE_K(3)
RC_DECLARE_K(5)
RC_DECLARE_K(7)

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

template <int K>
static rc_launch_fn pick_dtype(int dtype, bool bwd) {
    switch (dtype) {
        case RECNEXT_F32: return bwd ? rc_launch<K, float, true> : rc_launch<K, float, false>;
        case RECNEXT_BF16: return bwd ? rc_launch<K, __nv_bfloat16, true> : rc_launch<K, __nv_bfloat16, false>;
        case RECNEXT_F16: return bwd ? rc_launch<K, __half, true> : rc_launch<K, __half, false>;
    }
    return nullptr;
}
static rc_launch_fn pick(int k, int dtype, bool bwd) {
    switch (k) {
        case 3: return pick_dtype<3>(dtype, bwd);
        case 5: return pick_dtype<5>(dtype, bwd);
        case 7: return pick_dtype<7>(dtype, bwd);
    }
    return nullptr;
}

template <int K>
static w_launch_fn w_pick_dtype(int dtype, bool bwd) {
    switch (dtype) {
        case RECNEXT_F32: return bwd ? w_launch<K, float, true> : w_launch<K, float, false>;
        case RECNEXT_BF16: return bwd ? w_launch<K, __nv_bfloat16, true> : w_launch<K, __nv_bfloat16, false>;
        case RECNEXT_F16: return bwd ? w_launch<K, __half, true> : w_launch<K, __half, false>;
    }
    return nullptr;
}
static w_launch_fn w_pick(int k, int dtype, bool bwd) {
    switch (k) {
        case 3: return w_pick_dtype<3>(dtype, bwd);
        case 5: return w_pick_dtype<5>(dtype, bwd);
        case 7: return w_pick_dtype<7>(dtype, bwd);
    }
    return nullptr;
}
// RECNEXT_PATH=legacy forces the big-plane kernels (A/B measurements); default: team-resident path when it fits
static bool legacy_forced() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RECNEXT_PATH"); v = (e && strcmp(e, "legacy") == 0) ? 1 : 0; }
    return v == 1;
}

// RECNEXT_PROF=1: team 0 of CTA 0 records clock64() after every stage into a 32 KB device buffer that
// recnext_debug_prof() copies out (timing experiments only; not part of the documented ABI)
static long long* g_prof = nullptr;
static long long* prof_buffer() {
    static int init = 0;
    if (!init) {
        init = 1;
        const char* e = getenv("RECNEXT_PROF");
        if ((e && atoi(e)) || getenv("RECNEXT_FFN_PROF")) {
            const cudaError_t me = cudaMalloc(&g_prof, 4096 * sizeof(long long));
            if (me != cudaSuccess) { fprintf(stderr, "recnext: profiling buffer: %s\n", cudaGetErrorString(me)); g_prof = nullptr; }
        }
    }
    if (g_prof) cudaMemset(g_prof, 0, 4096 * sizeof(long long));
    return g_prof;
}

static int device_sms() { return rc_device_sms(); }   // (planning without a visible device, e.g. plan_describe on a CPU box: 148)

static int check_desc(const recconv_desc* d) {
    if (!d) return fail(RECNEXT_EINVAL, "recconv: null descriptor");
    if (d->B < 0 || d->C < 0 || d->H < 1 || d->W < 1) return fail(RECNEXT_EINVAL, "recconv: bad shape [%d,%d,%d,%d]", d->B, d->C, d->H, d->W);
    if (!(d->k == 3 || d->k == 5 || d->k == 7)) return fail(RECNEXT_EINVAL, "recconv: kernel_size %d not in {3,5,7}", d->k);
    if (d->level < 0 || d->level > RECNEXT_MAX_LEVEL) return fail(RECNEXT_EINVAL, "recconv: level %d outside 0..%d", d->level, RECNEXT_MAX_LEVEL);
    if (d->mode != RECNEXT_BILINEAR && d->mode != RECNEXT_NEAREST) return fail(RECNEXT_EINVAL, "recconv: mode %d (0 bilinear, 1 nearest)", d->mode);
    if (d->dtype < 0 || d->dtype > 2) return fail(RECNEXT_EINVAL, "recconv: dtype %d", d->dtype);
    if (d->wdtype != RECNEXT_F32 && d->wdtype != d->dtype) return fail(RECNEXT_EINVAL, "recconv: parameters must be fp32 or match dtype");
    return 0;
}

// 0: big-plane plan made; 1: the plane's pyramid does not fit in shared memory (caller takes the streamed path); < 0: error
static int make_plan(const recconv_desc* d, bool bwd, Plan& pl) {
    PlanOptions opt;
    opt.num_sms = device_sms();
    const int rc = rc_make_plan(pl, d->B, d->C, d->H, d->W, d->k, d->level, d->mode, d->dtype, d->wdtype, d->has_bias, bwd ? 1 : 0, opt);
    if (rc == 1) return 1;
    if (rc) return fail(RECNEXT_EINVAL, "recconv: bad arguments");
    return 0;
}
static bool fma_forced();
// 0: tensor-core backward plan made (16-bit activations, k = 5, the plane's four pyramid sets fit on chip); 1: not eligible
static int make_wplan(const recconv_desc* d, bool bwd, WPlan& pl);
static int make_mbplan(const recconv_desc* d, MBPlan& bp) {
    if (fma_forced() || d->k != 5 || !(d->dtype == RECNEXT_BF16 || d->dtype == RECNEXT_F16)) return 1;
    // planes below 20 x 20: the many tiny stages of the tensor-core schedule are pure latency and the FMA kernel, which packs
    // several planes into a warp, is as fast or faster (measured [256,256,14,14]: 0.86 ms FMA vs 0.90 ms); RECNEXT_PATH=mma overrides
    {
        const char* e = getenv("RECNEXT_PATH");
        WPlan wp;
        if (!(e && strcmp(e, "mma") == 0) && d->H * d->W < 400 && make_wplan(d, true, wp) == 0) return 1;
    }
    MBPlanOptions opt;
    opt.num_sms = device_sms();
    if (const char* e = getenv("RECNEXT_MBG")) opt.force_G = atoi(e);
    if (const char* e = getenv("RECNEXT_MBTW")) opt.force_TW = atoi(e);
    if (const char* e = getenv("RECNEXT_MBNT")) opt.force_NT = atoi(e);
    return mb_make_plan(bp, d->B, d->C, d->H, d->W, d->k, d->level, d->mode, d->dtype, d->wdtype, d->has_bias, opt) == 0 ? 0 : 1;
}
static GStreamDesc stream_desc(const recconv_desc* d) {
    GStreamDesc g;
    g.B = d->B; g.C = d->C; g.H = d->H; g.W = d->W; g.K = d->k; g.L = d->level; g.mode = d->mode; g.dtype = d->dtype; g.wdtype = d->wdtype;
    g.has_bias = d->has_bias; g.num_sms = device_sms();
    return g;
}
// RECNEXT_PATH=stream forces the streamed path (tests, A/B measurements)
static bool stream_forced() {   // read on every call so that one process can test both paths
    const char* e = getenv("RECNEXT_PATH");
    return e && strcmp(e, "stream") == 0;
}

// 0: team-resident plan made; 1: not eligible (use the big-plane path)
static int make_wplan(const recconv_desc* d, bool bwd, WPlan& pl) {
    if (legacy_forced() || !w_pick(d->k, d->dtype, bwd)) return 1;
    WPlanOptions opt;
    opt.num_sms = device_sms();
    // tuning overrides for experiments (tools/kbench.py): RECNEXT_G / RECNEXT_TW / RECNEXT_NT / RECNEXT_MAXW
    if (const char* e = getenv("RECNEXT_G")) opt.force_G = atoi(e);
    if (const char* e = getenv("RECNEXT_TW")) opt.force_TW = atoi(e);
    if (const char* e = getenv("RECNEXT_NT")) opt.force_NT = atoi(e);
    if (const char* e = getenv("RECNEXT_MAXW")) opt.max_warps = atoi(e);
    if (const char* e = getenv("RECNEXT_DBG")) opt.dbg = atoi(e);
    return w_make_plan(pl, d->B, d->C, d->H, d->W, d->k, d->level, d->mode, d->dtype, d->wdtype, d->has_bias, bwd ? 1 : 0, opt) == 0 ? 0 : 1;
}

// RECNEXT_PATH=fma keeps the FP32-FMA kernels for 16-bit activations too (A/B measurements, tests of both paths)
static bool fma_forced() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RECNEXT_PATH"); v = (e && (strcmp(e, "fma") == 0 || strcmp(e, "legacy") == 0)) ? 1 : 0; }
    return v == 1;
}
// 0: tensor-core forward plan made (16-bit activations, k = 5); 1: not eligible
static int make_mplan(const recconv_desc* d, MPlan& pl) {
    if (fma_forced() || d->k != 5 || !(d->dtype == RECNEXT_BF16 || d->dtype == RECNEXT_F16)) return 1;
    MPlanOptions opt;
    opt.num_sms = device_sms();
    if (const char* e = getenv("RECNEXT_MG")) opt.force_G = atoi(e);
    if (const char* e = getenv("RECNEXT_MTW")) opt.force_TW = atoi(e);
    if (const char* e = getenv("RECNEXT_MNT")) opt.force_NT = atoi(e);
    if (const char* e = getenv("RECNEXT_MNOTMA")) opt.force_no_tma = atoi(e);
    if (const char* e = getenv("RECNEXT_MDBG")) opt.dbg = atoi(e);
    // planes smaller than one 16x8 MMA tile waste most of the tensor work: the FMA kernels are faster there
    // (measured, [256,512,7,7] L1 bf16: 0.098 ms FMA vs 0.118 ms MMA); RECNEXT_PATH=mma overrides for experiments
    static int force_mma = -1;
    if (force_mma < 0) { const char* e = getenv("RECNEXT_PATH"); force_mma = (e && strcmp(e, "mma") == 0) ? 1 : 0; }
    if (!force_mma && d->H * d->W < 100) return 1;
    // the specialised (compile-time geometry) kernels of the small stage shapes run 24 warps per SM; anything that does
    // not match one of them is planned for the generic kernel (16 warps at 128 registers)
    if (!opt.force_NT && !opt.force_TW) {
        opt.max_warps = m_static_max_warps(d->H, d->W);
        if (opt.max_warps > 16) {
            if (m_make_plan(pl, d->B, d->C, d->H, d->W, d->k, d->level, d->mode, d->dtype, d->wdtype, d->has_bias, opt) == 0 &&
                (pl.threads <= 512 || m_static_geometry(pl)))
                return 0;
            opt.max_warps = 16;
        }
    }
    return m_make_plan(pl, d->B, d->C, d->H, d->W, d->k, d->level, d->mode, d->dtype, d->wdtype, d->has_bias, opt) == 0 ? 0 : 1;
}

static int fill_args(const recconv_desc* d, const recconv_params* p, KernelArgs& a) {
    if (!p) return fail(RECNEXT_EINVAL, "recconv: null params");
    memset(&a, 0, sizeof(a));
    a.w[0] = p->w_down; a.b[0] = d->has_bias ? p->b_down : nullptr;
    if (d->level > 0 && !p->w_down) return fail(RECNEXT_EINVAL, "recconv: down.weight is null");
    if (d->level > 0 && d->has_bias && !p->b_down) return fail(RECNEXT_EINVAL, "recconv: down.bias is null");
    for (int j = 0; j <= d->level; ++j) {
        if (!p->w_convs[j]) return fail(RECNEXT_EINVAL, "recconv: convs.%d.weight is null", j);
        if (d->has_bias && !p->b_convs[j]) return fail(RECNEXT_EINVAL, "recconv: convs.%d.bias is null", j);
        a.w[1 + j] = p->w_convs[j];
        a.b[1 + j] = d->has_bias ? p->b_convs[j] : nullptr;
    }
    return 0;
}

__global__ void recconv_wgrad_finalize(const float* __restrict__ partial, float* __restrict__ gw, float* __restrict__ gb,
                                       int n_chunk, int nstage, int C, int KK, int ws) {
    const long total = (long)nstage * C * ws;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long sc = i / ws;
        const int e = (int)(i - sc * ws);
        float s = 0.f;
        for (int ch = 0; ch < n_chunk; ++ch) s += partial[(long)ch * total + i];
        if (e < KK) gw[sc * KK + e] = s;
        else if (e == KK && gb) gb[sc] = s;
    }
}

}  // namespace recnext

using namespace recnext;

extern "C" {

RECNEXT_API int recnext_abi_version(void) { return RECNEXT_ABI_VERSION; }
RECNEXT_API const char* recnext_last_error(void) { return g_err; }

// 0: one of the on-chip kernels takes this forward; 1: streamed path (needs a workspace)
static int forward_needs_stream(const recconv_desc* d) {
    if (stream_forced()) return 1;
    MPlan mp;
    if (make_mplan(d, mp) == 0) return 0;
    WPlan wp;
    if (make_wplan(d, false, wp) == 0) return 0;
    Plan pl;
    return make_plan(d, false, pl) == 1 ? 1 : 0;
}

RECNEXT_API size_t recconv_forward_workspace_bytes(const recconv_desc* d) {
    if (check_desc(d)) return 0;
    if (d->B == 0 || d->C == 0) return 0;
    return forward_needs_stream(d) ? gstream_workspace_bytes(stream_desc(d), false) : 0;
}

RECNEXT_API int recconv_forward_ws(const recconv_desc* d, const recconv_params* p, const void* x, void* y, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    if (int rc = check_desc(d)) return rc;
    if (d->B == 0 || d->C == 0) return RECNEXT_OK;
    if (!x || !y) return fail(RECNEXT_EINVAL, "recconv_forward: null tensor");
    KernelArgs a;
    if (int rc = fill_args(d, p, a)) return rc;
    a.x = x; a.out = y;
    MPlan mp;
    if (!stream_forced() && (((uintptr_t)x | (uintptr_t)y) & 3) == 0 && make_mplan(d, mp) == 0) {
        if (((uintptr_t)x & 15) != 0) mp.use_tma = 0;  // bulk copies need 16-byte aligned sources (raw buffer stays allocated)
        const cudaError_t e = m_launch(mp, a, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recconv_forward(mma): %s", cudaGetErrorString(e));
        return RECNEXT_OK;
    }
    WPlan wp;
    if (!stream_forced() && make_wplan(d, false, wp) == 0) {
        a.prof = prof_buffer();
        if (((uintptr_t)x & 15) != 0) wp.use_tma = 0;  // bulk copies need 16-byte aligned sources
        const cudaError_t e = w_pick(d->k, d->dtype, false)(wp, a, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recconv_forward: %s", cudaGetErrorString(e));
        return RECNEXT_OK;
    }
    Plan pl;
    const int prc = stream_forced() ? 1 : make_plan(d, false, pl);
    if (prc < 0) return prc;
    if (prc == 1) {  // the pyramid of one plane does not fit on chip: level-by-level through the workspace (gstream.cu)
        const GStreamDesc g = stream_desc(d);
        const size_t need = gstream_workspace_bytes(g, false);
        if (!workspace || workspace_bytes < need)
            return fail(RECNEXT_EWORKSPACE, "recconv_forward: a %dx%d plane pyramid (level %d) does not fit in shared memory; the streamed path needs a "
                        "workspace of %zu bytes (recconv_forward_workspace_bytes), got %zu", d->H, d->W, d->level, need, workspace_bytes);
        const cudaError_t e = gstream_launch(g, a, workspace, false, nullptr, nullptr, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recconv_forward(stream): %s", cudaGetErrorString(e));
        return RECNEXT_OK;
    }
    rc_launch_fn fn = pick(d->k, d->dtype, false);
    const cudaError_t e = fn(pl, a, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recconv_forward: %s", cudaGetErrorString(e));
    return RECNEXT_OK;
}

RECNEXT_API int recconv_forward(const recconv_desc* d, const recconv_params* p, const void* x, void* y, void* stream) {
    return recconv_forward_ws(d, p, x, y, nullptr, 0, stream);
}

// RecAttn2d pieces: variants 1 / 2 of the tensor-core forward kernel
static int recattn_launch(const recconv_desc* d, int variant, const void* w, const void* b, const void* x, const void* z, int zH, int zW,
                          void* out, void* stream, const char* what) {
    if (int rc = check_desc(d)) return rc;
    if (d->B == 0 || d->C == 0) return RECNEXT_OK;
    if (!x || !out || !w || (variant == 2 && !z)) return fail(RECNEXT_EINVAL, "%s: null tensor", what);
    if ((d->has_bias != 0) != (b != nullptr)) return fail(RECNEXT_EINVAL, "%s: b must be given iff has_bias", what);
    if (d->k != 5 || !(d->dtype == RECNEXT_BF16 || d->dtype == RECNEXT_F16)) {
        // fp32 activations (the 1e-5 bar) or k = 3 / 7: the plain grid-stride kernel of gstream.cu (no workspace)
        const GStreamDesc g = stream_desc(d);
        const cudaError_t e = gstream_recattn(g, variant, w, b, x, z, zH, zW, out, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "%s(fp32): %s", what, cudaGetErrorString(e));
        return RECNEXT_OK;
    }
    if ((((uintptr_t)x | (uintptr_t)out | (uintptr_t)z) & 3) != 0) return fail(RECNEXT_EINVAL, "%s: tensors must be 4-byte aligned", what);
    MPlanOptions opt;
    opt.num_sms = device_sms();
    opt.variant = variant; opt.zH = zH; opt.zW = zW;
    if (const char* e = getenv("RECNEXT_MDBG")) opt.dbg = atoi(e);
    if (const char* e = getenv("RECNEXT_AG")) opt.force_G = atoi(e);       // team-shape experiments (tools/ra_prof.py)
    if (const char* e = getenv("RECNEXT_ATW")) opt.force_TW = atoi(e);
    if (const char* e = getenv("RECNEXT_ANT")) opt.force_NT = atoi(e);
    if (getenv("RECNEXT_APLAN")) {
        MPlan q;
        if (m_make_plan(q, d->B, d->C, d->H, d->W, d->k, 1, d->mode, d->dtype, d->wdtype, d->has_bias, opt) == 0)
            fprintf(stderr, "[recattn variant %d %dx%d] G=%d TW=%d NTEAM=%d threads=%d smem=%d grid=%d tma=%d\n", variant, d->H, d->W, q.G, q.TW, q.NTEAM, q.threads, q.smem_bytes, q.grid, q.use_tma);
    }
    MPlan mp;
    const int rc = m_make_plan(mp, d->B, d->C, d->H, d->W, d->k, 1, d->mode, d->dtype, d->wdtype, d->has_bias, opt);
    if (rc == 1) return fail(RECNEXT_EUNSUPPORTED, "%s: a %dx%d plane does not fit in 227 KB of shared memory", what, d->H, d->W);
    if (rc) return fail(RECNEXT_EINVAL, "%s: bad arguments", what);
    if (((uintptr_t)x & 15) != 0) mp.use_tma = 0;
    KernelArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.gy = z; a.out = out;
    if (variant == 1) { a.w[0] = w; a.b[0] = b; }
    else { a.w[2] = w; a.b[2] = b; }   // slot of convs[L] with L = 1
    const cudaError_t e = m_launch(mp, a, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    return RECNEXT_OK;
}

RECNEXT_API int recattn_down_forward(const recconv_desc* d, const void* w, const void* b, const void* x, void* out, void* stream) {
    return recattn_launch(d, 1, w, b, x, nullptr, 0, 0, out, stream, "recattn_down_forward");
}
RECNEXT_API int recattn_up_forward(const recconv_desc* d, const void* w, const void* b, const void* x, const void* z, int32_t zH, int32_t zW,
                       void* y, void* stream) {
    if (zH < 1 || zW < 1) return fail(RECNEXT_EINVAL, "recattn_up_forward: bad z size %dx%d", zH, zW);
    return recattn_launch(d, 2, w, b, x, z, zH, zW, y, stream, "recattn_up_forward");
}

static int linattn_call(const char* what, int32_t B, int32_t dim, int32_t heads, int32_t n, int32_t dtype, const void* q, const void* k, const float* qb,
                        const float* kb, const void* v, const void* pe, void* out, void* stream, const float* pew = nullptr, const float* peb = nullptr,
                        int32_t pw = 0) {
    if (B < 0 || dim < 1 || heads < 1 || n < 1) return fail(RECNEXT_EINVAL, "%s: bad shape [%d,%d,%d] heads %d", what, B, dim, n, heads);
    if (B == 0) return RECNEXT_OK;
    if (!q || !k || !v || !out) return fail(RECNEXT_EINVAL, "%s: null tensor", what);
    cudaError_t e = cudaSuccess;
    const int rc = linattn_launch(B, dim, heads, n, dtype, q, k, qb, kb, v, pe, pew, peb, pw, out, (cudaStream_t)stream, &e);
    if (rc == 1) return fail(RECNEXT_EUNSUPPORTED, "%s: head_dim in {4,8,16,20,24,28,32,40} only (dim %d, heads %d); n must be a multiple of the plane width", what, dim, heads);
    if (rc) return fail(RECNEXT_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    return RECNEXT_OK;
}

RECNEXT_API int recnext_linattn_forward(int32_t B, int32_t dim, int32_t heads, int32_t n, int32_t dtype, const void* qk, const void* v, const void* pe,
                            void* out, void* stream) {
    const size_t esz = dtype == RECNEXT_F32 ? 4 : 2;
    const void* k = qk ? (const void*)((const char*)qk + (size_t)dim * (size_t)n * esz) : nullptr;
    return linattn_call("recnext_linattn_forward", B, dim, heads, n, dtype, qk, k, nullptr, nullptr, v, pe, out, stream);
}

RECNEXT_API int recnext_linattn_forward_qk(int32_t B, int32_t dim, int32_t heads, int32_t n, int32_t dtype, const void* q, const void* k, const float* qbias,
                               const float* kbias, const void* v, const void* pe, void* out, void* stream) {
    return linattn_call("recnext_linattn_forward_qk", B, dim, heads, n, dtype, q, k, qbias, kbias, v, pe, out, stream);
}

RECNEXT_API int recnext_linattn_forward_pe(int32_t B, int32_t dim, int32_t heads, int32_t H, int32_t W, int32_t dtype, const void* q, const void* k,
                               const float* qbias, const float* kbias, const void* v, const float* pe_w, const float* pe_b, void* out, void* stream) {
    if (H < 1 || W < 1 || !pe_w) return fail(RECNEXT_EINVAL, "recnext_linattn_forward_pe: bad plane %dx%d or null pe_w", H, W);
    return linattn_call("recnext_linattn_forward_pe", B, dim, heads, H * W, dtype, q, k, qbias, kbias, v, nullptr, out, stream, pe_w, pe_b, W);
}

RECNEXT_API int recnext_dwdown_forward(int32_t B, int32_t C, int32_t H, int32_t W, int32_t dtype, const void* x, const float* w, const float* b,
                           void* out, void* stream) {
    if (B < 0 || C < 1 || H < 1 || W < 1) return fail(RECNEXT_EINVAL, "recnext_dwdown_forward: bad shape [%d,%d,%d,%d]", B, C, H, W);
    if (B == 0) return RECNEXT_OK;
    if (!x || !w || !b || !out) return fail(RECNEXT_EINVAL, "recnext_dwdown_forward: null tensor");
    if ((((uintptr_t)out) & 15) != 0 || (((uintptr_t)x) & 1) != 0) return fail(RECNEXT_EINVAL, "recnext_dwdown_forward: out must be 16-byte aligned (vector stores)");
    cudaError_t e = cudaSuccess;
    const int rc = dwdown_launch(B, C, H, W, dtype, x, w, b, out, (cudaStream_t)stream, &e);
    if (rc == 1) return fail(RECNEXT_EUNSUPPORTED, "recnext_dwdown_forward: a padded fp32 %dx%d plane must fit in shared memory", H, W);
    if (rc) return fail(RECNEXT_ECUDA, "recnext_dwdown_forward: %s", cudaGetErrorString(e));
    return RECNEXT_OK;
}

RECNEXT_API int recnext_ffn_forward(int32_t B, int32_t C, int32_t hidden, int32_t HW, int32_t dtype, const void* y, const void* x,
                        const void* w1, const float* b1, const void* w2, const float* b2, void* out, void* stream) {
    if (B < 0 || C < 1 || hidden < 1 || HW < 1) return fail(RECNEXT_EINVAL, "recnext_ffn_forward: bad shape [%d,%d,%d] hidden %d", B, C, HW, hidden);
    if (B == 0) return RECNEXT_OK;
    if (!y || !x || !w1 || !b1 || !w2 || !b2 || !out) return fail(RECNEXT_EINVAL, "recnext_ffn_forward: null tensor");
    if ((((uintptr_t)y | (uintptr_t)x | (uintptr_t)out) & 15) != 0 || (((uintptr_t)w1 | (uintptr_t)w2) & 3) != 0)
        return fail(RECNEXT_EINVAL, "recnext_ffn_forward: activations must be 16-byte aligned");
    FfnPlan pl;
    if (ffn_make_plan(pl, B, C, hidden, HW, dtype))
        return fail(RECNEXT_EUNSUPPORTED, "recnext_ffn_forward: needs 16-bit activations, C %% 16 == 0, hidden %% 16 == 0, HW %% 4 == 0 and "
                    "(2C + hidden) pixel-tile rows in 227 KB of shared memory (got C=%d hidden=%d HW=%d dtype=%d)", C, hidden, HW, dtype);
    const cudaError_t e = ffn_launch(pl, y, x, w1, b1, w2, b2, out, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recnext_ffn_forward: %s", cudaGetErrorString(e));
    return RECNEXT_OK;
}

// ---- fused stem (stem.cu)
RECNEXT_API int recnext_stem_forward(int32_t B, int32_t H, int32_t W, int32_t C1, int32_t C2, int32_t dtype, const void* x, const void* w1p,
                                     const float* b1p, const void* w2p, const float* b2p, void* out, void* stream) {
    if (B < 0 || H < 1 || W < 1 || C1 < 1 || C2 < 1) return fail(RECNEXT_EINVAL, "recnext_stem_forward: bad shape [%d,3,%d,%d] -> %d -> %d channels", B, H, W, C1, C2);
    if (B == 0) return RECNEXT_OK;
    if (!x || !w1p || !b1p || !w2p || !b2p || !out) return fail(RECNEXT_EINVAL, "recnext_stem_forward: null tensor");
    if ((((uintptr_t)w1p | (uintptr_t)w2p) & 15) != 0) return fail(RECNEXT_EINVAL, "recnext_stem_forward: packed weights must be 16-byte aligned");
    StemPlan pl;
    if (stem_make_plan(pl, B, H, W, C1, C2, dtype, device_sms()))
        return fail(RECNEXT_EUNSUPPORTED, "recnext_stem_forward: needs 16-bit activations, C1 <= 48 and C2 <= 80 (got C1=%d C2=%d dtype=%d)", C1, C2, dtype);
    const cudaError_t e = stem_launch(pl, x, w1p, b1p, w2p, b2p, out, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recnext_stem_forward: %s", cudaGetErrorString(e));
    return RECNEXT_OK;
}

// ---- tcgen05 channel mixer (ffn_tc.cu)
RECNEXT_API size_t recnext_ffn_packed_bytes(int32_t C, int32_t hidden) {
    FfnTcPlan pl;
    if (C < 1 || hidden < 1 || ffn_tc_make_plan(pl, 1, C, hidden, 8, RECNEXT_BF16, device_sms())) return 0;
    return pl.packed_bytes;
}

RECNEXT_API int recnext_ffn_pack(int32_t C, int32_t hidden, int32_t dtype, const void* w1, const void* w2, void* packed, void* stream) {
    if (!w1 || !w2 || !packed) return fail(RECNEXT_EINVAL, "recnext_ffn_pack: null tensor");
    FfnTcPlan pl;
    if (C < 1 || hidden < 1 || ffn_tc_make_plan(pl, 1, C, hidden, 8, dtype, device_sms()))
        return fail(RECNEXT_EUNSUPPORTED, "recnext_ffn_pack: needs 16-bit weights and C %% 8 == 0 (got C=%d hidden=%d dtype=%d)", C, hidden, dtype);
    const cudaError_t e = ffn_tc_pack(pl, w1, w2, packed, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recnext_ffn_pack: %s", cudaGetErrorString(e));
    return RECNEXT_OK;
}

RECNEXT_API int recnext_ffn_forward_packed(int32_t B, int32_t C, int32_t hidden, int32_t HW, int32_t dtype, const void* y, const void* x,
                                           const void* packed, const float* b1, const float* b2, void* out, void* stream) {
    if (B < 0 || C < 1 || hidden < 1 || HW < 1) return fail(RECNEXT_EINVAL, "recnext_ffn_forward_packed: bad shape [%d,%d,%d] hidden %d", B, C, HW, hidden);
    if (B == 0) return RECNEXT_OK;
    if (!y || !x || !packed || !b1 || !b2 || !out) return fail(RECNEXT_EINVAL, "recnext_ffn_forward_packed: null tensor");
    if ((((uintptr_t)y | (uintptr_t)x | (uintptr_t)out | (uintptr_t)packed) & 15) != 0)
        return fail(RECNEXT_EINVAL, "recnext_ffn_forward_packed: tensors must be 16-byte aligned");
    FfnTcPlan pl;
    if (ffn_tc_make_plan(pl, B, C, hidden, HW, dtype, device_sms()))
        return fail(RECNEXT_EUNSUPPORTED, "recnext_ffn_forward_packed: needs 16-bit activations, C %% 8 == 0 and C <= 768 (got C=%d hidden=%d HW=%d dtype=%d)",
                    C, hidden, HW, dtype);
    if (getenv("RECNEXT_FFN_PROF")) ffn_tc_set_prof(prof_buffer());   // stamps are read back with recnext_debug_prof()
    const cudaError_t e = ffn_tc_launch(pl, y, x, packed, b1, b2, out, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recnext_ffn_forward_packed: %s", cudaGetErrorString(e));
    return RECNEXT_OK;
}

RECNEXT_API size_t recconv_backward_workspace_bytes(const recconv_desc* d) {
    if (check_desc(d)) return 0;
    if (d->B == 0 || d->C == 0) return 0;
    MBPlan bp;
    if (!stream_forced() && make_mbplan(d, bp) == 0) return (size_t)bp.ws_floats * sizeof(float);
    WPlan wp;
    if (!stream_forced() && make_wplan(d, true, wp) == 0) return (size_t)wp.ws_partial_floats * sizeof(float);
    Plan pl;
    const int prc = stream_forced() ? 1 : make_plan(d, true, pl);
    if (prc == 1) return gstream_workspace_bytes(stream_desc(d), true);
    if (prc) return 0;
    return (size_t)pl.ws_partial_floats * sizeof(float);
}

RECNEXT_API int recconv_backward(const recconv_desc* d, const recconv_params* p, const void* x, const void* gy, void* gx, float* gw,
                     float* gb, void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = check_desc(d)) return rc;
    if (!gw) return fail(RECNEXT_EINVAL, "recconv_backward: gw is null");
    if ((d->has_bias != 0) != (gb != nullptr)) return fail(RECNEXT_EINVAL, "recconv_backward: gb must be given iff has_bias");
    const int KK = d->k * d->k;
    if (d->B == 0 || d->C == 0) {
        if (d->C) {
            cudaMemsetAsync(gw, 0, sizeof(float) * (d->level + 2) * d->C * KK, (cudaStream_t)stream);
            if (gb) cudaMemsetAsync(gb, 0, sizeof(float) * (d->level + 2) * d->C, (cudaStream_t)stream);
        }
        return RECNEXT_OK;
    }
    if (!x || !gy || !gx) return fail(RECNEXT_EINVAL, "recconv_backward: null tensor");
    KernelArgs a;
    if (int rc = fill_args(d, p, a)) return rc;
    a.x = x; a.gy = gy; a.out = gx; a.partial = reinterpret_cast<float*>(workspace);
    WPlan wp;
    Plan pl;
    int n_partials = 0, wstride = 0;
    cudaError_t e;
    MBPlan bp;
    if (!stream_forced() && (((uintptr_t)x | (uintptr_t)gy | (uintptr_t)gx) & 3) == 0 && make_mbplan(d, bp) == 0) {
        // 16-bit activations, k = 5: the tensor-core backward (mbwd.cuh) writes gw / gb through its own deterministic finalize
        const size_t need = (size_t)bp.ws_floats * sizeof(float);
        if (!workspace || workspace_bytes < need)
            return fail(RECNEXT_EWORKSPACE, "recconv_backward: workspace %zu bytes < %zu needed", workspace_bytes, need);
        a.prof = prof_buffer();
        if ((((uintptr_t)x | (uintptr_t)gy) & 15) != 0) bp.use_tma = 0;   // bulk copies need 16-byte aligned sources
        e = mb_launch(bp, a, gw, gb, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recconv_backward(mma): %s", cudaGetErrorString(e));
        return RECNEXT_OK;
    }
    if (!stream_forced() && make_wplan(d, true, wp) == 0) {
        const size_t need = (size_t)wp.ws_partial_floats * sizeof(float);
        if (!workspace || workspace_bytes < need)
            return fail(RECNEXT_EWORKSPACE, "recconv_backward: workspace %zu bytes < %zu needed", workspace_bytes, need);
        if ((((uintptr_t)x | (uintptr_t)gy) & 15) != 0) wp.use_tma = 0;
        a.prof = prof_buffer();
        e = w_pick(d->k, d->dtype, true)(wp, a, (cudaStream_t)stream);
        n_partials = wp.tpc; wstride = wp.wstride;
    } else {
        const int prc = stream_forced() ? 1 : make_plan(d, true, pl);
        if (prc < 0) return prc;
        if (prc == 1) {  // the plane's pyramid does not fit on chip: streamed backward (gstream.cu), writes gw / gb itself
            const GStreamDesc g = stream_desc(d);
            const size_t need = gstream_workspace_bytes(g, true);
            if (!workspace || workspace_bytes < need)
                return fail(RECNEXT_EWORKSPACE, "recconv_backward: workspace %zu bytes < %zu needed", workspace_bytes, need);
            e = gstream_launch(g, a, workspace, true, gw, gb, (cudaStream_t)stream);
            if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recconv_backward(stream): %s", cudaGetErrorString(e));
            return RECNEXT_OK;
        }
        const size_t need = (size_t)pl.ws_partial_floats * sizeof(float);
        if (!workspace || workspace_bytes < need)
            return fail(RECNEXT_EWORKSPACE, "recconv_backward: workspace %zu bytes < %zu needed", workspace_bytes, need);
        rc_launch_fn fn = pick(d->k, d->dtype, true);
        e = fn(pl, a, (cudaStream_t)stream);
        n_partials = pl.n_chunk; wstride = pl.wstride;
    }
    if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recconv_backward: %s", cudaGetErrorString(e));
    const long total = (long)(d->level + 2) * d->C * wstride;
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads);
    recconv_wgrad_finalize<<<blocks, threads, 0, (cudaStream_t)stream>>>(a.partial, gw, gb, n_partials, d->level + 2, d->C, KK, wstride);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RECNEXT_ECUDA, "recconv_backward(finalize): %s", cudaGetErrorString(e));
    return RECNEXT_OK;
}

RECNEXT_API int recnext_debug_prof(long long* host_out, int n) {
    if (!g_prof) return 0;
    cudaDeviceSynchronize();
    return cudaMemcpy(host_out, g_prof, sizeof(long long) * (n < 4096 ? n : 4096), cudaMemcpyDeviceToHost) == cudaSuccess ? 1 : 0;
}

RECNEXT_API int recconv_plan_describe(const recconv_desc* d, int backward, char* buf, size_t buflen) {
    if (int rc = check_desc(d)) return rc;
    if (!buf || !buflen) return fail(RECNEXT_EINVAL, "recconv_plan_describe: null buffer");
    MPlan mp;
    if (!stream_forced() && !backward && make_mplan(d, mp) == 0) {
        snprintf(buf, buflen,
                 "fwd tensor-core k=5 L=%d [%d,%d,%d,%d] planes/batch=%d warps/team=%d teams/CTA=%d threads=%d grid=%d smem=%d B team=%d B "
                 "plane=%d B frag-regs/channel=%d tma=%d geometry=%s",
                 mp.L, mp.B, mp.C, mp.H, mp.W, mp.G, mp.TW, mp.NTEAM, mp.threads, mp.grid, mp.smem_bytes, mp.team_bytes, mp.l0_bytes + mp.upper_bytes,
                 mp.nregs, mp.use_tma, m_static_geometry(mp) ? "compile-time" : "run-time");
        return RECNEXT_OK;
    }
    MBPlan bp;
    if (!stream_forced() && backward && make_mbplan(d, bp) == 0) {
        snprintf(buf, buflen,
                 "bwd tensor-core k=5 L=%d [%d,%d,%d,%d] planes/batch=%d warps/team=%d teams/CTA=%d threads=%d grid=%d smem=%d B team=%d B "
                 "frag-regs/channel=%d tma=%d",
                 bp.f.L, bp.f.B, bp.f.C, bp.f.H, bp.f.W, bp.f.G, bp.f.TW, bp.f.NTEAM, bp.f.threads, bp.grid, bp.smem_bytes, bp.team_bytes, bp.nregs, bp.use_tma);
        return RECNEXT_OK;
    }
    WPlan wp;
    if (!stream_forced() && make_wplan(d, backward != 0, wp) == 0) {
        snprintf(buf, buflen,
                 "%s team-resident k=%d L=%d [%d,%d,%d,%d] planes/batch=%d warps/team=%d teams/CTA=%d threads=%d grid=%d lanes/plane=%d "
                 "teams/channel-group=%d smem=%d B team=%d B plane=%d B tma=%d",
                 backward ? "bwd" : "fwd", wp.K, wp.L, wp.B, wp.C, wp.H, wp.W, wp.G, wp.TW, wp.NT, wp.threads, wp.grid, wp.LPP, wp.tpc,
                 wp.smem_bytes, wp.team_bytes, wp.plane_floats * 4, wp.use_tma);
        return RECNEXT_OK;
    }
    Plan pl;
    const int prc = stream_forced() ? 1 : make_plan(d, backward != 0, pl);
    if (prc < 0) return prc;
    if (prc == 1) {
        snprintf(buf, buflen, "%s streamed k=%d L=%d [%d,%d,%d,%d] level-by-level through a %zu-byte fp32 workspace (plane pyramid exceeds shared memory)",
                 backward ? "bwd" : "fwd", d->k, d->level, d->B, d->C, d->H, d->W, gstream_workspace_bytes(stream_desc(d), backward != 0));
        return RECNEXT_OK;
    }
    int n = snprintf(buf, buflen,
                     "%s k=%d L=%d [%d,%d,%d,%d] planes/CTA=%d lanes/plane=%d units=%d threads=%d grid=%dx%d (cg x chunk, %d img/chunk) "
                     "smem=%d B plane=%d B tma=%d share_raw=%d rpi=",
                     backward ? "bwd" : "fwd", pl.K, pl.L, pl.B, pl.C, pl.H, pl.W, pl.P, pl.g, pl.n_units, pl.T, pl.n_cg, pl.n_chunk,
                     pl.img_per_chunk, pl.smem_bytes, pl.plane_floats * 4, pl.use_tma, pl.share_raw);
    for (int l = 0; l <= pl.L && n > 0 && (size_t)n < buflen; ++l)
        n += snprintf(buf + n, buflen - n, "%s%d/%d(%d)", l ? "," : "", pl.lv[l].g1.rpi, pl.lv[l].g2.rpi, pl.lv[l].g1.lanes);
    return RECNEXT_OK;
}

RECNEXT_API int recconv_source_index(int mode, int in_size, int out_size, int32_t* i0, int32_t* i1, float* lambda) {
    if (in_size < 1 || out_size < 1 || !i0) return fail(RECNEXT_EINVAL, "recconv_source_index: bad arguments");
    for (int d = 0; d < out_size; ++d) {
        if (mode == RECNEXT_NEAREST) {
            i0[d] = rc_nearest_src(in_size, out_size, d);
            if (i1) i1[d] = i0[d];
            if (lambda) lambda[d] = 0.f;
        } else {
            int a, b; float l;
            rc_bilinear_src(in_size, out_size, d, a, b, l);
            i0[d] = a;
            if (i1) i1[d] = b;
            if (lambda) lambda[d] = l;
        }
    }
    return RECNEXT_OK;
}

}  // extern "C"
