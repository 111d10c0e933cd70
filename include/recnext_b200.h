/*
 * recnext_b200.h — C ABI of the B200-native RecConv hot path.
 *
 * This is the drop-in boundary.  The reference has no native layer: its hot path is the
 * Python module RecConv2d (reference model/recnext.py:8-34), whose arithmetic is dispatched to ATen.  These
 * entry points are what a binding for that module calls instead of ATen:
 *
 *   recconv_forward   replaces  RecConv2d.forward          model/recnext.py:24-34
 *                               (down loop :27-29, up loop :31-33, final conv :34)
 *   recconv_backward  replaces  the autograd graph of the same lines (conv dgrad/wgrad, upsample backward)
 *   recconv_*_workspace_bytes / recconv_plan_describe: sizing and introspection helpers
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers owned by the caller (PyTorch); the library borrows them for
 *     the duration of the stream-ordered call.  It never allocates or frees device memory, never
 *     synchronises the device, and launches only on `stream` (safe under CUDA-graph capture).
 *   - Layout is NCHW contiguous (what the reference feeds the module).  Every (n, c) plane is independent.
 *   - Parameters follow the reference state_dict (model/recnext.py:21-22): `down.weight` [C,1,k,k],
 *     `convs.{j}.weight` [C,1,k,k] for j = 0..level, optional `down.bias` / `convs.{j}.bias` [C].
 *     They are passed as HOST arrays of device pointers so that no packing kernel is needed.
 *   - Return value: 0 on success, a negative RECNEXT_E* code otherwise; recnext_last_error() returns a
 *     thread-local message.  There is no CPU fallback: unsupported arguments are errors.
 */
#ifndef RECNEXT_B200_H_
#define RECNEXT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RECNEXT_ABI_VERSION 3

#if defined(__GNUC__)
#define RECNEXT_API __attribute__((visibility("default")))
#else
#define RECNEXT_API
#endif

/* element types of x / y / gy / gx (dtype) and of the parameters (wdtype) */
#define RECNEXT_F32 0
#define RECNEXT_BF16 1
#define RECNEXT_F16 2

/* interpolation modes of F.interpolate(..., size=s, mode=...) (model/recnext.py:33) */
#define RECNEXT_BILINEAR 0
#define RECNEXT_NEAREST 1

#define RECNEXT_MAX_LEVEL 6

#define RECNEXT_OK 0
#define RECNEXT_EINVAL (-1)     /* bad argument (shape, dtype, k, level, mode, null pointer) */
#define RECNEXT_EUNSUPPORTED (-2) /* valid for the reference but not built here (e.g. plane too large) */
#define RECNEXT_EWORKSPACE (-3) /* workspace too small */
#define RECNEXT_ECUDA (-4)      /* CUDA runtime error at launch */

typedef struct recconv_desc {
    int32_t B, C, H, W;   /* input  x: [B, C, H, W] */
    int32_t k;            /* odd kernel size: 3, 5 or 7 (reference default 5) */
    int32_t level;        /* number of stride-2 downsamples, 0..RECNEXT_MAX_LEVEL */
    int32_t mode;         /* RECNEXT_BILINEAR | RECNEXT_NEAREST */
    int32_t dtype;        /* RECNEXT_F32 | RECNEXT_BF16 | RECNEXT_F16 : x, y, gy, gx */
    int32_t wdtype;       /* dtype of weights and biases (F32 master weights, or same as dtype) */
    int32_t has_bias;     /* 0 | 1 */
} recconv_desc;

typedef struct recconv_params {
    const void* w_down;                              /* [C,1,k,k]   down.weight          */
    const void* w_convs[RECNEXT_MAX_LEVEL + 1];      /* [C,1,k,k]   convs.{j}.weight     */
    const void* b_down;                              /* [C] or NULL down.bias            */
    const void* b_convs[RECNEXT_MAX_LEVEL + 1];      /* [C] or NULL convs.{j}.bias       */
} recconv_params;

RECNEXT_API int recnext_abi_version(void);
RECNEXT_API const char* recnext_last_error(void);

/*
 * y = RecConv2d(x).  Planes whose whole pyramid fits in one SM's shared memory (every RecNeXt stage shape at 224 px,
 * detection stages 2-3, 16-bit detection stages 0-1) run fused and need no workspace.  Larger planes (fp32 detection
 * stages 0-1) are STREAMED level by level through an fp32 workspace of recconv_forward_workspace_bytes(d) bytes
 * (0 when the fused kernels take the call); recconv_forward == recconv_forward_ws without a workspace and returns
 * RECNEXT_EWORKSPACE for those.
 */
RECNEXT_API int recconv_forward(const recconv_desc* d, const recconv_params* p, const void* x, void* y, void* stream);
RECNEXT_API size_t recconv_forward_workspace_bytes(const recconv_desc* d);
RECNEXT_API int recconv_forward_ws(const recconv_desc* d, const recconv_params* p, const void* x, void* y, void* workspace,
                                   size_t workspace_bytes, void* stream);

/*
 * Backward.  gx: [B,C,H,W] in d->dtype.  Weight grads are fp32 and PACKED:
 *   gw [(level+2), C, k*k]: slot 0 = down.weight (summed over all levels — the filter is shared,
 *                            model/recnext.py:21,28), slot 1+j = convs.{j}.weight
 *   gb [(level+2), C] or NULL (same slot order); must be non-NULL iff has_bias.
 * Both are overwritten (not accumulated).  The result is deterministic (fixed reduction order).
 * workspace: recconv_backward_workspace_bytes(d) bytes of device memory, 16-byte aligned: a few hundred KB of partial
 * sums for the fused kernels; planes whose pyramid does not fit on chip (detection stages 0-1) take the streamed path,
 * whose workspace holds the fp32 pyramid and its gradients (~14 bytes per element of x).
 */
RECNEXT_API size_t recconv_backward_workspace_bytes(const recconv_desc* d);
RECNEXT_API int recconv_backward(const recconv_desc* d, const recconv_params* p, const void* x, const void* gy, void* gx,
                     float* gw, float* gb, void* workspace, size_t workspace_bytes, void* stream);

/*
 * RecAttn2d (A-series token mixer, reference model/recattn.py:54-67), the two plane-independent pieces around the
 * linear attention (BatchNorm folded into w, b as ConvNorm.fuse does, model/recattn.py:87-111).  16-bit activations with k = 5
 * (every RecNeXt-A model under autocast) run on the tensor-core kernel; fp32 activations (the 1e-5 parity bar) and k = 3 / 7 run
 * on a plain fp32-accumulating kernel without workspace (csrc/gstream.cu).  Inference entry points: there is no backward yet.
 *
 *   recattn_down_forward  replaces  RecAttn2d.down[0]                      model/recattn.py:60,67
 *       out[B,C,H1,W1] = depthwise k x k stride-2 conv of x[B,C,H,W] (+ b),  H1 = (H-1)/2+1, W1 = (W-1)/2+1
 *   recattn_up_forward    replaces  self.conv(x + F.interpolate(z, size=x.shape[2:], mode))   model/recattn.py:67
 *       y[B,C,H,W] = depthwise k x k conv of (x + interpolate(z[B,C,zH,zW])) (+ b)
 *
 * d->level is ignored; d->mode selects the interpolation of recattn_up_forward (the reference default is nearest).
 * w: [C,1,k,k], b: [C] or NULL (must be non-NULL iff d->has_bias), dtype d->wdtype.
 */
RECNEXT_API int recattn_down_forward(const recconv_desc* d, const void* w, const void* b, const void* x, void* out, void* stream);
RECNEXT_API int recattn_up_forward(const recconv_desc* d, const void* w, const void* b, const void* x, const void* z, int32_t zH,
                                   int32_t zW, void* y, void* stream);

/*
 * Fused channel mixer of a RecNeXt block on NCHW tensors: replaces
 *     x + channel_mixer(norm(y))        with y = token_mixer(x)          model/recnext.py:157-158
 * where channel_mixer = mlp = 1x1 conv -> GELU -> 1x1 conv (model/recnext.py:125-131) with its ConvNorms folded
 * (ConvNorm.fuse :75-97) and the eval-mode BatchNorm `norm` (:153) folded into (w1, b1) by the caller:
 *     out[b,c,p] = x[b,c,p] + b2[c] + sum_h w2[c,h] * gelu(b1[h] + sum_k w1[h,k] * y[b,k,p]),   p = pixel of H*W
 * y, x, out: [B, C, HW] (= NCHW) in dtype (RECNEXT_BF16 | RECNEXT_F16), 16-byte aligned; w1 [hidden, C], w2 [C, hidden] in
 * the same dtype; b1 [hidden], b2 [C] fp32.  Inference entry point.  Returns RECNEXT_EUNSUPPORTED (nothing launched)
 * unless C % 16 == 0, hidden % 16 == 0, HW % 4 == 0 and the tiles fit in shared memory (C <= 256 for hidden = 2C).
 */
RECNEXT_API int recnext_ffn_forward(int32_t B, int32_t C, int32_t hidden, int32_t HW, int32_t dtype, const void* y, const void* x,
                                    const void* w1, const float* b1, const void* w2, const float* b2, void* out, void* stream);

/*
 * The same channel mixer on the 5th-generation tensor cores (tcgen05.mma with TMEM accumulators, weights streamed by TMA
 * bulk copies; csrc/ffn_tc.cu) — the default path of recnext_b200.model.  It serves every RecNeXt width (C % 8 == 0,
 * C <= 768, any hidden width, any plane size).  The weights are consumed as a PACKED stream: 128 x 64 tiles in the
 * kernel's consumption order, each stored as the shared-memory image of a K-major tcgen05 operand, so that a tile is one
 * contiguous bulk copy.
 *   recnext_ffn_packed_bytes(C, hidden)          size of the packed stream (0 if the shape is not served)
 *   recnext_ffn_pack(..., w1, w2, packed)        w1 [hidden, C], w2 [C, hidden] (16-bit, row-major) -> packed (device kernel)
 *   recnext_ffn_forward_packed(...)              same contract as recnext_ffn_forward with (w1, w2) replaced by `packed`
 * Returns RECNEXT_EUNSUPPORTED (nothing launched) for other dtypes / widths.
 */
RECNEXT_API size_t recnext_ffn_packed_bytes(int32_t C, int32_t hidden);
RECNEXT_API int recnext_ffn_pack(int32_t C, int32_t hidden, int32_t dtype, const void* w1, const void* w2, void* packed, void* stream);
RECNEXT_API int recnext_ffn_forward_packed(int32_t B, int32_t C, int32_t hidden, int32_t HW, int32_t dtype, const void* y, const void* x,
                                           const void* packed, const float* b1, const float* b2, void* out, void* stream);

/*
 * The stem (SURVEY.md §8 f: the caller in front of the first RecConv stage): ConvNorm(3 -> C1, 3x3, stride 2, pad 1) -> GELU ->
 * ConvNorm(C1 -> C2, 3x3, stride 2, pad 1) with both BatchNorms folded by the caller — replaces
 *     self.stem(x)                         model/recnext.py:139-146
 * as one kernel: the intermediate map stays in shared memory.  x: [B, 3, H, W], out: [B, C2, H2, W2] with H1 = (H-1)/2+1,
 * H2 = (H1-1)/2+1, 16-bit dtype (RECNEXT_BF16 | RECNEXT_F16), fp32 accumulation, intermediates rounded where the reference's
 * autocast graph rounds them.  Weights in the kernel's operand order (what recnext_b200.model.stem_pack writes), 16-byte aligned:
 *   w1p [C1P][32]      k = ci * 9 + ky * 3 + kx, zero padded; C1P = 32 (C1 <= 32) or 48        b1p [C1P] fp32
 *   w2p [C2P][9 C1P]   k = (ky * 3 + kx) * C1P + ci, zero padded; C2P = C2 rounded up to 16     b2p [C2P] fp32
 * Inference entry point.  RECNEXT_EUNSUPPORTED for fp32 activations, C1 > 48 or C2 > 80 (the caller keeps its conv path).
 */
RECNEXT_API int recnext_stem_forward(int32_t B, int32_t H, int32_t W, int32_t C1, int32_t C2, int32_t dtype, const void* x, const void* w1p,
                                     const float* b1p, const void* w2p, const float* b2p, void* out, void* stream);

/*
 * Token mixer of a `Downsample` block: depthwise 7x7 stride-2 conv with channel multiplier 2 and the
 * eval-mode BatchNorm that follows it folded into (w, b) by the caller — replaces
 *     self.norm(self.token_mixer(x))      model/recnext.py:137-138,145
 * x: [B, C, H, W], out: [B, 2C, (H-1)/2+1, (W-1)/2+1] in dtype (RECNEXT_F32 | RECNEXT_BF16 | RECNEXT_F16; fp32 arithmetic);
 * w: [2C, 1, 7, 7] fp32, b: [2C] fp32.
 * out must be 16-byte aligned.  Inference entry point.  RECNEXT_EUNSUPPORTED if a padded plane does not fit in shared memory.
 */
RECNEXT_API int recnext_dwdown_forward(int32_t B, int32_t C, int32_t H, int32_t W, int32_t dtype, const void* x, const float* w,
                                       const float* b, void* out, void* stream);

/*
 * Contraction core of the A-series linear attention: replaces everything between the `qk` ConvNorm and
 * the final `+ self.pe(v)` of LinearAttention1.forward (model/recattn.py:21-28) / LinearAttention2.forward (:44-51) — the two
 * are the same function:   q, k = elu(qk) + 1;  out = q^T (k v^T / n) / (q^T mean(k) + 1e-6)  (+ pe), per image and head.
 * qk: [B, 2*dim, n] (output of the grouped 1x1 ConvNorm, pre-activation; q = first dim channels), v, pe, out: [B, dim, n]
 * (= NCHW with n = h*w) in dtype (RECNEXT_F32 | RECNEXT_BF16 | RECNEXT_F16; fp32 accumulation throughout); pe may be NULL.
 * head_dim = dim / heads in {4,8,16,20,24,28,32,40}
 * (every RecNeXt-A model), else RECNEXT_EUNSUPPORTED.  Inference entry point.
 */
RECNEXT_API int recnext_linattn_forward(int32_t B, int32_t dim, int32_t heads, int32_t n, int32_t dtype, const void* qk, const void* v,
                                        const void* pe, void* out, void* stream);
/*
 * The same contraction with q and k as two [B, dim, n] tensors (the pre-activation outputs of the two halves of the grouped 1x1 `qk`
 * conv, e.g. written by two batched GEMMs) and optional fp32 per-channel biases [dim] that are added before the elu — so the caller's
 * GEMM needs no bias pass.  model/recattn.py:13,21 (`qk`), :21-28.
 */
RECNEXT_API int recnext_linattn_forward_qk(int32_t B, int32_t dim, int32_t heads, int32_t n, int32_t dtype, const void* q, const void* k,
                                           const float* qbias, const float* kbias, const void* v, const void* pe, void* out, void* stream);
/*
 * ... and with the positional-encoding ConvNorm `pe` (depthwise 3x3, pad 1, BatchNorm folded: pe_w [dim, 9] fp32, pe_b [dim] fp32 or
 * NULL) evaluated on v inside the kernel instead of read as a tensor: everything of LinearAttention1/2.forward after the `qk` GEMMs
 * (model/recattn.py:14,21-28).  v / out: [B, dim, H, W].
 */
RECNEXT_API int recnext_linattn_forward_pe(int32_t B, int32_t dim, int32_t heads, int32_t H, int32_t W, int32_t dtype, const void* q, const void* k,
                                           const float* qbias, const float* kbias, const void* v, const float* pe_w, const float* pe_b, void* out,
                                           void* stream);

/* Writes a one-line description of the launch plan (tiling, shared memory, grid) for logs/benchmarks. */
RECNEXT_API int recconv_plan_describe(const recconv_desc* d, int backward, char* buf, size_t buflen);

/*
 * Source-index tables the kernels use for F.interpolate(size=out) from `in` (bit-exact contract with ATen
 * UpSample.h:259-311,441-476).  Host-side helper for tests/diagnostics: fills i0[out], i1[out], lambda[out]
 * (bilinear) or i0[out] only (nearest; i1/lambda may be NULL).
 */
RECNEXT_API int recconv_source_index(int mode, int in_size, int out_size, int32_t* i0, int32_t* i1, float* lambda);

#ifdef __cplusplus
}
#endif
#endif /* RECNEXT_B200_H_ */
