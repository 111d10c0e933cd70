// linattn_mma.cu — the A-series linear attention (reference model/recattn.py:16-28 / :39-51; see linattn.cu for the algebra) with its two
// contractions on the tensor cores, 16-bit activations:
//     kv[i, j] = (1/n) sum_p k[i, p] v[j, p]     (a ones-row appended to v makes column j = D the mean of k: the denominator's vector)
//     out[j, p] = (sum_i q[i, p] kv[i, j]) / (sum_i q[i, p] kbar[i] + 1e-6) + pe[j, p]
// One CTA per (image, head), 8 warps.  Phase 1 streams k and v in 128-pixel chunks through shared memory (elu + 1 and the optional bias
// are applied on the way in, values rounded to 16 bits exactly where the reference's autocast graph holds 16-bit tensors) and accumulates
// kv as mma.sync.m16n8k16 tiles: A = k rows (K = pixels, contiguous: row-major A straight from NCHW), B = v rows (col-major B straight from
// NCHW); a warp owns whole output tiles and walks all pixels, so there is no cross-warp reduction and the result is deterministic.
// Phase 2 streams q: A = q^T through ldmatrix.trans, B = kv^T (16-bit, as the reference's `kv` tensor is), one 16-pixel m-tile per warp;
// the quotient is staged in shared memory and leaves with the lanes along the pixels (+ pe: a tensor, or the depthwise 3x3 conv of v).
// K = D = 20..40 is far below a tcgen05 tile (tools/tc_probe.cu: a 128 x N x 16 tcgen05.mma is operand-fetch bound at any N), hence mma.sync.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace recnext {

namespace {

template <typename T> struct LH;
template <> struct LH<__nv_bfloat16> {
    static __device__ __forceinline__ float to_f(unsigned short u) { return __uint_as_float((uint32_t)u << 16); }
    static __device__ __forceinline__ unsigned short from_f(float v) { __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<unsigned short*>(&h); }
    static constexpr unsigned short one = 0x3f80;
    static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};
template <> struct LH<__half> {
    static __device__ __forceinline__ float to_f(unsigned short u) { return __half2float(*reinterpret_cast<__half*>(&u)); }
    static __device__ __forceinline__ unsigned short from_f(float v) { __half h = __float2half_rn(v); return *reinterpret_cast<unsigned short*>(&h); }
    static constexpr unsigned short one = 0x3c00;
    static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};

__device__ __forceinline__ void la_ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void la_ldsm4t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void la_ldsm2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ float la_elu1(float v) { return v > 0.f ? v + 1.f : expf(v); }   // elu(v) + 1

constexpr int kCH = 128;               // pixels per chunk
constexpr int kCP = (kCH + 8) * 2;     // bytes per shared-memory row: 272 = 17 x 16 (conflict-free ldmatrix rows)

// D = head dim; DP = D rounded up to 16 (M of phase 1, K of phase 2); NP = (D + 1) rounded up to 8 (N of both phases: v rows + the ones row)
template <typename T, int D>
__global__ void __launch_bounds__(256) recnext_linattn_mma_kernel(const T* __restrict__ q_pre, const T* __restrict__ k_pre, const float* __restrict__ qbias,
                                                                  const float* __restrict__ kbias, const T* __restrict__ v, const T* __restrict__ pe,
                                                                  const float* __restrict__ pew, const float* __restrict__ peb, int pw,
                                                                  T* __restrict__ out, int heads, int n) {
    constexpr int DP = (D + 15) / 16 * 16, NP = (D + 1 + 7) / 8 * 8;
    constexpr int MT = DP / 16, NT = NP / 8, NTILE = MT * NT, TPW = (NTILE + 7) / 8;   // phase-1 output tiles, tiles per warp
    constexpr int KVP = (DP + 8) * 2;                   // bytes per row of kv^T [NP][DP]: (DP / 8 + 1) x 16, odd multiple of 16
    __shared__ __align__(16) unsigned char s_a[DP * kCP];    // k chunk (phase 1) / q chunk (phase 2): [i][pixel]
    __shared__ __align__(16) unsigned char s_b[NP * kCP];    // v chunk + ones row (phase 1): [j][pixel]; output staging (phase 2)
    __shared__ __align__(16) unsigned char s_kv[NP * KVP];   // kv^T, 16-bit, scaled by 1/n: [j][i]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;
    const int dim = heads * D;
    const long img = (k_pre == q_pre + (long)dim * n) ? 2l * dim * n : (long)dim * n;   // (one [B, 2 dim, n] tensor, or two [B, dim, n] tensors)
    const unsigned short* qp = reinterpret_cast<const unsigned short*>(q_pre) + (long)b * img + (long)h * D * n;
    const unsigned short* kp = reinterpret_cast<const unsigned short*>(k_pre) + (long)b * img + (long)h * D * n;
    const unsigned short* vp = reinterpret_cast<const unsigned short*>(v) + ((long)b * dim + h * D) * n;
    const float* qbp = qbias ? qbias + h * D : nullptr;
    const float* kbp = kbias ? kbias + h * D : nullptr;
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s_a), sbb = (uint32_t)__cvta_generic_to_shared(s_b), skv = (uint32_t)__cvta_generic_to_shared(s_kv);
    unsigned short* a16 = reinterpret_cast<unsigned short*>(s_a);
    unsigned short* b16 = reinterpret_cast<unsigned short*>(s_b);
    constexpr int RP = kCP / 2;                            // elements per shared-memory row

    // rows that are padding stay constant: zero k / q rows i >= D, v rows j > D zero, row j = D ones
    for (int i = tid; i < (DP - D) * RP; i += 256) a16[D * RP + i] = 0;
    for (int i = tid; i < (NP - D) * RP; i += 256) b16[D * RP + i] = (i < RP) ? LH<T>::one : (unsigned short)0;
    // a chunk row (one channel, 128 pixels) -> shared memory, transformed; pixels past n are zero (also in the ones row: handled by `cn` below)
    // Only the first `cw` pixels of a row are touched (cw = cn rounded up to a power of two >= 16: the MMAs read whole 16-pixel steps, and
    // shifts replace divisions): small planes (7 x 7, 4 x 4) do not pay for 128-pixel rows.
    const bool flat = n <= kCH && ((D * n) & 7) == 0 && (n & 3) != 0 && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(pe)) & 15) == 0;
    auto load_rows = [&](const unsigned short* src, unsigned short* dst, const float* bias, bool act, int c0, int cn) {
        int cwl = 4;
        while ((1 << cwl) < cn) ++cwl;
        const int vw = (n & 7) == 0 ? 8 : ((n & 3) == 0 ? 4 : 1);   // pixels per global access (rows are n elements apart; chunks start at multiples of 128)
        if (vw > 1) {
            const int sl = cwl - (vw == 8 ? 3 : 2);             // log2(segments per row)
            for (int i = tid; i < (D << sl); i += 256) {
                const int row = i >> sl, px = vw * (i - (row << sl));
                uint32_t w[4] = {0u, 0u, 0u, 0u};
                if (px < cn) {
                    const unsigned short* gsrc = src + (long)row * n + c0 + px;
                    if (vw == 8) { const uint4 u = __ldg(reinterpret_cast<const uint4*>(gsrc)); w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w; }
                    else { const uint2 u = __ldg(reinterpret_cast<const uint2*>(gsrc)); w[0] = u.x; w[1] = u.y; }
                    if (act) {
                        const float bb = bias ? bias[row] : 0.f;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (e >= vw / 2) break;
                            const unsigned short lo = LH<T>::from_f(la_elu1(LH<T>::to_f((unsigned short)(w[e] & 0xffffu)) + bb));
                            const unsigned short hi = LH<T>::from_f(la_elu1(LH<T>::to_f((unsigned short)(w[e] >> 16)) + bb));
                            w[e] = (uint32_t)lo | ((uint32_t)hi << 16);
                        }
                    }
                    // (n % vw == 0 and chunks start at multiples of 128: cn is a multiple of vw, a segment is whole or empty)
                }
                if (vw == 8) *reinterpret_cast<uint4*>(dst + row * RP + px) = make_uint4(w[0], w[1], w[2], w[3]);
                else *reinterpret_cast<uint2*>(dst + row * RP + px) = make_uint2(w[0], w[1]);
            }
        } else if (flat) {
            // odd plane sizes that fit one chunk (7 x 7): the [D x n] block of this (image, head) is CONTIGUOUS and 16-byte aligned, so it is
            // read as flat 16-byte vectors and scattered into the rows (one division per vector instead of 2-byte loads with one each)
            const int cw = 1 << cwl;
            for (int i = tid; i < D * (cw - n); i += 256) { const int row = i / (cw - n); dst[row * RP + n + (i - row * (cw - n))] = 0; }
            for (int i = tid; i < (D * n) / 8; i += 256) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
                int row = (8 * i) / n, px = 8 * i - row * n;
                float bb = (act && bias) ? bias[row] : 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    unsigned short val = (unsigned short)((w[e >> 1] >> (16 * (e & 1))) & 0xffffu);
                    if (act) val = LH<T>::from_f(la_elu1(LH<T>::to_f(val) + bb));
                    dst[row * RP + px] = val;
                    if (++px == n) { px = 0; ++row; if (act && bias && row < D) bb = bias[row]; }
                }
            }
        } else {
            for (int i = tid; i < (D << cwl); i += 256) {
                const int row = i >> cwl, px = i - (row << cwl);
                unsigned short e = 0;
                if (px < cn) {
                    e = __ldg(src + (long)row * n + c0 + px);
                    if (act) e = LH<T>::from_f(la_elu1(LH<T>::to_f(e) + (bias ? bias[row] : 0.f)));
                }
                dst[row * RP + px] = e;
            }
        }
    };

    // ---- phase 1: kv = k v'^T over all pixels; warp w owns output tiles w, w + 8, ... (tile = (m-tile, n-tile))
    float acc1[TPW][4];
#pragma unroll
    for (int e = 0; e < TPW; ++e) acc1[e][0] = acc1[e][1] = acc1[e][2] = acc1[e][3] = 0.f;
    const int lrow = lane & 7, lmat = lane >> 3;
    for (int c0 = 0; c0 < n; c0 += kCH) {
        const int cn = (n - c0) < kCH ? (n - c0) : kCH;
        __syncthreads();                                   // the previous chunk's fragments have been read
        load_rows(kp, a16, kbp, true, c0, cn);
        load_rows(vp, b16, nullptr, false, c0, cn);
        if (cn < kCH) for (int i = tid + cn; i < kCH; i += 256) b16[D * RP + i] = 0;   // the ones row ends with the image
        __syncthreads();
        const int ksteps = (cn + 15) / 16;
#pragma unroll
        for (int e = 0; e < TPW; ++e) {
            const int tile = warp + 8 * e;
            if (tile < NTILE) {
                const int mt = tile / NT, nt = tile - mt * NT;
                // A (k rows): matrices (i 0-7, p 0-7), (i 8-15, p 0-7), (i 0-7, p 8-15), (i 8-15, p 8-15); B (v rows): (j 0-7, p 0-7), (j 0-7, p 8-15)
                const uint32_t aaddr = sa + (uint32_t)((mt * 16 + (lmat & 1) * 8 + lrow) * kCP) + (uint32_t)((lmat >> 1) * 16);
                const uint32_t baddr = sbb + (uint32_t)((nt * 8 + lrow) * kCP) + (uint32_t)((lmat & 1) * 16);
                for (int ks = 0; ks < ksteps; ++ks) {
                    uint32_t af[4], b0, b1;
                    la_ldsm4(af, aaddr + (uint32_t)(ks * 32));
                    la_ldsm2(b0, b1, baddr + (uint32_t)(ks * 32));
                    LH<T>::mma(acc1[e], af, b0, b1);
                }
            }
        }
    }
    // kv^T[j][i] = acc / n as 16-bit (the reference's kv is a 16-bit tensor under autocast); fragment: rows i = g, g + 8; columns j = 2 t, 2 t + 1
    {
        const float inv_n = 1.f / (float)n;
        unsigned short* kv16 = reinterpret_cast<unsigned short*>(s_kv);
#pragma unroll
        for (int e = 0; e < TPW; ++e) {
            const int tile = warp + 8 * e;
            if (tile < NTILE) {
                const int mt = tile / NT, nt = tile - mt * NT;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int i = mt * 16 + g + 8 * (c >> 1), j = nt * 8 + 2 * t4 + (c & 1);
                    kv16[j * (KVP / 2) + i] = LH<T>::from_f(acc1[e][c] * inv_n);
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: per 128-pixel chunk, warp w takes pixels 16 w .. 16 w + 15: num[p][j] = sum_i q[i][p] kv[i][j]; column j = D is the denominator
    const unsigned short* pep = pe ? reinterpret_cast<const unsigned short*>(pe) + ((long)b * dim + h * D) * n : nullptr;
    unsigned short* op = reinterpret_cast<unsigned short*>(out) + ((long)b * dim + h * D) * n;
    for (int c0 = 0; c0 < n; c0 += kCH) {
        const int cn = (n - c0) < kCH ? (n - c0) : kCH;
        __syncthreads();                                   // the previous chunk's staging buffer has been written out
        load_rows(qp, a16, qbp, true, c0, cn);
        __syncthreads();
        if (16 * warp < cn) {
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
            // A = q^T by ldmatrix.trans of the blocks (i 0-7, p 0-7), (i 0-7, p 8-15), (i 8-15, p 0-7), (i 8-15, p 8-15)
            const uint32_t aaddr = sa + (uint32_t)(((lmat >> 1) * 8 + lrow) * kCP) + (uint32_t)((16 * warp + (lmat & 1) * 8) * 2);
            const uint32_t baddr = skv + (uint32_t)(lrow * KVP) + (uint32_t)((lmat & 1) * 16);
#pragma unroll
            for (int ks = 0; ks < MT; ++ks) {
                uint32_t af[4];
                la_ldsm4t(af, aaddr + (uint32_t)(ks * 16 * kCP));
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    uint32_t b0, b1;
                    la_ldsm2(b0, b1, baddr + (uint32_t)(nt * 8 * KVP) + (uint32_t)(ks * 32));
                    LH<T>::mma(acc[nt], af, b0, b1);
                }
            }
            // denominator: column D of rows g / g + 8 lives in the quad's lane t = (D % 8) / 2
            constexpr int ND = D / 8, CD = D % 8;
            const int src = (lane & ~3) | (CD >> 1);
            const float d0 = __shfl_sync(0xffffffffu, acc[ND][CD & 1], src), d1 = __shfl_sync(0xffffffffu, acc[ND][2 + (CD & 1)], src);
            const float r0 = 1.f / (d0 + 1e-6f), r1 = 1.f / (d1 + 1e-6f);
            // stage the quotient as [j][pixel] (the chunk's v rows are dead: phase 1 is over)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = nt * 8 + 2 * t4 + (c & 1), p = 16 * warp + g + 8 * (c >> 1);
                    if (j < D) b16[j * RP + p] = LH<T>::from_f(acc[nt][c] * ((c >> 1) ? r1 : r0));
                }
        }
        __syncthreads();
        // out = staged + pe, lanes along the pixels
        auto pe_of = [&](int j, int pidx, int yy, int xx) {   // pe[j, pixel]: a tensor, or the depthwise 3x3 conv (+ bias) of v's plane j at (yy, xx)
            float add = 0.f;
            if (pep) add = LH<T>::to_f(pep[(long)j * n + pidx]);
            if (pew) {
                const int ph = n / pw;
                const float* wj = pew + (long)(h * D + j) * 9;
                const unsigned short* vj = vp + (long)j * n;
                float a9 = peb ? peb[h * D + j] : 0.f;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int y2 = yy + dy;
                    if (y2 < 0 || y2 >= ph) continue;
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int x2 = xx + dx;
                        if (x2 >= 0 && x2 < pw) a9 = fmaf(wj[(dy + 1) * 3 + dx + 1], LH<T>::to_f(vj[(long)y2 * pw + x2]), a9);
                    }
                }
                add += a9;
            }
            return add;
        };
        if (flat) {   // (single chunk) the [D x n] output block is contiguous: 16-byte stores of 8 flat elements
            const int pwd = pew ? pw : n;
            for (int i = tid; i < (D * n) / 8; i += 256) {
                int j = (8 * i) / n, px = 8 * i - j * n;
                int yy = px / pwd, xx = px - yy * pwd;
                uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float val = LH<T>::to_f(b16[j * RP + px]) + pe_of(j, px, yy, xx);
                    w[e >> 1] |= (uint32_t)LH<T>::from_f(val) << (16 * (e & 1));
                    if (++xx == pwd) { xx = 0; ++yy; }
                    if (++px == n) { px = 0; ++j; yy = 0; xx = 0; }
                }
                reinterpret_cast<uint4*>(op)[i] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
            int owl = 4;
            while ((1 << owl) < cn) ++owl;
            for (int i = tid; i < (D << owl); i += 256) {
                const int j = i >> owl, px = i - (j << owl);
                if (px >= cn) continue;
                const int pidx = c0 + px, pwd = pew ? pw : n;
                const int yy = pidx / pwd, xx = pidx - yy * pwd;
                op[(long)j * n + pidx] = LH<T>::from_f(LH<T>::to_f(b16[j * RP + px]) + pe_of(j, pidx, yy, xx));
            }
        }
    }
}

template <typename T>
cudaError_t la_mma_launch_t(int B, int heads, int d, int n, const void* q, const void* k, const float* qb, const float* kb, const void* v, const void* pe,
                            const float* pew, const float* peb, int pw, void* out, cudaStream_t s) {
    const int grid = B * heads;
#define LAM_CASE(DD) case DD: recnext_linattn_mma_kernel<T, DD><<<grid, 256, 0, s>>>((const T*)q, (const T*)k, qb, kb, (const T*)v, (const T*)pe, pew, peb, pw, (T*)out, heads, n); break;
    switch (d) {
        LAM_CASE(4) LAM_CASE(8) LAM_CASE(16) LAM_CASE(20) LAM_CASE(24) LAM_CASE(28) LAM_CASE(32) LAM_CASE(40)
        default: return cudaErrorInvalidValue;
    }
#undef LAM_CASE
    return cudaGetLastError();
}

}  // namespace

// 16-bit activations only (dtype 1 = bf16, 2 = fp16); same contract as linattn_launch
cudaError_t linattn_mma_launch(int B, int heads, int d, int n, int dtype, const void* q, const void* k, const float* qb, const float* kb, const void* v,
                               const void* pe, const float* pew, const float* peb, int pw, void* out, cudaStream_t stream) {
    if (dtype == 1) return la_mma_launch_t<__nv_bfloat16>(B, heads, d, n, q, k, qb, kb, v, pe, pew, peb, pw, out, stream);
    return la_mma_launch_t<__half>(B, heads, d, n, q, k, qb, kb, v, pe, pew, peb, pw, out, stream);
}

}  // namespace recnext
