// dwdown.cu — the depthwise 7x7 stride-2 convolution of a RecNeXt `Downsample` block with channel multiplier 2 and the
// eval-mode BatchNorm that follows it folded into (w, b):   reference model/recnext.py:134-146
//     out[n, 2c + m, i, j] = b[2c + m] + sum_{r,s} w[2c + m, r, s] * x[n, c, 2i + r - 3, 2j + s - 3],   m = 0, 1
// (nn.Conv2d(C, 2C, 7, padding=3, stride=2, groups=C) followed by BatchNorm2d(2C)).  PyTorch dispatches this to its generic
// depthwise kernel: 0.62 ms per launch at [256, 64, 56, 56], 16 % of the RecNeXt-M3 inference step
// (profiles/r1_e_launches_bench_step.txt).  Here a persistent CTA takes groups of PP input planes: a plane lands in shared memory as
// padded fp32 (zero border) with its columns SPLIT BY RESIDUE MOD 8, each thread produces a 2 x 4 block of BOTH output channels from
// a 9 x 13 window (117 shared loads, conflict-free: a thread's 13 columns are 8 bx + s, so lanes with consecutive bx read consecutive
// words of residue plane s & 7) and 392 packed FFMA2 (the two output channels of an input value are one fp32x2 accumulator); the
// filters come from shared memory as 16-byte broadcasts.  x read once, out written once.
#include <cuda_runtime.h>
#include "devcfg.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace recnext {

template <typename T> __device__ __forceinline__ float dd_to_f(T v);
template <> __device__ __forceinline__ float dd_to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float dd_to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float dd_to_f<float>(float v) { return v; }
template <typename T> __device__ __forceinline__ T dd_from_f(float v);
template <> __device__ __forceinline__ __nv_bfloat16 dd_from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half dd_from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float dd_from_f<float>(float v) { return v; }


namespace {

constexpr int kWRow = 16;                 // floats per filter row in shared memory: [7 taps + pad][2 channels]
constexpr int kWSlot = 7 * kWRow;         // 112 floats of filters in front of every plane slot

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <typename T> struct Pack4;
template <> struct Pack4<__nv_bfloat16> {
    static __device__ __forceinline__ uint2 pack(float a, float b, float c, float d) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
        return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
};
template <> struct Pack4<__half> {
    static __device__ __forceinline__ uint2 pack(float a, float b, float c, float d) {
        __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
        return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
};
template <> struct Pack4<float> {
    static __device__ __forceinline__ uint2 pack(float, float, float, float) { return make_uint2(0u, 0u); }   // (unused: fp32 rows are stored element-wise)
};

struct DwPlan {
    int C, H, W, Ho, Wo;
    int rows, cw, rs;         // padded rows (H + 8), words per row of a residue plane, words per residue plane (rows * cw)
    int slot;                 // floats per plane slot: filters + 8 residue planes
    int bh, bwn, items;       // 2 x 4 output blocks per plane
    int PP;                   // planes per group
    int vw, vpr, lpr_shift;   // pixels per global load (8 / 4 / 2 / 1), loads per row, log2(lanes per row)
    long nplanes, ngroups;
};

template <typename T>
__global__ void __launch_bounds__(256) recnext_dwdown_kernel(const __grid_constant__ DwPlan p, const T* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ b, T* __restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = p.C, H = p.H, W = p.W, Ho = p.Ho, Wo = p.Wo, cw = p.cw, rs = p.rs, slot = p.slot, PP = p.PP;
    for (int i = tid; i < PP * slot; i += 256) sm[i] = 0.f;    // borders stay zero for the whole kernel: only interiors are rewritten
    const int lpr = 1 << p.lpr_shift, rpw = 32 >> p.lpr_shift;  // lanes per row, rows per warp pass of the loader
    const int lj = lane & (lpr - 1), lr = lane >> p.lpr_shift;
    const long HW = (long)H * W;
    for (long grp = blockIdx.x; grp < p.ngroups; grp += gridDim.x) {
        const long plane0 = grp * PP;
        __syncthreads();   // the previous group's windows have been read (first pass: the zero fill is done)
        // ---- filters of the group's planes: [fr][tap][2 channels]
        for (int i = tid; i < PP * 98; i += 256) {
            const int pl = i / 98, e = i - pl * 98, m = e / 49, k = e - 49 * m, fr = k / 7, tap = k - 7 * fr;
            if (plane0 + pl < p.nplanes) sm[pl * slot + fr * kWRow + tap * 2 + m] = __ldg(w + (long)(2 * ((plane0 + pl) % C) + m) * 49 + k);
        }
        // ---- planes: lanes along the row in vectors of `vw` pixels; a pixel at padded column pc goes to residue plane pc & 7, word pc >> 3
        for (int gr = warp * rpw + lr; gr < PP * H; gr += 8 * rpw) {
            const int pl = gr / H, r = gr - pl * H;
            if (plane0 + pl >= p.nplanes) continue;
            for (int j = lj; j < p.vpr; j += lpr) {   // (more than 32 vectors per row: the lanes stride along the row)
                const T* src = x + (plane0 + pl) * HW + (long)r * W + j * p.vw;
                float* dst = sm + pl * slot + kWSlot + (r + 3) * cw;
                float v[8];
                if (sizeof(T) == 2 && p.vw == 8) {
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
                    const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        T lo, hi;
                        *reinterpret_cast<unsigned short*>(&lo) = (unsigned short)(q[e] & 0xffffu); *reinterpret_cast<unsigned short*>(&hi) = (unsigned short)(q[e] >> 16);
                        v[2 * e] = dd_to_f<T>(lo); v[2 * e + 1] = dd_to_f<T>(hi);
                    }
                } else if (sizeof(T) == 2 && p.vw == 4) {
                    const uint2 u = __ldg(reinterpret_cast<const uint2*>(src));
                    const uint32_t q[2] = {u.x, u.y};
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        T lo, hi;
                        *reinterpret_cast<unsigned short*>(&lo) = (unsigned short)(q[e] & 0xffffu); *reinterpret_cast<unsigned short*>(&hi) = (unsigned short)(q[e] >> 16);
                        v[2 * e] = dd_to_f<T>(lo); v[2 * e + 1] = dd_to_f<T>(hi);
                    }
                } else {
                    for (int e = 0; e < p.vw; ++e) v[e] = dd_to_f<T>(src[e]);
                }
                const int pc0 = j * p.vw + 3;
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (e < p.vw) { const int pc = pc0 + e; dst[(pc & 7) * rs + (pc >> 3)] = v[e]; }
            }
        }
        __syncthreads();
        // ---- 2 x 4 output blocks x 2 channels
        for (int it = tid; it < PP * p.items; it += 256) {
            const int pl = it / p.items, blk = it - pl * p.items;
            const long plane = plane0 + pl;
            if (plane >= p.nplanes) break;
            const int c = (int)(plane % C);
            const int by = blk / p.bwn, bx = blk - by * p.bwn;
            const float* ws = sm + pl * slot;
            const float* pr = ws + kWSlot + (4 * by) * cw + bx;     // residue plane 0, window row 0, word bx
            const float2 bias = make_float2(__ldg(b + 2 * c), __ldg(b + 2 * c + 1));
            float2 a[2][4];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 4; ++dx) a[dy][dx] = bias;
#pragma unroll
            for (int r = 0; r < 9; ++r) {
                float v[13];
#pragma unroll
                for (int s2 = 0; s2 < 13; ++s2) v[s2] = pr[(s2 & 7) * rs + r * cw + (s2 >> 3)];
#pragma unroll
                for (int dy = 0; dy < 2; ++dy) {
                    const int fr = r - 2 * dy;   // filter row for output row 2 by + dy
                    if (fr < 0 || fr > 6) continue;
                    const float4* wr = reinterpret_cast<const float4*>(ws + fr * kWRow);
                    const float4 w01 = wr[0], w23 = wr[1], w45 = wr[2], w6 = wr[3];
                    const float2 wt[7] = {make_float2(w01.x, w01.y), make_float2(w01.z, w01.w), make_float2(w23.x, w23.y), make_float2(w23.z, w23.w),
                                          make_float2(w45.x, w45.y), make_float2(w45.z, w45.w), make_float2(w6.x, w6.y)};
#pragma unroll
                    for (int tap = 0; tap < 7; ++tap)
#pragma unroll
                        for (int dx = 0; dx < 4; ++dx) a[dy][dx] = ffma2(wt[tap], make_float2(v[tap + 2 * dx], v[tap + 2 * dx]), a[dy][dx]);
                }
            }
            const int oy = 2 * by, ox = 4 * bx;
            T* o0 = out + ((plane / C) * 2 * C + 2 * c) * (long)(Ho * Wo);
            T* o1 = o0 + (long)Ho * Wo;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                if (oy + dy >= Ho) continue;
                const long ro = (long)(oy + dy) * Wo + ox;
                if (sizeof(T) == 2 && (Wo & 3) == 0) {
                    *reinterpret_cast<uint2*>(o0 + ro) = Pack4<T>::pack(a[dy][0].x, a[dy][1].x, a[dy][2].x, a[dy][3].x);
                    *reinterpret_cast<uint2*>(o1 + ro) = Pack4<T>::pack(a[dy][0].y, a[dy][1].y, a[dy][2].y, a[dy][3].y);
                } else {
#pragma unroll
                    for (int dx = 0; dx < 4; ++dx)
                        if (ox + dx < Wo) { o0[ro + dx] = dd_from_f<T>(a[dy][dx].x); o1[ro + dx] = dd_from_f<T>(a[dy][dx].y); }
                }
            }
        }
    }
}


// ---- 16-bit activations with even W: the planes stay 16-bit in shared memory, as 32-bit words (two pixels) split by word residue mod 4,
// copied there by 4-byte cp.async (no register staging: a thread issues its copies of the NEXT group of planes and goes on computing
// this one: two buffers).  Left padding is 4 columns, so global pixel pairs are also padded-column pairs; a thread's 13 window
// columns 8 bx + 1 .. 8 bx + 13 are the words 4 bx + k, k = 0 .. 6: residue plane k & 3, word bx + (k >> 2) -- conflict-free 32-bit loads.
struct DwPlanH {
    int C, H, W, Ho, Wo;
    int rows, cw, rs;         // padded rows (H + 8), words per row of a residue plane, words per residue plane
    int slot;                 // 32-bit words per plane slot: 112 filter floats + 4 residue planes
    int bh, bwn, items;
    int PP;
    int wpr, lpr_shift;       // words per row (W / 2), log2(lanes per row)
    long nplanes, ngroups;
};

template <typename T> __device__ __forceinline__ float2 dd_unpack(uint32_t w);
template <> __device__ __forceinline__ float2 dd_unpack<__nv_bfloat16>(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
template <> __device__ __forceinline__ float2 dd_unpack<__half>(uint32_t w) { return __half22float2(*reinterpret_cast<__half2*>(&w)); }

__device__ __forceinline__ void dd_cp_async4(uint32_t dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void dd_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void dd_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T>
__global__ void __launch_bounds__(256) recnext_dwdown16_kernel(const __grid_constant__ DwPlanH p, const T* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ b, T* __restrict__ out) {
    extern __shared__ __align__(16) uint32_t smw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = p.C, H = p.H, W = p.W, Ho = p.Ho, Wo = p.Wo, cw = p.cw, rs = p.rs, slot = p.slot, PP = p.PP;
    for (int i = tid; i < 2 * PP * slot; i += 256) smw[i] = 0u;   // borders stay zero for the whole kernel: only interiors are rewritten
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smw);
    const int lpr = 1 << p.lpr_shift, rpw = 32 >> p.lpr_shift;
    const int lj = lane & (lpr - 1), lr = lane >> p.lpr_shift;
    const long HW = (long)H * W;
    // the copies of one group: filters by plain stores (few), planes by cp.async
    auto issue = [&](long grp, int buf) {
        const long plane0 = grp * PP;
        uint32_t* bufw = smw + (size_t)buf * PP * slot;
        const uint32_t bb = sbase + (uint32_t)((size_t)buf * PP * slot * 4);
        const int c0 = (int)((unsigned long long)plane0 % (unsigned)C);   // channel of the group's first plane; the others follow by increments
        const int npl = (int)((p.nplanes - plane0 < PP) ? (p.nplanes - plane0) : PP);
        for (int i = tid; i < npl * 98; i += 256) {
            const int pl = i / 98, e = i - pl * 98, m = e / 49, k = e - 49 * m, fr = k / 7, tap = k - 7 * fr;
            int c = c0 + pl;
            while (c >= C) c -= C;
            reinterpret_cast<float*>(bufw)[pl * slot + fr * kWRow + tap * 2 + m] = __ldg(w + (long)(2 * c + m) * 49 + k);
        }
        // rows of all planes of the group, flattened: (pl, r) advance by increments (no division per row)
        {
            const int step = 8 * rpw;
            int r = warp * rpw + lr, pl = 0;
            while (r >= H) { r -= H; ++pl; }
            while (pl < npl) {
                const T* srow = x + (plane0 + pl) * HW + (long)r * W;
                const uint32_t drow = bb + (uint32_t)((pl * slot + kWSlot + (r + 3) * cw) * 4);
                for (int j = lj; j < p.wpr; j += lpr) {
                    const int wd = j + 2;                           // word index in the padded row (4 columns of left padding)
                    dd_cp_async4(drow + (uint32_t)(((wd & 3) * rs + (wd >> 2)) * 4), srow + 2 * j);
                }
                r += step;
                while (r >= H) { r -= H; ++pl; }
            }
        }
        dd_cp_commit();
    };
    int buf = 0;
    if ((long)blockIdx.x < p.ngroups) issue(blockIdx.x, 0);
    for (long grp = blockIdx.x; grp < p.ngroups; grp += gridDim.x, buf ^= 1) {
        const long plane0 = grp * PP;
        if (grp + gridDim.x < p.ngroups) { issue(grp + gridDim.x, buf ^ 1); dd_cp_wait<1>(); }
        else dd_cp_wait<0>();
        __syncthreads();   // this group's planes and filters are in place
        const uint32_t* bufw = smw + (size_t)buf * PP * slot;
        const unsigned c0 = (unsigned)((unsigned long long)plane0 % (unsigned)C);
        const long n0 = plane0 / C;
        for (int it = tid; it < PP * p.items; it += 256) {
            const int pl = it / p.items, blk = it - pl * p.items;
            if (plane0 + pl >= p.nplanes) break;
            unsigned c = c0 + (unsigned)pl;
            long n = n0;
            while (c >= (unsigned)C) { c -= (unsigned)C; ++n; }
            const int by = blk / p.bwn, bx = blk - by * p.bwn;
            const float* ws = reinterpret_cast<const float*>(bufw + pl * slot);
            const uint32_t* pr = bufw + pl * slot + kWSlot + (4 * by) * cw + bx;
            const float2 bias = make_float2(__ldg(b + 2 * c), __ldg(b + 2 * c + 1));
            float2 a[2][4];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 4; ++dx) a[dy][dx] = bias;
#pragma unroll
            for (int r = 0; r < 9; ++r) {
                float v[14];     // padded columns 8 bx + 0 .. 8 bx + 13; the window is columns 1 .. 13
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const float2 f = dd_unpack<T>(pr[(k & 3) * rs + r * cw + (k >> 2)]);
                    v[2 * k] = f.x; v[2 * k + 1] = f.y;
                }
#pragma unroll
                for (int dy = 0; dy < 2; ++dy) {
                    const int fr = r - 2 * dy;
                    if (fr < 0 || fr > 6) continue;
                    const float4* wr = reinterpret_cast<const float4*>(ws + fr * kWRow);
                    const float4 w01 = wr[0], w23 = wr[1], w45 = wr[2], w6 = wr[3];
                    const float2 wt[7] = {make_float2(w01.x, w01.y), make_float2(w01.z, w01.w), make_float2(w23.x, w23.y), make_float2(w23.z, w23.w),
                                          make_float2(w45.x, w45.y), make_float2(w45.z, w45.w), make_float2(w6.x, w6.y)};
#pragma unroll
                    for (int tap = 0; tap < 7; ++tap)
#pragma unroll
                        for (int dx = 0; dx < 4; ++dx) a[dy][dx] = ffma2(wt[tap], make_float2(v[1 + tap + 2 * dx], v[1 + tap + 2 * dx]), a[dy][dx]);
                }
            }
            const int oy = 2 * by, ox = 4 * bx;
            T* o0 = out + (n * 2 * C + 2 * c) * (long)(Ho * Wo);
            T* o1 = o0 + (long)Ho * Wo;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                if (oy + dy >= Ho) continue;
                const long ro = (long)(oy + dy) * Wo + ox;
                if ((Wo & 3) == 0) {
                    *reinterpret_cast<uint2*>(o0 + ro) = Pack4<T>::pack(a[dy][0].x, a[dy][1].x, a[dy][2].x, a[dy][3].x);
                    *reinterpret_cast<uint2*>(o1 + ro) = Pack4<T>::pack(a[dy][0].y, a[dy][1].y, a[dy][2].y, a[dy][3].y);
                } else {
#pragma unroll
                    for (int dx = 0; dx < 4; ++dx)
                        if (ox + dx < Wo) { o0[ro + dx] = dd_from_f<T>(a[dy][dx].x); o1[ro + dx] = dd_from_f<T>(a[dy][dx].y); }
                }
            }
        }
        __syncthreads();   // the windows of this buffer have been read: the group after next may land in it
    }
}

static int dwdown16_launch(int B, int C, int H, int W, int dtype, const void* x, const float* w, const float* b, void* out, cudaStream_t stream, cudaError_t* err) {
    DwPlanH p{};
    p.C = C; p.H = H; p.W = W;
    p.Ho = (H - 1) / 2 + 1; p.Wo = (W - 1) / 2 + 1;
    p.bh = (p.Ho + 1) / 2; p.bwn = (p.Wo + 3) / 4; p.items = p.bh * p.bwn;
    p.rows = H + 8;
    int best_cw = p.bwn + 1, best_conf = 1 << 30;
    for (int cw = p.bwn + 1; cw <= p.bwn + 9; ++cw) {
        int cnt[32] = {0}, conf = 0;
        for (int lane = 0; lane < 32; ++lane) { const int bank = ((lane / p.bwn) * 4 * cw + (lane % p.bwn)) & 31; if (++cnt[bank] > conf) conf = cnt[bank]; }
        if (conf < best_conf) { best_conf = conf; best_cw = cw; }
    }
    p.cw = best_cw; p.rs = p.rows * p.cw;
    p.slot = (kWSlot + 4 * p.rs + 3) / 4 * 4;
    const size_t sbytes = (size_t)p.slot * 4;
    if (2 * sbytes > 227 * 1024) return 1;
    // planes per group: fill the 256 threads' rounds as evenly as possible; two buffers within ~110 KB (two CTAs per SM)
    int bestPP = 1; double best_eff = 0.0;
    for (int PP = 1; PP <= 32 && 2 * PP * sbytes <= 110 * 1024; ++PP) {
        const int work = PP * p.items, rounds = (work + 255) / 256;
        const double eff = (double)work / (256.0 * rounds);
        if (eff > best_eff + 0.02) { best_eff = eff; bestPP = PP; }
    }
    p.PP = bestPP;
    p.wpr = W / 2;
    p.lpr_shift = 0;
    while ((1 << p.lpr_shift) < p.wpr && p.lpr_shift < 5) ++p.lpr_shift;
    p.nplanes = (long)B * C;
    p.ngroups = (p.nplanes + p.PP - 1) / p.PP;
    const size_t smem = 2 * (size_t)p.PP * sbytes;
    static DeviceOnce configured = {};
    *err = rc_once_per_device(configured, [] {
        cudaError_t e = cudaFuncSetAttribute(recnext_dwdown16_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(recnext_dwdown16_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        return e;
    });
    if (*err != cudaSuccess) return 2;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    long grid = (long)rc_device_sms() * per_sm;
    if (grid > p.ngroups) grid = p.ngroups;
    if (dtype == 1) recnext_dwdown16_kernel<__nv_bfloat16><<<(unsigned)grid, 256, smem, stream>>>(p, (const __nv_bfloat16*)x, w, b, (__nv_bfloat16*)out);
    else recnext_dwdown16_kernel<__half><<<(unsigned)grid, 256, smem, stream>>>(p, (const __half*)x, w, b, (__half*)out);
    *err = cudaGetLastError();
    return *err == cudaSuccess ? 0 : 2;
}

}  // namespace

// 0 ok, 1 unsupported (plane does not fit), 2 CUDA error in *err
int dwdown_launch(int B, int C, int H, int W, int dtype, const void* x, const float* w, const float* b, void* out, cudaStream_t stream, cudaError_t* err) {
    if (dtype < 0 || dtype > 2) return 1;
    if (dtype != 0 && (W & 1) == 0 && (((uintptr_t)x) & 3) == 0) {   // 16-bit planes with even rows: the cp.async kernel (falls through when it does not fit)
        const int rc = dwdown16_launch(B, C, H, W, dtype, x, w, b, out, stream, err);
        if (rc != 1) return rc;
    }
    DwPlan p{};
    p.C = C; p.H = H; p.W = W;
    p.Ho = (H - 1) / 2 + 1; p.Wo = (W - 1) / 2 + 1;
    p.bh = (p.Ho + 1) / 2; p.bwn = (p.Wo + 3) / 4; p.items = p.bh * p.bwn;
    p.rows = H + 8;
    // words per residue-plane row: >= bwn + 1 (+1: the window's columns 8 .. 12 live one word further).  Lanes of a warp are (by, bx)
    // with bx fastest and rows 4 cw apart: take the candidate with the fewest lanes of warp 0 on one bank.
    int best_cw = p.bwn + 1, best_conf = 1 << 30;
    for (int cw = p.bwn + 1; cw <= p.bwn + 9; ++cw) {
        int cnt[32] = {0}, conf = 0;
        for (int lane = 0; lane < 32; ++lane) { const int bank = ((lane / p.bwn) * 4 * cw + (lane % p.bwn)) & 31; if (++cnt[bank] > conf) conf = cnt[bank]; }
        if (conf < best_conf) { best_conf = conf; best_cw = cw; }
    }
    p.cw = best_cw; p.rs = p.rows * p.cw;
    p.slot = (kWSlot + 8 * p.rs + 3) / 4 * 4;
    const size_t sbytes = (size_t)p.slot * sizeof(float);
    if (sbytes > 227 * 1024) return 1;
    // planes per group: fill the 256 threads' rounds as evenly as possible within ~110 KB (two CTAs per SM)
    int bestPP = 1; double best_eff = 0.0;
    for (int PP = 1; PP <= 32 && PP * sbytes <= 110 * 1024; ++PP) {
        const int work = PP * p.items, rounds = (work + 255) / 256;
        const double eff = (double)work / (256.0 * rounds);
        if (eff > best_eff + 0.02) { best_eff = eff; bestPP = PP; }
    }
    p.PP = bestPP;
    const int esz = dtype == 0 ? 4 : 2;
    p.vw = (esz == 2 && W % 8 == 0) ? 8 : ((esz == 2 && W % 4 == 0) ? 4 : 1);
    if ((((uintptr_t)x) & 15) != 0) p.vw = 1;
    p.vpr = W / p.vw;
    p.lpr_shift = 0;
    while ((1 << p.lpr_shift) < p.vpr && p.lpr_shift < 5) ++p.lpr_shift;

    p.nplanes = (long)B * C;
    p.ngroups = (p.nplanes + p.PP - 1) / p.PP;
    const size_t smem = (size_t)p.PP * sbytes;
    static DeviceOnce configured = {};
    *err = rc_once_per_device(configured, [] {
        cudaError_t e = cudaFuncSetAttribute(recnext_dwdown_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(recnext_dwdown_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(recnext_dwdown_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        return e;
    });
    if (*err != cudaSuccess) return 2;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    long grid = (long)rc_device_sms() * per_sm;
    if (grid > p.ngroups) grid = p.ngroups;
    if (dtype == 0) recnext_dwdown_kernel<float><<<(unsigned)grid, 256, smem, stream>>>(p, (const float*)x, w, b, (float*)out);
    else if (dtype == 1) recnext_dwdown_kernel<__nv_bfloat16><<<(unsigned)grid, 256, smem, stream>>>(p, (const __nv_bfloat16*)x, w, b, (__nv_bfloat16*)out);
    else recnext_dwdown_kernel<__half><<<(unsigned)grid, 256, smem, stream>>>(p, (const __half*)x, w, b, (__half*)out);
    *err = cudaGetLastError();
    return *err == cudaSuccess ? 0 : 2;
}

}  // namespace recnext
