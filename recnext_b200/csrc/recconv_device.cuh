// recconv_device.cuh — CUDA execution context of the RecConv schedules: barriers, TMA bulk copies
// (cp.async.bulk, SASS UBLKCP) of whole plane groups, warp-shuffle weight-gradient reduction, and the
// __global__ entry.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include "recconv_body.cuh"
#include "devcfg.h"

namespace recnext {

__device__ __forceinline__ uint32_t rc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void rc_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void rc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool rc_mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, 16-byte aligned)
__device__ __forceinline__ void rc_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     rc_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
// shared -> global bulk copy, tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void rc_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(rc_smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void rc_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void rc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// plain cooperative copy for plane groups that are not 16-byte aligned (ragged shapes)
__device__ __forceinline__ void rc_coop_copy(void* dst, const void* src, long bytes, int tid, int nthreads) {
    const uintptr_t a = (uintptr_t)dst | (uintptr_t)src | (uintptr_t)bytes;
    if ((a & 15) == 0) {
        const uint4* s = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
        for (long i = tid; i < bytes / 16; i += nthreads) d[i] = s[i];
    } else if ((a & 3) == 0) {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        for (long i = tid; i < bytes / 4; i += nthreads) d[i] = s[i];
    } else {
        const uint16_t* s = reinterpret_cast<const uint16_t*>(src);
        uint16_t* d = reinterpret_cast<uint16_t*>(dst);
        for (long i = tid; i < bytes / 2; i += nthreads) d[i] = s[i];
    }
}

// transposed butterfly: after the call lane `r` of every GG-lane group holds, in v[j], the group total of element
// base(r) + j, base(r) = r * (32 / GG) (elements padded to 32).  31 shuffles for GG = 32 instead of 5 per element.
template <int GG, int N>
__device__ __forceinline__ void rc_group_reduce(float (&v)[32], const float (&acc)[N], int r) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i < N ? acc[i] : 0.f;
    int cnt = 32;
#pragma unroll
    for (int s = GG / 2; s > 0; s >>= 1) {
        const int half = cnt / 2;
        const bool upper = (r & s) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < half) {
                const float lo = v[i], hi = v[i + half];
                const float send = upper ? lo : hi;
                const float keep = upper ? hi : lo;
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
            }
        }
        cnt = half;
    }
}

struct DeviceCtx {
    int tid, T, use_tma, g, u, ul, unit_lanes;
    uint32_t bar, phase;

    ThreadPos pos;  // computed once per thread
    template <class F> __device__ __forceinline__ void run_all(F f) { f(tid); }
    template <class F> __device__ __forceinline__ void run(int n, F f) { if (ul < n) f(pos); }
    __device__ __forceinline__ void cta_sync() { __syncthreads(); }
    // Units never interact.  A warp-resident unit needs only __syncwarp; a multi-warp team uses one named barrier
    // per participant count (64 / 128 / 256 leading lanes), so warps that skip a stage never touch its barrier.
    __device__ __forceinline__ void sync(int n) {
        if (unit_lanes <= 32 || n <= 32) { if (ul < 32) __syncwarp(); return; }
        if (ul < n) {
            const int id = 1 + u * 3 + (n == 64 ? 0 : (n == 128 ? 1 : 2));
            asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
        }
    }
    __device__ __forceinline__ void sync_unit() { sync(unit_lanes); }

    template <class F> __device__ __forceinline__ void load_begin(F desc) {
        void *d0, *d1; const void *s0, *s1; long b0, b1;
        desc(u, d0, s0, b0, d1, s1, b1);
        if (b0 + b1 == 0) return;
        if (use_tma) {
            if (ul == 0) {
                rc_mbar_expect_tx(bar, (uint32_t)(b0 + b1));
                rc_bulk_g2s(d0, s0, (uint32_t)b0, bar);
                if (b1) rc_bulk_g2s(d1, s1, (uint32_t)b1, bar);
            }
        } else {
            rc_coop_copy(d0, s0, b0, ul, unit_lanes);
            if (b1) rc_coop_copy(d1, s1, b1, ul, unit_lanes);
        }
    }
    __device__ __forceinline__ void load_wait() {
        if (use_tma) {
            while (!rc_mbar_try_wait(bar, phase)) {}
            phase ^= 1u;
        } else {
            sync_unit();
        }
    }
    template <class F> __device__ __forceinline__ void store(F desc) {
        void* dst; const void* src; long bytes;
        desc(u, dst, src, bytes);
        if (use_tma) {
            rc_fence_proxy_async();  // make this thread's shared-memory writes visible to the async proxy
            sync_unit();
            if (ul == 0 && bytes) rc_bulk_s2g(dst, src, (uint32_t)bytes);
        } else {
            sync_unit();
            rc_coop_copy(dst, src, bytes, ul, unit_lanes);
        }
    }
    __device__ __forceinline__ void store_drain() {
        if (use_tma) { if (ul == 0) rc_bulk_wait_read0(); }
        else sync_unit();  // cooperative stores of the previous image have finished reading raw out
    }

    // Sum the per-lane partials of one plane (team lanes inside a warp) with a transposed butterfly, then every
    // lane adds "its" elements into the plane's (or warp's) accumulation slot: no atomics, fixed order.
    template <int N>
    __device__ __forceinline__ void wgrad_commit(const ThreadPos& t, const Plan& pl, float (&acc)[N], float* slot) {
        static_assert(N <= 32 || N == 50, "kernel sizes 3, 5, 7");
        if constexpr (N <= 32) {
            float v[32];
            const int gg = g < 32 ? g : 32;
            const int r = tid & (gg - 1);
            switch (gg) {
                case 32: rc_group_reduce<32, N>(v, acc, r); break;
                case 16: rc_group_reduce<16, N>(v, acc, r); break;
                case 8: rc_group_reduce<8, N>(v, acc, r); break;
                case 4: rc_group_reduce<4, N>(v, acc, r); break;
                case 2: rc_group_reduce<2, N>(v, acc, r); break;
                default: rc_group_reduce<1, N>(v, acc, r); break;
            }
            if (t.p < pl.P) {
                const int per = 32 / gg, base = r * per;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < per && base + j < N) slot[base + j] += v[j];
            }
        } else {  // k = 7: 50 values, plain xor butterfly per element
            const int gg = g < 32 ? g : 32;
            for (int off = gg >> 1; off > 0; off >>= 1) {
#pragma unroll
                for (int i = 0; i < N; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
            }
            if (t.p < pl.P) {
                const int r = tid & (gg - 1);
#pragma unroll
                for (int i = 0; i < N; ++i)
                    if ((i & (gg - 1)) == r) slot[i] += acc[i];
            }
        }
    }
};

template <int K, typename T, bool BWD>
__global__ void __launch_bounds__(kMaxThreads, BWD ? 1 : 2) recconv_kernel(const __grid_constant__ Plan pl, const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    DeviceCtx ctx;
    ctx.tid = threadIdx.x; ctx.T = pl.T; ctx.use_tma = pl.use_tma; ctx.g = pl.g; ctx.unit_lanes = pl.unit_lanes;
    ctx.u = threadIdx.x / pl.unit_lanes; ctx.ul = threadIdx.x - ctx.u * pl.unit_lanes;
    ctx.bar = rc_smem_u32(smem + pl.smBar + 8 * ctx.u); ctx.phase = 0;
    if (pl.use_tma && ctx.ul == 0) rc_mbar_init(ctx.bar, 1);
    __syncthreads();
    const int cg = blockIdx.x % pl.n_cg, chunk = blockIdx.x / pl.n_cg;
    ctx.pos = rc_thread_pos(pl, threadIdx.x, cg);
    if (BWD) rc_backward_body<K, T>(ctx, pl, a, smem, cg, chunk);
    else rc_forward_body<K, T>(ctx, pl, a, smem, cg, chunk);
}

// one launcher per (K, dtype, direction); defined in recconv_k{3,5,7}.cu
typedef cudaError_t (*rc_launch_fn)(const Plan&, const KernelArgs&, cudaStream_t);

template <int K, typename T, bool BWD>
cudaError_t rc_launch(const Plan& pl, const KernelArgs& a, cudaStream_t stream) {
    static DeviceOnce configured = {};
    const cudaError_t e0 = rc_once_per_device(configured, [] { return cudaFuncSetAttribute(recconv_kernel<K, T, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    if (e0 != cudaSuccess) return e0;
    recconv_kernel<K, T, BWD><<<pl.n_cg * pl.n_chunk, pl.T, pl.smem_bytes, stream>>>(pl, a);
    return cudaGetLastError();
}

#define RC_INSTANTIATE_K(K)                                                                                    \
    template cudaError_t rc_launch<K, float, false>(const Plan&, const KernelArgs&, cudaStream_t);             \
    template cudaError_t rc_launch<K, float, true>(const Plan&, const KernelArgs&, cudaStream_t);              \
    template cudaError_t rc_launch<K, __nv_bfloat16, false>(const Plan&, const KernelArgs&, cudaStream_t);     \
    template cudaError_t rc_launch<K, __nv_bfloat16, true>(const Plan&, const KernelArgs&, cudaStream_t);      \
    template cudaError_t rc_launch<K, __half, false>(const Plan&, const KernelArgs&, cudaStream_t);            \
    template cudaError_t rc_launch<K, __half, true>(const Plan&, const KernelArgs&, cudaStream_t);

}  // namespace recnext
