// wdevice.cuh — CUDA execution context, kernels and launchers of the team-resident RecConv path.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include "recconv_device.cuh"  // mbarrier / cp.async.bulk wrappers, rc_group_reduce
#include "wbody.cuh"

namespace recnext {

struct WDeviceCtx {
    int tl, team, team_lanes, use_tma, LPP;
    uint32_t bar[2], phase[2];
    bool pending[2];

    __device__ __forceinline__ void team_sync() {
        if (team_lanes == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(team_lanes) : "memory");
    }
    long long* prof;
    int prof_i;
    template <class F> __device__ __forceinline__ void stage(F f) {
        f(tl);
        team_sync();
        if (prof && tl == 0 && prof_i < 4000) prof[prof_i++] = clock64();
    }
    // The team has passed a barrier since its last access to dst.  With TMA, lane 0 issues one bulk copy that
    // completes on the team's mbarrier of queue q; otherwise the lanes copy cooperatively (unaligned batches).
    __device__ __forceinline__ void load(void* dst, const void* src, int bytes, int q) {
        if (use_tma) {
            rc_fence_proxy_async();  // order this lane's generic accesses to dst before the async-proxy writes
            team_sync();
            if (tl == 0) {
                rc_mbar_expect_tx(bar[q], (uint32_t)bytes);
                rc_bulk_g2s(dst, src, (uint32_t)bytes, bar[q]);
            }
        } else {
            rc_coop_copy(dst, src, bytes, tl, team_lanes);
        }
        pending[q] = true;
    }
    __device__ __forceinline__ void wait(int q) {
        if (!pending[q]) return;
        if (use_tma) {
            while (!rc_mbar_try_wait(bar[q], phase[q])) {}
            phase[q] ^= 1u;
        } else {
            team_sync();
        }
        pending[q] = false;
    }
    // Adds the per-lane filter-gradient partials of one stage into the plane's accumulation slot.  The LPP lanes
    // of a plane (an aligned group inside a warp, or whole warps with one slot each) are summed with a transposed
    // butterfly; afterwards every lane owns distinct elements of the slot: no atomics, fixed order.
    template <int N>
    __device__ __forceinline__ void reduce(const WPlan& pl, int, float (&acc)[N], float* slot) {
        static_assert(N <= 32 || N == 50, "kernel sizes 3, 5, 7");
        const int gg = LPP < 32 ? LPP : 32;
        const int r = tl & (gg - 1);
        if constexpr (N <= 32) {
            float v[32];
            switch (gg) {
                case 32: rc_group_reduce<32, N>(v, acc, r); break;
                case 16: rc_group_reduce<16, N>(v, acc, r); break;
                case 8: rc_group_reduce<8, N>(v, acc, r); break;
                case 4: rc_group_reduce<4, N>(v, acc, r); break;
                case 2: rc_group_reduce<2, N>(v, acc, r); break;
                default: rc_group_reduce<1, N>(v, acc, r); break;
            }
            const int per = 32 / gg, base = r * per;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < per && base + j < N) slot[base + j] += v[j];
        } else {
            for (int off = gg >> 1; off > 0; off >>= 1) {
#pragma unroll
                for (int i = 0; i < N; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
            }
#pragma unroll
            for (int i = 0; i < N; ++i)
                if ((i & (gg - 1)) == r) slot[i] += acc[i];
        }
        (void)pl;
    }
};

template <int K, typename T, int MAXT>
__device__ __forceinline__ void w_kernel_prologue(WDeviceCtx& ctx, const WPlan& pl, unsigned char* smem) {
    ctx.team_lanes = pl.team_lanes;
    ctx.team = threadIdx.x / pl.team_lanes;
    ctx.tl = threadIdx.x - ctx.team * pl.team_lanes;
    ctx.use_tma = pl.use_tma;
    ctx.LPP = pl.LPP;
    ctx.prof = nullptr; ctx.prof_i = 0;
    for (int q = 0; q < 2; ++q) {
        ctx.bar[q] = rc_smem_u32(smem + pl.smBar + 16 * ctx.team + 8 * q);
        ctx.phase[q] = 0; ctx.pending[q] = false;
        if (pl.use_tma && ctx.tl == 0) rc_mbar_init(ctx.bar[q], 1);
    }
    w_cta_init(pl, smem, threadIdx.x, blockDim.x);
    __syncthreads();
}

// MAXT = 512: up to 16 warps per SM at <= 128 registers; MAXT = 256: up to 8 warps with the full register file
// (K = 7 windows, and the backward's wider live ranges)
template <int K, typename T, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) recconv_wfwd_kernel(const __grid_constant__ WPlan pl, const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    WDeviceCtx ctx;
    w_kernel_prologue<K, T, MAXT>(ctx, pl, smem);
    if (a.prof && blockIdx.x == 0 && ctx.team == 0) { ctx.prof = a.prof; if (ctx.tl == 0) ctx.prof[ctx.prof_i++] = clock64(); }
    w_forward_team<K, T>(ctx, pl, a, smem, ctx.team, blockIdx.x * pl.NT + ctx.team);
}

template <int K, typename T, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) recconv_wbwd_kernel(const __grid_constant__ WPlan pl, const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    WDeviceCtx ctx;
    w_kernel_prologue<K, T, MAXT>(ctx, pl, smem);
    w_cta_init_bwd(pl, smem, threadIdx.x, blockDim.x);
    __syncthreads();
    if (a.prof && blockIdx.x == 0 && ctx.team == 0) { ctx.prof = a.prof; if (ctx.tl == 0) ctx.prof[ctx.prof_i++] = clock64(); }
    w_backward_team<K, T>(ctx, pl, a, smem, ctx.team, blockIdx.x * pl.NT + ctx.team);
}

typedef cudaError_t (*w_launch_fn)(const WPlan&, const KernelArgs&, cudaStream_t);

template <int K, typename T, bool BWD>
cudaError_t w_launch(const WPlan& pl, const KernelArgs& a, cudaStream_t stream);

}  // namespace recnext
