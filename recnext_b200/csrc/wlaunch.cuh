// wlaunch.cuh — launchers of the team-resident kernels (one per K, dtype, direction)
#pragma once
#include "wdevice.cuh"
#include "devcfg.h"

namespace recnext {

template <int K, typename T, bool BWD>
cudaError_t w_launch(const WPlan& pl, const KernelArgs& a, cudaStream_t stream) {
    if constexpr (BWD) {
        static DeviceOnce configured_b = {};
        const cudaError_t e0 = rc_once_per_device(configured_b, [] {
            cudaError_t e = cudaFuncSetAttribute(recconv_wbwd_kernel<K, T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess && K < 7) e = cudaFuncSetAttribute(recconv_wbwd_kernel<K, T, (K < 7 ? 384 : 256)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess && K < 7) e = cudaFuncSetAttribute(recconv_wbwd_kernel<K, T, (K < 7 ? 512 : 256)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            return e;
        });
        if (e0 != cudaSuccess) return e0;
        if (pl.threads <= 256) recconv_wbwd_kernel<K, T, 256><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
        else if (K < 7 && pl.threads <= 384) recconv_wbwd_kernel<K, T, (K < 7 ? 384 : 256)><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
        else if (K < 7) recconv_wbwd_kernel<K, T, (K < 7 ? 512 : 256)><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
        else return cudaErrorInvalidConfiguration;
        return cudaGetLastError();
    }
    static DeviceOnce configured = {};
    const cudaError_t e1 = rc_once_per_device(configured, [] {
        cudaError_t e = cudaFuncSetAttribute(recconv_wfwd_kernel<K, T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess && K < 7) e = cudaFuncSetAttribute(recconv_wfwd_kernel<K, T, (K < 7 ? 512 : 256)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        return e;
    });
    if (e1 != cudaSuccess) return e1;
    if (pl.threads <= 256) recconv_wfwd_kernel<K, T, 256><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
    else if (K < 7) recconv_wfwd_kernel<K, T, (K < 7 ? 512 : 256)><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
    else return cudaErrorInvalidConfiguration;
    return cudaGetLastError();
}

#define W_INSTANTIATE_K(K)                                                                                \
    template cudaError_t w_launch<K, float, false>(const WPlan&, const KernelArgs&, cudaStream_t);          \
    template cudaError_t w_launch<K, __nv_bfloat16, false>(const WPlan&, const KernelArgs&, cudaStream_t);  \
    template cudaError_t w_launch<K, __half, false>(const WPlan&, const KernelArgs&, cudaStream_t);         \
    template cudaError_t w_launch<K, float, true>(const WPlan&, const KernelArgs&, cudaStream_t);           \
    template cudaError_t w_launch<K, __nv_bfloat16, true>(const WPlan&, const KernelArgs&, cudaStream_t);   \
    template cudaError_t w_launch<K, __half, true>(const WPlan&, const KernelArgs&, cudaStream_t);

}  // namespace recnext
