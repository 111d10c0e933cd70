// wlaunch.cuh — launchers of the team-resident kernels (one per K, dtype, direction)
#pragma once
#include "wdevice.cuh"

namespace recnext {

template <int K, typename T, bool BWD>
cudaError_t w_launch(const WPlan& pl, const KernelArgs& a, cudaStream_t stream) {
    if constexpr (BWD) {
        static int configured_b = 0;
        if (!configured_b) {
            cudaError_t e = cudaFuncSetAttribute(recconv_wbwd_kernel<K, T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return e;
            if (K < 7) {
                e = cudaFuncSetAttribute(recconv_wbwd_kernel<K, T, (K < 7 ? 384 : 256)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (e != cudaSuccess) return e;
                e = cudaFuncSetAttribute(recconv_wbwd_kernel<K, T, (K < 7 ? 512 : 256)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (e != cudaSuccess) return e;
            }
            configured_b = 1;
        }
        if (pl.threads <= 256) recconv_wbwd_kernel<K, T, 256><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
        else if (K < 7 && pl.threads <= 384) recconv_wbwd_kernel<K, T, (K < 7 ? 384 : 256)><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
        else if (K < 7) recconv_wbwd_kernel<K, T, (K < 7 ? 512 : 256)><<<pl.grid, pl.threads, pl.smem_bytes, stream>>>(pl, a);
        else return cudaErrorInvalidConfiguration;
        return cudaGetLastError();
    }
    static int configured = 0;  // benign race: idempotent
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(recconv_wfwd_kernel<K, T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        if (K < 7) {
            e = cudaFuncSetAttribute(recconv_wfwd_kernel<K, T, (K < 7 ? 512 : 256)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return e;
        }
        configured = 1;
    }
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
