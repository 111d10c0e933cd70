// devcfg.h — per-DEVICE one-time configuration.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are
// properties of a (function, device) pair, not of the process: a process that drives several GPUs (per-GPU serving
// threads, a model moved to cuda:1) must opt in on each of them.
#pragma once
#include <cuda_runtime.h>

namespace recnext {

constexpr int kMaxDevices = 64;

struct DeviceOnce {
    unsigned char done[kMaxDevices];
};

// runs `f` (returns cudaError_t) the first time it is reached with the current device; benign race: f is idempotent
template <class F>
inline cudaError_t rc_once_per_device(DeviceOnce& flags, F f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return f();
    if (flags.done[dev]) return cudaSuccess;
    e = f();
    if (e == cudaSuccess) flags.done[dev] = 1;
    return e;
}

// SM count of the CURRENT device (cached per device: one process may drive several GPUs); 148 when no device is visible
inline int rc_device_sms() {
    static int sms[kMaxDevices] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms[dev] = v;
        else return 148;
    }
    return sms[dev];
}

}  // namespace recnext
