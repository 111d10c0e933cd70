// gstream.h — streamed (level-by-level, global-memory) RecConv path for planes that do not fit on chip; see gstream.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace recnext {

struct KernelArgs;

struct GStreamDesc {
    int B, C, H, W, K, L, mode, dtype, wdtype, has_bias, num_sms;
};

size_t gstream_workspace_bytes(const GStreamDesc& d, bool bwd);
// forward: a.x -> a.out; backward: a.x, a.gy -> a.out (gx), gw [(L+2), C, K*K], gb [(L+2), C] or null
cudaError_t gstream_launch(const GStreamDesc& d, const KernelArgs& a, void* workspace, bool bwd, float* gw, float* gb, cudaStream_t stream);

}  // namespace recnext
