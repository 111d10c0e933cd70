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

// RecAttn2d pieces without the tensor-core kernels' dtype / size limits (fp32 activations, k = 3 / 5 / 7, any plane size); no workspace.
// variant 1: out = conv_s2(x) + b; variant 2: out = conv_s1(x + interpolate(z)) + b   (model/recattn.py:60, :67)
cudaError_t gstream_recattn(const GStreamDesc& d, int variant, const void* w, const void* b, const void* x, const void* z, int zH, int zW, void* out,
                            cudaStream_t stream);

}  // namespace recnext
