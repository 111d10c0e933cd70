// stem.h — plan / launch interface of the fused stem kernel (stem.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace recnext {

struct StemPlan {
    int B, H, W, C1, C2, dtype;
    int H1, W1, H2, W2;      // sizes after conv1 / conv2 (3x3, stride 2, pad 1 each)
    int C1P, C2P;            // C1 padded to 32 or 48 (K of conv2 per tap), C2 padded to a multiple of 16
    int threads, grid;
    int offMid, offW2, offB2, smem_bytes;
    int dbg;                 // RECNEXT_STEM_DBG=1: CTA 0 prints the clocks it spent per phase (timing experiments)
};

// 0: plan made; 1: unsupported (dtype not 16-bit, C1 > 48, C2 > 80)
int stem_make_plan(StemPlan& p, int B, int H, int W, int C1, int C2, int dtype, int num_sms);
// x [B, 3, H, W]; w1p [C1P][32] (k = ci * 9 + ky * 3 + kx, zero padded), b1p [C1P]; w2p [C2P][9 * C1P] (k = (ky * 3 + kx) * C1P + ci), b2p [C2P];
// out [B, C2, H2, W2]; all activations / weights of the plan's 16-bit type, biases fp32
cudaError_t stem_launch(const StemPlan& p, const void* x, const void* w1p, const float* b1p, const void* w2p, const float* b2p, void* out, cudaStream_t st);

}  // namespace recnext
