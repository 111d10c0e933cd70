// ffn_tc.h — plan / launch interface of the tcgen05 channel-mixer kernel (ffn_tc.cu)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace recnext {

struct FfnTcPlan {
    int B, C, CP, HID, HIDP, HW, dtype;
    int HWp;                // pixel columns one image takes in tile space: HW, or HW rounded up to 8 when vec = 1 (no 8-pixel chunk straddles two images)
    long P;                 // pixel columns of the whole batch (B x HWp)
    int NT;                 // pixels per tile (128, or 64 for C > 256: TMEM holds D1 x 2 + D2 x ceil(C / 128) accumulators of NT columns)
    int nH, nK1, kwLast, nCT;  // hidden chunks of 128 rows; K tiles (64 wide, the last kwLast wide) of W1; 128-row tiles of W2
    int vec;                // pixels per global access: 8 (HW % 8 == 0), 4 (HW % 4 == 0) or 1
    int nY;                 // activation tile buffers (2 when they fit)
    int RS;                 // weight ring slots (4 .. 8, as many as fit)
    int nD2;                // output accumulators in TMEM (2 when they fit: C <= 128)
    int cs;                 // CTAs per cluster sharing one multicast weight stream (1 or 2)
    int tiles_per_cta;      // the same for every CTA (tiles past the end are idle)
    uint32_t sboY, yBytes, hBytes, offW, offY, offH, offBias, offBar;
    int smem_bytes, ntiles, grid;
    int dbg;                // RECNEXT_FFN_DBG: timing experiments (1 no weight copies, 2 no GELU, 4 no residual / output traffic, 8 no activation loads)
    size_t packed_bytes;    // size of the packed weight stream
};

// 0: plan made; 1: unsupported (dtype not 16-bit, C % 8 != 0, tiles do not fit)
int ffn_tc_make_plan(FfnTcPlan& p, int B, int C, int HID, int HW, int dtype, int num_sms);
// w1 [HID, C], w2 [C, HID] (16-bit, row-major) -> packed weight stream of p.packed_bytes bytes
cudaError_t ffn_tc_pack(const FfnTcPlan& p, const void* w1, const void* w2, void* packed, cudaStream_t st);
void ffn_tc_set_prof(long long* buf);   // timing experiments: device buffer of 4 x 512 clock64 stamps, or null
cudaError_t ffn_tc_launch(const FfnTcPlan& p, const void* y, const void* x, const void* packed, const float* b1, const float* b2, void* out, cudaStream_t st);

}  // namespace recnext
