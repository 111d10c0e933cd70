// tensor-core RecConv backward (mbplan.h / mbwd.cuh): 16-bit activations, K = 5 — the reference's kernel size
// (model/recnext.py:152)
#include "mbwd.cuh"
namespace recnext {
cudaError_t mb_launch(const MBPlan& bp, const KernelArgs& a, float* gw, float* gb, cudaStream_t stream) { return mb_launch_bwd(bp, a, gw, gb, stream); }
}
