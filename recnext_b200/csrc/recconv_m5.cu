// tensor-core RecConv forward (mplan.h / mfwd.cuh): 16-bit activations, K = 5 — the reference's kernel size
// (model/recnext.py:152)
#include "mfwd.cuh"
namespace recnext {
cudaError_t m_launch(const MPlan& pl, const KernelArgs& a, cudaStream_t stream) { return m_launch_fwd(pl, a, stream); }
bool m_static_geometry(const MPlan& pl) { return m_is_static(pl); }
}
