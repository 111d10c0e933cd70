// Fused RecConv kernels for kernel_size = 5 (all element types, forward and backward).
#include "recconv_device.cuh"
namespace recnext { RC_INSTANTIATE_K(5) }
