// team-resident RecConv kernels (wplan.h), K = 5: the reference's kernel size (model/recnext.py:152)
#include "wdevice.cuh"
#include "wlaunch.cuh"
namespace recnext {
W_INSTANTIATE_K(5)
}
