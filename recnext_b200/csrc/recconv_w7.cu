// team-resident RecConv kernels (wplan.h), K = 7
#include "wdevice.cuh"
#include "wlaunch.cuh"
namespace recnext {
W_INSTANTIATE_K(7)
}
