// step_kernel_nm4.cu — instantiates the stepping kernels for 4 motors (see step_kernel.cuh)
#include "step_kernel.cuh"

template void launch_step_nm<4>(const DevState&, const DevParams*, double, int, int, bool, cudaStream_t, int*);
