// G2 (Fq2) instantiations of the MSM kernels.
#define MSM_BUILD_G2 1
#include "msm_impl.inc"
