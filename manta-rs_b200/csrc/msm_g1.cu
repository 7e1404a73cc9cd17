// G1 (Fq) instantiations of the MSM kernels + the field-independent sort / plan kernels.
#define MSM_BUILD_G1 1
#include "msm_impl.inc"
