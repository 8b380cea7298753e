// lk_mi_aff.cu -- the affine instantiations of the MI update kernels (see the end of lk_mi.cu): a second translation unit so
// that the two halves compile in parallel.
#define MTFB_MI_AFFINE_TU 1
#include "lk_mi.cu"
