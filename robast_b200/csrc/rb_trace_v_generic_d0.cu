// bounce-kernel instantiations (k_trace + k_step): any scene without boolean composites
#include "rb_trace_kernel.cuh"
RB_DEFINE_TRACE_VARIANT(generic_d0, 0, (RB_SHAPES_ALL), (RB_PH_ALL), 2, 256, 2)
