// bounce-kernel instantiations (k_trace + k_step): any scene, composites nested 6 levels (tutorials/AshraOptics.C reaches 6)
#include "rb_trace_kernel.cuh"
RB_DEFINE_TRACE_VARIANT(generic_d6, 6, (RB_SHAPES_ALL), (RB_PH_ALL), 2, 256, 2)
