// bounce-kernel instantiations (k_trace + k_step): refractive aspheric systems with union obscurations (SchmidtCassegrain)
#include "rb_trace_kernel.cuh"
RB_DEFINE_TRACE_VARIANT(cfg4_schmidt_cassegrain, 3, (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_ASPHERE)|RB_SBIT(RBG_SHAPE_UNION)), (RB_PH_LENS), 4, 512, 2)
