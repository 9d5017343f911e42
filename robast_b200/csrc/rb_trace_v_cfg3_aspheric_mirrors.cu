// bounce-kernel instantiations (k_trace + k_step): aspheric-disk mirror systems (SchwarzschildCouder)
#include "rb_trace_kernel.cuh"
RB_DEFINE_TRACE_VARIANT(cfg3_aspheric_mirrors, 0, (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_ASPHERE)), (RB_PH_OVERLAP), 4, 512, 2)
