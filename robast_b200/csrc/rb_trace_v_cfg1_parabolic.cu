// bounce-kernel instantiations (k_trace + k_step): SimpleParabolicTelescope-class scenes
#include "rb_trace_kernel.cuh"
RB_DEFINE_TRACE_VARIANT(cfg1_parabolic, 1, (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_PARABOLOID)|RB_SBIT(RBG_SHAPE_SUBTRACTION)), (0u), 4, 512, 2)
