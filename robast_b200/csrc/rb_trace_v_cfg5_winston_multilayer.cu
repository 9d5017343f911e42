// bounce-kernel instantiations (k_trace + k_step): Winston-cone light guides with multilayer-coated walls (HexWinstonCone)
#include "rb_trace_kernel.cuh"
RB_DEFINE_TRACE_VARIANT(cfg5_winston_multilayer, 1, (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_PGON)|RB_SBIT(RBG_SHAPE_WINSTONPOLY)|RB_SBIT(RBG_SHAPE_SUBTRACTION)), (RB_PH_MULTILAYER), 4, 512, 2)
