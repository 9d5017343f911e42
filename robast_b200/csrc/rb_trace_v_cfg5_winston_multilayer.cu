// bounce-kernel instantiation: Winston-cone light guides with multilayer coatings (HexWinstonCone)
#include "rb_trace_kernel.cuh"
typedef Combos<B2<RBG_SHAPE_SUBTRACTION, RBG_SHAPE_PGON, RBG_SHAPE_WINSTONPOLY>> rb_combos_cfg5_winston_multilayer;
RB_DEFINE_TRACE_VARIANT(cfg5_winston_multilayer, 1, (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_PGON)|RB_SBIT(RBG_SHAPE_WINSTONPOLY)|RB_SBIT(RBG_SHAPE_SUBTRACTION)), (RB_PH_MULTILAYER), 256, 4, rb_combos_cfg5_winston_multilayer)
