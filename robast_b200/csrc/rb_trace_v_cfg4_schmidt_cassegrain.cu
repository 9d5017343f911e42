// bounce-kernel instantiation: catadioptric systems with aspheric lenses (SchmidtCassegrain)
#include "rb_trace_kernel.cuh"
typedef Combos<> rb_combos_cfg4_schmidt_cassegrain;
RB_DEFINE_TRACE_VARIANT(cfg4_schmidt_cassegrain, 3, (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_ASPHERE)|RB_SBIT(RBG_SHAPE_UNION)), (RB_PH_LENS), 256, 4, rb_combos_cfg4_schmidt_cassegrain)
