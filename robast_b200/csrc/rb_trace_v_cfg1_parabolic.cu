// bounce-kernel instantiation: paraboloid mirrors (SimpleParabolicTelescope)
#define RB_VARIANT_FUSED 2  // one launch per bounce; one launch per trace once the rays are seen to end within two steps (rb_variant::fused_bounce)
#include "rb_trace_kernel.cuh"
typedef Combos<B2<RBG_SHAPE_SUBTRACTION, RBG_SHAPE_PARABOLOID, RBG_SHAPE_PARABOLOID>> rb_combos_cfg1_parabolic;
RB_DEFINE_TRACE_VARIANT(cfg1_parabolic, 1, (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_PARABOLOID)|RB_SBIT(RBG_SHAPE_SUBTRACTION)), (0u), 256, 4, rb_combos_cfg1_parabolic)
