// bounce-kernel instantiation: aspheric mirror systems (SchwarzschildCouder)
#define RB_VARIANT_FUSED 1  // one launch per bounce (rb_variant::fused_bounce)
#include "rb_trace_kernel.cuh"
typedef Combos<> rb_combos_cfg3_aspheric_mirrors;
RB_DEFINE_TRACE_VARIANT(cfg3_aspheric_mirrors, 0, (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_ASPHERE)), (RB_PH_OVERLAP), 256, 4, rb_combos_cfg3_aspheric_mirrors)
