// rb_variants.h — registry of the compiled bounce-kernel instantiations (most specialised first)
#ifndef RB_VARIANTS_H
#define RB_VARIANTS_H
#include "rb_trace_kernel.cuh"
extern const rb_variant rb_variant_cfg1_parabolic;
extern const rb_variant rb_variant_cfg2_davies_cotton;
extern const rb_variant rb_variant_cfg3_aspheric_mirrors;
extern const rb_variant rb_variant_cfg4_schmidt_cassegrain;
extern const rb_variant rb_variant_cfg5_winston_multilayer;
extern const rb_variant rb_variant_generic_d0;
extern const rb_variant rb_variant_generic_d1;
extern const rb_variant rb_variant_generic_d2;
extern const rb_variant rb_variant_generic_d3;
extern const rb_variant rb_variant_generic_d6;
static const rb_variant* const rb_variants[] = {&rb_variant_cfg1_parabolic, &rb_variant_cfg2_davies_cotton, &rb_variant_cfg3_aspheric_mirrors, &rb_variant_cfg4_schmidt_cassegrain, &rb_variant_cfg5_winston_multilayer, &rb_variant_generic_d0, &rb_variant_generic_d1, &rb_variant_generic_d2, &rb_variant_generic_d3, &rb_variant_generic_d6};
#endif
