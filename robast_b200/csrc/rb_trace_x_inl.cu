// experiment: primitives inlined into the boolean functions / kernel body (built only with `make EXP=1`)
#define RB_PRIM_CALL inline
#include "rb_trace_kernel.cuh"
#define CFG2_MASK (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_SPHERE)|RB_SBIT(RBG_SHAPE_PGON)|RB_SBIT(RBG_SHAPE_INTERSECTION)|RB_SBIT(RBG_SHAPE_SUBTRACTION))
RB_DEFINE_TRACE_VARIANT(x2i_512_2, 1, CFG2_MASK, (0u), 4, 512, 2)
RB_DEFINE_TRACE_VARIANT(x2i_384_2, 1, CFG2_MASK, (0u), 4, 384, 2)
RB_DEFINE_TRACE_VARIANT(x2i_256_2, 1, CFG2_MASK, (0u), 4, 256, 2)
extern const rb_variant* const rb_xi_variants[] = {&rb_variant_x2i_512_2, &rb_variant_x2i_384_2, &rb_variant_x2i_256_2, nullptr};
