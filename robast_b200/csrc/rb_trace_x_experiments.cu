// launch-bound experiments for the bounce kernels (built only with `make EXP=1`; selected with RB_VARIANT=<name>)
#include "rb_trace_kernel.cuh"
#define CFG2_MASK (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_SPHERE)|RB_SBIT(RBG_SHAPE_PGON)|RB_SBIT(RBG_SHAPE_INTERSECTION)|RB_SBIT(RBG_SHAPE_SUBTRACTION))
RB_DEFINE_TRACE_VARIANT(x2_256_4, 1, CFG2_MASK, (0u), 4, 256, 4)
RB_DEFINE_TRACE_VARIANT(x2_256_3, 1, CFG2_MASK, (0u), 4, 256, 3)
RB_DEFINE_TRACE_VARIANT(x2_256_2, 1, CFG2_MASK, (0u), 4, 256, 2)
RB_DEFINE_TRACE_VARIANT(x2_384_2, 1, CFG2_MASK, (0u), 4, 384, 2)
RB_DEFINE_TRACE_VARIANT(x2_128_8, 1, CFG2_MASK, (0u), 4, 128, 8)
RB_DEFINE_TRACE_VARIANT(x2_128_6, 1, CFG2_MASK, (0u), 4, 128, 6)
RB_DEFINE_TRACE_VARIANT(x2_1024_1, 1, CFG2_MASK, (0u), 4, 1024, 1)
RB_DEFINE_TRACE_VARIANT(x2_512_1, 1, CFG2_MASK, (0u), 4, 512, 1)
extern const rb_variant* const rb_x_variants[] = {&rb_variant_x2_256_4, &rb_variant_x2_256_3, &rb_variant_x2_256_2, &rb_variant_x2_384_2, &rb_variant_x2_128_8, &rb_variant_x2_128_6, &rb_variant_x2_1024_1, &rb_variant_x2_512_1, nullptr};
