// Experiment (make EXP=1, RB_VARIANT=xsm_512_2): k_step with the geometry tables staged in shared memory (RB_SMEM_SCENE,
// rb_trace_kernel.cuh).
#define RB_X_TAG 400
#define RB_SMEM_SCENE 1
#include "rb_trace_kernel.cuh"
#define CFG2_MASK (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_SPHERE)|RB_SBIT(RBG_SHAPE_PGON)|RB_SBIT(RBG_SHAPE_INTERSECTION)|RB_SBIT(RBG_SHAPE_SUBTRACTION))
RB_DEFINE_X_VARIANT(xsm_512_2, 1, CFG2_MASK, (0u), 4, 512, 2)
RB_DEFINE_X_VARIANT(xsm_384_2, 1, CFG2_MASK, (0u), 4, 384, 2)
RB_DEFINE_X_VARIANT(xsm_256_4, 1, CFG2_MASK, (0u), 4, 256, 4)
