// Experiment instantiations of the bounce kernel (built only with `make EXP=1`, selected at run time with
// RB_VARIANT=<name>, swept by profiles/sweep_variants.py).  Each experiment unit sets its own RB_X_TAG: TraceCfg carries it
// as a template parameter so that instantiations whose other parameters coincide stay distinct kernels.
// This unit: k_step without its phase barriers (RB_STEP_BARRIERS mask, rb_trace_kernel.cuh) — measured 1.11e9 rays/s
// against 1.37e9 with the barriers on DaviesCotton (profiles/r1m_ncu_summary.md).
#define RB_X_TAG 100
#define RB_STEP_BARRIERS 0
#include "rb_trace_kernel.cuh"
#define CFG2_MASK (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_SPHERE)|RB_SBIT(RBG_SHAPE_PGON)|RB_SBIT(RBG_SHAPE_INTERSECTION)|RB_SBIT(RBG_SHAPE_SUBTRACTION))
RB_DEFINE_X_VARIANT(xb0_512_2, 1, CFG2_MASK, (0u), 4, 512, 2)
RB_DEFINE_X_VARIANT(xb0_256_4, 1, CFG2_MASK, (0u), 4, 256, 4)
RB_DEFINE_X_VARIANT(xb0_128_8, 1, CFG2_MASK, (0u), 4, 128, 8)
