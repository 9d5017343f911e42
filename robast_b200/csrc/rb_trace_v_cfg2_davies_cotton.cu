// bounce-kernel instantiation: segmented spherical-facet reflectors (DaviesCotton, HESS1, MST)
#include "rb_trace_kernel.cuh"
typedef Combos<B2<RBG_SHAPE_INTERSECTION, RBG_SHAPE_SPHERE, RBG_SHAPE_PGON>, B2<RBG_SHAPE_SUBTRACTION, RBG_SHAPE_BBOX, RBG_SHAPE_BBOX>, B2<RBG_SHAPE_SUBTRACTION, RBG_SHAPE_TUBE, RBG_SHAPE_TUBE>> rb_combos_cfg2_davies_cotton;
RB_DEFINE_TRACE_VARIANT(cfg2_davies_cotton, 1, (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_SPHERE)|RB_SBIT(RBG_SHAPE_PGON)|RB_SBIT(RBG_SHAPE_INTERSECTION)|RB_SBIT(RBG_SHAPE_SUBTRACTION)), (0u), 256, 4, rb_combos_cfg2_davies_cotton)
