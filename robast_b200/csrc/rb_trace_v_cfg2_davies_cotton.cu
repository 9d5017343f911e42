// bounce-kernel instantiations (k_trace + k_step): segmented spherical-facet reflectors (DaviesCotton, HESS1, MST)
#include "rb_trace_kernel.cuh"
RB_DEFINE_TRACE_VARIANT(cfg2_davies_cotton, 1, (RB_SBIT(RBG_SHAPE_BBOX)|RB_SBIT(RBG_SHAPE_TUBE)|RB_SBIT(RBG_SHAPE_SPHERE)|RB_SBIT(RBG_SHAPE_PGON)|RB_SBIT(RBG_SHAPE_INTERSECTION)|RB_SBIT(RBG_SHAPE_SUBTRACTION)), (0u), 4, 512, 2)
