// bounce-kernel instantiation: every shape and every physics branch, boolean nesting up to 0
#include "rb_trace_kernel.cuh"
typedef Combos<> rb_combos_generic_d0;
RB_DEFINE_TRACE_VARIANT(generic_d0, 0, (RB_SHAPES_ALL), (RB_PH_ALL), 128, 4, rb_combos_generic_d0)
