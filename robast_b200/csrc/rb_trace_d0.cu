// bounce kernel instantiation for scenes whose deepest boolean composite nests 0 level(s)
#include "rb_trace_kernel.cuh"
RB_DEFINE_TRACE_LAUNCH(0)
