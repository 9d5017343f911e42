#include "rb_trace_tune.cuh"
