// rb_trace_kernel.cuh — the bounce kernel template; instantiated once per boolean-nesting depth in
// rb_trace_d{0,1,2,3}.cu so the four instantiations compile in parallel.
#ifndef RB_TRACE_KERNEL_CUH
#define RB_TRACE_KERNEL_CUH
#include <cuda_runtime.h>

#include "rb_device.cuh"

struct DRays {
  const double *x, *y, *z, *t, *dx, *dy, *dz, *lambda;
  double *ox, *oy, *oz, *ot, *odx, *ody, *odz;
  int32_t *status, *last_node, *npoints;
  int32_t* cur;     // scratch (multi-launch only)
  uint32_t* ndraw;  // scratch: bit31 = on_boundary, low bits = draw counter
};

#define TRACE_THREADS 128

template <int DEPTH>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace(DScene sc, DTraceParams tp, DRays R, const int32_t* __restrict__ live, long long n, int init,
                                                        int keep_state) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long idx = live ? (long long)live[i] : i;
  RayReg r;
  r.lambda = R.lambda[idx];
  if (init) {
    r.p = v3(R.x[idx], R.y[idx], R.z[idx]);
    r.t = R.t[idx];
    V3 d = v3(R.dx[idx], R.dy[idx], R.dz[idx]);
    double mag = sqrt(dot(d, d));
    r.d = mag > 0 ? (1. / mag) * d : d;  // ARay::SetDirection normalises (src/ARay.cxx:210-223)
    r.status = RBG_RUN;
    r.npoints = 1;
    r.last_node = -1;
    r.ndraw = 0;
    r.on_boundary = 0;
    r.cur = locate_start<DEPTH>(sc, r.p);
  } else {
    r.p = v3(R.ox[idx], R.oy[idx], R.oz[idx]);
    r.t = R.ot[idx];
    r.d = v3(R.odx[idx], R.ody[idx], R.odz[idx]);
    r.status = R.status[idx];
    r.npoints = R.npoints[idx];
    r.last_node = R.last_node[idx];
    uint32_t nd = R.ndraw[idx];
    r.ndraw = nd & 0x7fffffffu;
    r.on_boundary = nd >> 31;
    r.cur = R.cur[idx];
  }
  unsigned long long id = tp.ray_id_offset + (unsigned long long)idx;
  Philox g;
  g.k0 = (uint32_t)tp.seed;
  g.k1 = (uint32_t)(tp.seed >> 32);
  g.id0 = (uint32_t)id;
  g.id1 = (uint32_t)(id >> 32);
  g.ndraw = r.ndraw;
  int steps = 0;
  while (r.status == RBG_RUN && (tp.max_steps <= 0 || steps < tp.max_steps)) {
    trace_step<DEPTH>(sc, tp, r, g);
    steps++;
  }
  R.ox[idx] = r.p.x; R.oy[idx] = r.p.y; R.oz[idx] = r.p.z; R.ot[idx] = r.t;
  R.odx[idx] = r.d.x; R.ody[idx] = r.d.y; R.odz[idx] = r.d.z;
  R.status[idx] = r.status;
  R.last_node[idx] = r.last_node;
  R.npoints[idx] = r.npoints;
  if (keep_state) {
    R.cur[idx] = r.cur;
    R.ndraw[idx] = (g.ndraw & 0x7fffffffu) | ((uint32_t)r.on_boundary << 31);
  }
}


// returns the cudaGetLastError() code of the launch
#define RB_DECLARE_TRACE_LAUNCH(N)                                                                                                 \
  int rb_launch_trace_d##N(const DScene& sc, const DTraceParams& tp, const DRays& R, const int32_t* live, long long n, int init, \
                           int keep, cudaStream_t st)
RB_DECLARE_TRACE_LAUNCH(0);
RB_DECLARE_TRACE_LAUNCH(1);
RB_DECLARE_TRACE_LAUNCH(2);
RB_DECLARE_TRACE_LAUNCH(3);
#define RB_DEFINE_TRACE_LAUNCH(N)                                                                        \
  RB_DECLARE_TRACE_LAUNCH(N) {                                                                           \
    long long blocks = (n + TRACE_THREADS - 1) / TRACE_THREADS;                                          \
    k_trace<N><<<(unsigned)blocks, TRACE_THREADS, 0, st>>>(sc, tp, R, live, n, init, keep);              \
    return (int)cudaGetLastError();                                                                      \
  }
#endif
