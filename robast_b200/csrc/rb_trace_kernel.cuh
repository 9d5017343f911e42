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
  DHist hist;       // optional polyline record (hist.x == nullptr: off); per-ray loop kernel only
};

#define TRACE_THREADS 128
// k_step phase barriers: 1 = before the candidate shapes, 2 = barrier + vote around the candidate loop (re-synchronises the
// batches of rays with more than RB_MAXVIS candidates), 4 = before the interaction
#ifndef RB_STEP_BARRIERS
#define RB_STEP_BARRIERS 7
#endif

// Ray state streams through once per launch: mark it evict-first (ld/st .cs) so that it does not push the kernels'
// local-memory working set (spill slots and the call stack, ~100 MB over all resident threads) out of the 126 MB L2.
#if defined(__CUDA_ARCH__) && !defined(RB_NO_STREAM_HINTS)
template <class T> __device__ inline T rb_ldcs(const T* p) { return __ldcs(p); }
template <class T> __device__ inline void rb_stcs(T* p, T v) { __stcs(p, v); }
#else
template <class T> __device__ inline T rb_ldcs(const T* p) { return *p; }
template <class T> __device__ inline void rb_stcs(T* p, T v) { *p = v; }
#endif

template <class K> __device__ inline void load_ray(const DScene& sc, const DTraceParams& tp, const DRays& R, long long idx, int init, RayReg& r, Philox& g) {
  r.lambda = rb_ldcs(R.lambda + idx);
  if (init) {
    r.p = v3(rb_ldcs(R.x + idx), rb_ldcs(R.y + idx), rb_ldcs(R.z + idx));
    r.t = rb_ldcs(R.t + idx);
    V3 d = v3(rb_ldcs(R.dx + idx), rb_ldcs(R.dy + idx), rb_ldcs(R.dz + idx));
    double mag = sqrt(dot(d, d));
    r.d = mag > 0 ? (1. / mag) * d : d;  // ARay::SetDirection normalises (src/ARay.cxx:210-223)
    r.status = RBG_RUN;
    r.npoints = 1;
    r.last_node = -1;
    r.ndraw = 0;
    r.on_boundary = 0;
    r.cur = locate_start<K>(sc, r.p);
  } else {
    r.p = v3(rb_ldcs(R.ox + idx), rb_ldcs(R.oy + idx), rb_ldcs(R.oz + idx));
    r.t = rb_ldcs(R.ot + idx);
    r.d = v3(rb_ldcs(R.odx + idx), rb_ldcs(R.ody + idx), rb_ldcs(R.odz + idx));
    r.status = R.status[idx];
    r.npoints = rb_ldcs(R.npoints + idx);
    r.last_node = rb_ldcs(R.last_node + idx);
    uint32_t nd = rb_ldcs(R.ndraw + idx);
    r.ndraw = nd & 0x7fffffffu;
    r.on_boundary = nd >> 31;
    r.cur = rb_ldcs(R.cur + idx);
  }
  unsigned long long id = tp.ray_id_offset + (unsigned long long)idx;
  g.k0 = (uint32_t)tp.seed;
  g.k1 = (uint32_t)(tp.seed >> 32);
  g.id0 = (uint32_t)id;
  g.id1 = (uint32_t)(id >> 32);
  g.ndraw = r.ndraw;
}
__device__ inline void store_ray(const DRays& R, long long idx, const RayReg& r, const Philox& g, int keep_state) {
  rb_stcs(R.ox + idx, r.p.x); rb_stcs(R.oy + idx, r.p.y); rb_stcs(R.oz + idx, r.p.z); rb_stcs(R.ot + idx, r.t);
  rb_stcs(R.odx + idx, r.d.x); rb_stcs(R.ody + idx, r.d.y); rb_stcs(R.odz + idx, r.d.z);
  R.status[idx] = r.status;  // read again by k_compact right after this launch: leave it cached
  rb_stcs(R.last_node + idx, (int32_t)r.last_node);
  rb_stcs(R.npoints + idx, (int32_t)r.npoints);
  if (keep_state) {
    rb_stcs(R.cur + idx, (int32_t)r.cur);
    rb_stcs(R.ndraw + idx, (uint32_t)((g.ndraw & 0x7fffffffu) | ((uint32_t)r.on_boundary << 31)));
  }
}

// ---- per-ray loop kernel: `max_steps` boundary steps per launch (<= 0: until the ray terminates)
template <class K>
__global__ void __launch_bounds__(TRACE_THREADS, K::min_blocks) k_trace(const __grid_constant__ DScene sc, const __grid_constant__ DTraceParams tp,
                                                                        const __grid_constant__ DRays R, const int32_t* __restrict__ live, long long n,
                                                                        int init, int keep_state) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long idx = live ? (long long)live[i] : i;
  RayReg r;
  Philox g;
  load_ray<K>(sc, tp, R, idx, init, r, g);
  HistSink sink;
  sink.h = &R.hist;
  sink.idx = idx;
  const HistSink* hs = R.hist.x ? &sink : nullptr;
  if (hs && init && R.hist.max_points > 0) {  // point 0 = the start point (no node entry)
    R.hist.x[idx] = r.p.x; R.hist.y[idx] = r.p.y; R.hist.z[idx] = r.p.z; R.hist.t[idx] = r.t;
    R.hist.node[idx] = -1;
  }
  int steps = 0;
  while (r.status == RBG_RUN && (tp.max_steps <= 0 || steps < tp.max_steps)) {
    trace_step<K>(sc, tp, r, g, hs);
    steps++;
  }
  store_ray(R, idx, r, g, keep_state);
}

// ---- wavefront bounce kernel: exactly one boundary step per live ray, executed in block-wide lock step.
// The phases of the step (DistFromInside / BVH walk / candidate DistFromOutside / relocation / interaction) are
// separated by barriers so that all warps of a block run the same code region at the same time: the per-SM
// instruction cache is then shared instead of thrashed (profiles/r1d: icc hit rate 58 %, L1.5 saturated by
// instruction requests without this).  Idle lanes of the last block shadow the last ray and store nothing.
#ifndef RB_SMEM_SCENE
#define RB_SMEM_SCENE 0
#endif
__host__ __device__ inline size_t rb_smem_round(size_t b) { return (b + 15) & ~size_t(15); }
// block-cooperative copy of `bytes` (multiple of 16; the source tables are cudaMalloc'ed, i.e. 256-byte aligned and padded by the
// allocator's granularity) into shared memory
__device__ inline void rb_smem_copy(char* dst, const void* src, size_t bytes) {
  const int4* s4 = (const int4*)src;
  int4* d4 = (int4*)dst;
  for (size_t k = threadIdx.x; k < bytes / 16; k += blockDim.x) d4[k] = s4[k];
}
// Split ray load for k_step: the navigation phases need only the point, the direction, the node and the on-boundary bit;
// time, wavelength, counters and the segment start are read when the interaction is evaluated, so that they are not
// live (= spilled around every non-inlined shape call) during navigation.
template <class K> __device__ inline void load_nav(const DScene& sc, const DRays& R, long long idx, int from_out, RayReg& r) {
  if (!from_out) {
    r.p = v3(rb_ldcs(R.x + idx), rb_ldcs(R.y + idx), rb_ldcs(R.z + idx));
    V3 d = v3(rb_ldcs(R.dx + idx), rb_ldcs(R.dy + idx), rb_ldcs(R.dz + idx));
    double mag = sqrt(dot(d, d));
    r.d = mag > 0 ? (1. / mag) * d : d;  // ARay::SetDirection normalises (src/ARay.cxx:210-223)
    r.status = RBG_RUN;
    r.on_boundary = 0;
    r.cur = locate_start<K>(sc, r.p);
  } else {
    r.p = v3(rb_ldcs(R.ox + idx), rb_ldcs(R.oy + idx), rb_ldcs(R.oz + idx));
    r.d = v3(rb_ldcs(R.odx + idx), rb_ldcs(R.ody + idx), rb_ldcs(R.odz + idx));
    r.status = R.status[idx];
    r.on_boundary = R.ndraw[idx] >> 31;
    r.cur = rb_ldcs(R.cur + idx);
  }
}
__device__ inline void load_rest(const DTraceParams& tp, const DRays& R, long long idx, int from_out, RayReg& r, Philox& g) {
  r.lambda = rb_ldcs(R.lambda + idx);
  if (!from_out) {
    r.p = v3(rb_ldcs(R.x + idx), rb_ldcs(R.y + idx), rb_ldcs(R.z + idx));
    r.t = rb_ldcs(R.t + idx);
    r.npoints = 1;
    r.last_node = -1;
    r.ndraw = 0;
  } else {
    r.p = v3(rb_ldcs(R.ox + idx), rb_ldcs(R.oy + idx), rb_ldcs(R.oz + idx));
    r.t = rb_ldcs(R.ot + idx);
    r.npoints = rb_ldcs(R.npoints + idx);
    r.last_node = rb_ldcs(R.last_node + idx);
    r.ndraw = rb_ldcs(R.ndraw + idx) & 0x7fffffffu;
  }
  unsigned long long id = tp.ray_id_offset + (unsigned long long)idx;
  g.k0 = (uint32_t)tp.seed;
  g.k1 = (uint32_t)(tp.seed >> 32);
  g.id0 = (uint32_t)id;
  g.id1 = (uint32_t)(id >> 32);
  g.ndraw = r.ndraw;
}

template <class K>
__global__ void __launch_bounds__(K::step_threads, K::step_min_blocks) k_step(const __grid_constant__ DScene sc_, const __grid_constant__ DTraceParams tp,
                                                                              const __grid_constant__ DRays R, const int32_t* __restrict__ live,
                                                                              long long n, int init, int smem_bytes) {
#if RB_SMEM_SCENE
  // Geometry tables (nodes, BVH, shapes, parameters, operand matrices: tens of kB) staged in shared memory when they fit the
  // block's dynamic allocation: their loads then no longer compete for L1 with the kernel's local-memory traffic.
  extern __shared__ double4 rb_smem[];
  __shared__ DScene ssc;
  {
    char* base = (char*)rb_smem;
    const size_t b0 = rb_smem_round((size_t)sc_.nnodes * sizeof(DNode)), b1 = rb_smem_round((size_t)sc_.nbvh * sizeof(DBvh)),
                 b2 = rb_smem_round((size_t)sc_.nshapes * sizeof(DShape)), b3 = rb_smem_round((size_t)sc_.ndpar * 8),
                 b4 = rb_smem_round((size_t)sc_.nmats * sizeof(DMat));
    const bool fits = b0 + b1 + b2 + b3 + b4 <= (size_t)smem_bytes;
    if (threadIdx.x == 0) {
      ssc = sc_;
      if (fits) {
        ssc.nodes = (const DNode*)base; ssc.bvh = (const DBvh*)(base + b0); ssc.shapes = (const DShape*)(base + b0 + b1);
        ssc.dpar = (const double*)(base + b0 + b1 + b2); ssc.mats = (const DMat*)(base + b0 + b1 + b2 + b3);
      }
    }
    if (fits) {
      rb_smem_copy(base, sc_.nodes, b0); rb_smem_copy(base + b0, sc_.bvh, b1); rb_smem_copy(base + b0 + b1, sc_.shapes, b2);
      rb_smem_copy(base + b0 + b1 + b2, sc_.dpar, b3); rb_smem_copy(base + b0 + b1 + b2 + b3, sc_.mats, b4);
    }
    __syncthreads();
  }
  const DScene& sc = ssc;
#else
  const DScene& sc = sc_;
#endif
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  if (!active) i = n - 1;
  long long idx = live ? (long long)live[i] : i;
  int from_out = !init;
  RayReg nav;
  load_nav<K>(sc, R, idx, from_out, nav);
  // a ray shot from outside the top volume first has to enter it (no daughters to examine, no interaction
  // besides AddPoint): take that step here instead of spending a whole bounce launch on it
  if (active && init && nav.cur < 0) {
    RayReg r = nav;
    Philox g;
    load_rest(tp, R, idx, 0, r, g);
    r.p = nav.p;
    trace_step<K>(sc, tp, r, g);
    store_ray(R, idx, r, g, 1);
    nav.p = r.p; nav.d = r.d; nav.cur = r.cur; nav.status = r.status; nav.on_boundary = r.on_boundary;
    from_out = 1;
  }
  const bool run = active && nav.status == RBG_RUN;
  const bool push = (tp.quirks & RBG_QUIRK_BOUNDARY_PUSH) != 0;
  const int cur0 = nav.cur;
  NavStep st;
  st.mode = 0;
  st.bvh_next = -1;
  st.o.nvis = -1;
  if (run) nb_begin<K>(sc, nav, push, st);
#if (RB_STEP_BARRIERS & 16)
  __syncthreads();
#endif
  bool overflow = false;
  bool want = run && st.mode == 1 && st.bvh_next >= 0;
#if (RB_STEP_BARRIERS & 2)
  do {  // one pass unless some ray touches more than RB_MAXVIS daughter boxes
    int first = 0;
    if (want) {
      if (st.o.nvis >= RB_MAXVIS) { st.o.nvis = 0; overflow = true; }
      first = st.o.nvis;
      nb_collect<K>(sc, nav, st);
    }
    __syncthreads();  // all warps enter the shape code together; inside the phase they run unsynchronised
#if (RB_STEP_BARRIERS & 8)
    for (int k = first; __syncthreads_or(want && k < st.o.nvis); k++)
      if (want && k < st.o.nvis) nb_eval<K>(sc, nav, st, k);
#else
    if (want)
      for (int k = first; k < st.o.nvis; k++) nb_eval<K>(sc, nav, st, k);
#endif
    want = run && st.mode == 1 && st.bvh_next >= 0;
  } while (__syncthreads_or(want));
#else
  if (want) nb_collect<K>(sc, nav, st);
#if (RB_STEP_BARRIERS & 1)
  __syncthreads();  // all warps enter the shape code together; inside the phase they run unsynchronised
#endif
  if (want) {
    int first = 0;
    while (true) {
      for (int k = first; k < st.o.nvis; k++) nb_eval<K>(sc, nav, st, k);
      if (st.bvh_next < 0) break;  // rare: more than RB_MAXVIS daughter boxes touched, continue the walk in batches
      st.o.nvis = 0;
      overflow = true;
      nb_collect<K>(sc, nav, st);
    }
  }
#endif
  if (overflow) st.o.nvis = -1;
  if (run) nb_finish<K>(sc, nav, st);
#if (RB_STEP_BARRIERS & 4)
  __syncthreads();
#endif
  if (run) {
    RayReg r;
    Philox g;
    load_rest(tp, R, idx, from_out, r, g);
    r.d = nav.d;
    r.cur = cur0;
    r.status = RBG_RUN;
    r.on_boundary = 0;
    trace_shade<K>(sc, tp, r, nav, st.o, g);
    store_ray(R, idx, r, g, 1);
  } else if (active && !from_out) {  // a first launch always leaves a complete record (not reached: every ray starts running)
    RayReg r = nav;
    Philox g;
    load_rest(tp, R, idx, 0, r, g);
    store_ray(R, idx, r, g, 1);
  }
}

// Launch entry points per compiled variant (each variant in its own translation unit so they build in parallel).
typedef int (*rb_launch_fn)(const DScene& sc, const DTraceParams& tp, const DRays& R, const int32_t* live, long long n, int init, int keep,
                            cudaStream_t st);
struct rb_variant {
  const char* name;
  int depth;
  unsigned shapes, phys;
  rb_launch_fn launch;       // k_trace: per-ray loop
  rb_launch_fn launch_step;  // k_step : one lock-step boundary step (wavefront)
};
#ifndef RB_X_TAG
#define RB_X_TAG 0
#endif
#define RB_DEFINE_TRACE_VARIANT(NAME, D, S, P, MB, STH, SMB)                                                                          \
  typedef TraceCfg<D, S, P, MB, STH, SMB, RB_X_TAG> rb_cfg_##NAME;                                                                              \
  int rb_launch_trace_##NAME(const DScene& sc, const DTraceParams& tp, const DRays& R, const int32_t* live, long long n, int init,   \
                             int keep, cudaStream_t st) {                                                                            \
    long long blocks = (n + TRACE_THREADS - 1) / TRACE_THREADS;                                                                      \
    k_trace<rb_cfg_##NAME><<<(unsigned)blocks, TRACE_THREADS, 0, st>>>(sc, tp, R, live, n, init, keep);                              \
    return (int)cudaGetLastError();                                                                                                  \
  }                                                                                                                                  \
  int rb_launch_step_##NAME(const DScene& sc, const DTraceParams& tp, const DRays& R, const int32_t* live, long long n, int init,    \
                            int, cudaStream_t st) {                                                                                  \
    long long blocks = (n + STH - 1) / STH;                                                                                          \
    int smem = 0;                                                                                                                    \
    if (RB_SMEM_SCENE) {                                                                                                             \
      size_t need = rb_smem_round((size_t)sc.nnodes * sizeof(DNode)) + rb_smem_round((size_t)sc.nbvh * sizeof(DBvh)) +               \
                    rb_smem_round((size_t)sc.nshapes * sizeof(DShape)) + rb_smem_round((size_t)sc.ndpar * 8) +                      \
                    rb_smem_round((size_t)sc.nmats * sizeof(DMat));                                                                 \
      if (need <= 46 * 1024) smem = (int)need;                                                                                       \
    }                                                                                                                                \
    k_step<rb_cfg_##NAME><<<(unsigned)blocks, STH, smem, st>>>(sc, tp, R, live, n, init, smem);                                      \
    return (int)cudaGetLastError();                                                                                                  \
  }                                                                                                                                  \
  extern const rb_variant rb_variant_##NAME = {#NAME, D, S, P, rb_launch_trace_##NAME, rb_launch_step_##NAME};
// experiment instantiations (make EXP=1) add themselves to the RB_VARIANT=<name> lookup
int rb_register_x_variant(const rb_variant* v);
#define RB_DEFINE_X_VARIANT(NAME, D, S, P, MB, STH, SMB) \
  RB_DEFINE_TRACE_VARIANT(NAME, D, S, P, MB, STH, SMB)   \
  static int rb_reg_##NAME = rb_register_x_variant(&rb_variant_##NAME);
#endif
