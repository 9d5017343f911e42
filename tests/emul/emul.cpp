// emul.cpp — DEBUGGING AID FOR THE GPU-LESS BUILD BOX, NOT PART OF THE PRODUCT.
// Compiles the device functions of robast_b200/csrc/rb_device.cuh for the host (RB_HD expands to
// nothing under g++) and runs the per-ray loop of k_trace serially, so that the CUDA path's math can
// be compared with the oracle before a GPU run.  Only tests/ build and load this library.
#include <cstdio>
#include <cstring>

#include "../../robast_b200/csrc/rb_build.h"
#include "../../robast_b200/csrc/rb_device.cuh"

namespace {
// the typed two-primitive booleans the scene-specialised kernels carry (rb_trace_v_*.cu), so that the host build runs them too
typedef Combos<B2<RBG_SHAPE_INTERSECTION, RBG_SHAPE_SPHERE, RBG_SHAPE_PGON>, B2<RBG_SHAPE_INTERSECTION, RBG_SHAPE_SPHERE, RBG_SHAPE_TUBE>,
               B2<RBG_SHAPE_SUBTRACTION, RBG_SHAPE_BBOX, RBG_SHAPE_BBOX>, B2<RBG_SHAPE_SUBTRACTION, RBG_SHAPE_PGON, RBG_SHAPE_WINSTONPOLY>,
               B2<RBG_SHAPE_SUBTRACTION, RBG_SHAPE_PARABOLOID, RBG_SHAPE_PARABOLOID>, B2<RBG_SHAPE_SUBTRACTION, RBG_SHAPE_PGON, RBG_SHAPE_PGON>,
               B2<RBG_SHAPE_SUBTRACTION, RBG_SHAPE_PCON, RBG_SHAPE_PCON>>
    EmulCombos;
template <class K> void run(const DScene& sc, const DTraceParams& tp, const rbg_rays* R, const rbg_history* H) {
  DHist dh;
  memset(&dh, 0, sizeof(dh));
  if (H && H->max_points > 0) {
    dh.x = H->hx; dh.y = H->hy; dh.z = H->hz; dh.t = H->ht; dh.node = H->hnode;
    dh.stride = R->n;
    dh.max_points = H->max_points;
  }
  for (long long idx = 0; idx < R->n; idx++) {
    RayReg r;
    r.lambda = R->lambda[idx];
    r.p = v3(R->x[idx], R->y[idx], R->z[idx]);
    r.t = R->t[idx];
    V3 d = v3(R->dx[idx], R->dy[idx], R->dz[idx]);
    double mag = sqrt(dot(d, d));
    r.d = mag > 0 ? (1. / mag) * d : d;
    r.status = RBG_RUN; r.npoints = 1; r.last_node = -1; r.ndraw = 0; r.on_boundary = 0;
    r.cur = locate_start<K>(sc, r.p);
    unsigned long long id = tp.ray_id_offset + (unsigned long long)idx;
    Philox g;
    g.k0 = (uint32_t)tp.seed; g.k1 = (uint32_t)(tp.seed >> 32); g.id0 = (uint32_t)id; g.id1 = (uint32_t)(id >> 32); g.ndraw = 0;
    HistSink sink;
    sink.h = &dh;
    sink.idx = idx;
    const HistSink* hs = dh.x ? &sink : nullptr;
    if (hs) { dh.x[idx] = r.p.x; dh.y[idx] = r.p.y; dh.z[idx] = r.p.z; dh.t[idx] = r.t; dh.node[idx] = -1; }
#ifdef RB_EMUL_STATS
    for (g_eval_step = 0; r.status == RBG_RUN; g_eval_step++) { RB_STAT(steps); trace_step<K>(sc, tp, r, g, hs); }
#else
    while (r.status == RBG_RUN) trace_step<K>(sc, tp, r, g, hs);
#endif
    R->ox[idx] = r.p.x; R->oy[idx] = r.p.y; R->oz[idx] = r.p.z; R->ot[idx] = r.t;
    R->odx[idx] = r.d.x; R->ody[idx] = r.d.y; R->odz[idx] = r.d.z;
    R->status[idx] = r.status; R->last_node[idx] = r.last_node; R->npoints[idx] = r.npoints;
  }
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int emul_trace_history(const rbg_scene_desc* D, const rbg_trace_opts* o, const rbg_rays* R,
                                                                         const rbg_history* H, int nthreads);
extern "C" __attribute__((visibility("default"))) int emul_trace(const rbg_scene_desc* D, const rbg_trace_opts* o, const rbg_rays* R, int nthreads) {
  return emul_trace_history(D, o, R, nullptr, nthreads);
}
extern "C" __attribute__((visibility("default"))) int emul_trace_history(const rbg_scene_desc* D, const rbg_trace_opts* o, const rbg_rays* R,
                                                                         const rbg_history* H, int /*nthreads*/) {
  try {
    validate_desc(D);
    SceneBuilder B;
    B.D = D;
    B.build_shapes();
    B.flatten(D->top_volume, mat_identity(), -1, 0, "top_1");
    DScene sc;
    memset(&sc, 0, sizeof(sc));
    sc.nodes = B.nodes.data(); sc.bvh = B.bvh.data(); sc.boxes = B.boxes.data(); sc.shapes = B.shapes.data(); sc.dpar = B.dpar.data(); sc.mats = B.mats.data();
    sc.volumes = D->volumes; sc.borders = D->borders; sc.graphs = D->graphs; sc.gx = D->gx; sc.gy = D->gy; sc.th2 = D->th2; sc.th2v = D->th2v;
    sc.indices = D->indices; sc.mirrors = D->mirrors; sc.focals = D->focals; sc.multilayers = D->multilayers; sc.layers = D->layers;
    sc.graph2d = D->graph2d; sc.tri = D->tri; sc.g2x = D->g2x; sc.g2y = D->g2y; sc.g2z = D->g2z;
    sc.nnodes = (int)B.nodes.size();
    sc.top_shape = D->volumes[D->top_volume].shape;
    sc.top_leaf = B.leaf_kind(sc.top_shape);
    for (const DNode& nd : B.nodes) sc.has_many |= nd.overlap != 0;
    DTraceParams tp;
    tp.limit = o->limit > 0 ? o->limit : 100; tp.disable_fresnel = o->disable_fresnel; tp.quirks = o->quirks; tp.max_steps = 0;
    tp.seed = o->seed; tp.ray_id_offset = o->ray_id_offset;
    switch (scene_depth_needed(B)) {
      case 0: run<TraceCfg<0, RB_SHAPES_ALL, RB_PH_ALL, 128, 4, EmulCombos>>(sc, tp, R, H); break;
      case 1: run<TraceCfg<1, RB_SHAPES_ALL, RB_PH_ALL, 128, 4, EmulCombos>>(sc, tp, R, H); break;
      case 2: run<TraceCfg<2, RB_SHAPES_ALL, RB_PH_ALL, 128, 4, EmulCombos>>(sc, tp, R, H); break;
      case 3: run<TraceCfg<3, RB_SHAPES_ALL, RB_PH_ALL, 128, 4, EmulCombos>>(sc, tp, R, H); break;
      case 4: case 5: case 6: run<TraceCfg<6, RB_SHAPES_ALL, RB_PH_ALL, 128, 4, EmulCombos>>(sc, tp, R, H); break;
      default: return RBG_ENOTSUP;
    }
    return RBG_OK;
  } catch (std::exception& e) {
    fprintf(stderr, "emul_trace: %s\n", e.what());
    return RBG_EINTERNAL;
  }
}
#ifdef RB_EMUL_STATS
extern "C" __attribute__((visibility("default"))) void emul_stats(long long* out, int reset) {
  memcpy(out, &g_eval_stats, sizeof(g_eval_stats));
  if (reset) memset(&g_eval_stats, 0, sizeof(g_eval_stats));
}
#endif
extern "C" __attribute__((visibility("default"))) double emul_div(double a, double b) { return rb_div(a, b); }
extern "C" __attribute__((visibility("default"))) int emul_tmm(const rbg_scene_desc* D, int ml, int pol, double th, double lam, double* R, double* T) {
  DScene sc;
  memset(&sc, 0, sizeof(sc));
  sc.graphs = D->graphs; sc.gx = D->gx; sc.gy = D->gy; sc.th2 = D->th2; sc.th2v = D->th2v; sc.indices = D->indices;
  sc.multilayers = D->multilayers; sc.layers = D->layers;
  if (pol == 2) tmm_mixed(sc, ml, th, lam, *R, *T);
  else tmm_coherent(sc, ml, pol, th, lam, *R, *T);
  return 0;
}
extern "C" __attribute__((visibility("default"))) int emul_tmm_general(const rbg_scene_desc* D, int ml, int mode, int pol, int reverse, double th_re, double th_im,
                                                                      double lam, double* R, double* T) {
  DScene sc;
  memset(&sc, 0, sizeof(sc));
  sc.graphs = D->graphs; sc.gx = D->gx; sc.gy = D->gy; sc.th2 = D->th2; sc.th2v = D->th2v; sc.indices = D->indices;
  sc.multilayers = D->multilayers; sc.layers = D->layers;
  const rbg_multilayer M = sc.multilayers[ml];
  if (mode == 0) tmm_coherent_sub(sc, M.first, 0, M.n - 1, reverse != 0, pol, cx(th_re, th_im), lam, *R, *T);
  else tmm_incoherent(sc, ml, pol, cx(th_re, th_im), lam, *R, *T);
  return 0;
}
