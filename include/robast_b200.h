/*
 * robast_b200.h — C ABI of the B200-native non-sequential ray tracer.
 *
 * This is the drop-in boundary for ONE path of ROBAST:
 *   AOpticsManager::TraceNonSequential(ARayArray&)   (reference src/AOpticsManager.cxx:523-587)
 * The reference has no FFI layer of its own: the boundary is the C++ member
 * overload set in include/AOpticsManager.h:92-100.  The entry points below are what a thin
 * AOpticsManager/ARayArray/ARayShooter C++ layer (include/robast/, shipped here) or a ROOT-side
 * exporter binds.  Plain pointers and sizes only; never throws; every function
 * returns 0 on success or a negative RBG_E* code, with text in rbg_last_error().
 *
 * Units follow the reference (include/AOpticsManager.h:59-71): lengths in cm, time in s,
 * wavelengths are lengths in cm, angles in rad unless a TGeo shape parameter is in degrees.
 */
#ifndef ROBAST_B200_H
#define ROBAST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBG_ABI_VERSION 1

/* ---------------------------------------------------------------- error codes */
#define RBG_OK 0
#define RBG_EINVAL (-1)      /* bad argument / malformed scene */
#define RBG_ENOTSUP (-2)     /* shape or option not supported on the device path */
#define RBG_ECUDA (-3)       /* CUDA runtime error (text in rbg_last_error) */
#define RBG_ENOMEM (-4)
#define RBG_EINTERNAL (-5)   /* caught C++ exception at the boundary */

/* ---------------------------------------------------------------- enums */
/* volume classification, reference include/AOpticsManager.h:45-53 (exact class identity :76-90) */
enum { RBG_LENS = 0, RBG_OBS = 1, RBG_MIRROR = 2, RBG_FOCUS = 3, RBG_OPT = 4, RBG_OTHER = 5, RBG_NULL = 6 };

/* ray status, enum order of reference include/ARay.h:26 */
enum { RBG_RUN = 0, RBG_STOP = 1, RBG_EXIT = 2, RBG_FOCUSED = 3, RBG_SUSPEND = 4, RBG_ABSORB = 5 };

/* shapes: ROOT primitives used by the five configs + the reference's own AGeo* shapes */
enum {
  RBG_SHAPE_BBOX = 0,        /* TGeoBBox      dpar: dx,dy,dz,ox,oy,oz */
  RBG_SHAPE_TUBE = 1,        /* TGeoTube      dpar: rmin,rmax,dz */
  RBG_SHAPE_SPHERE = 2,      /* TGeoSphere    dpar: rmin,rmax,theta1,theta2,phi1,phi2 (deg) */
  RBG_SHAPE_PARABOLOID = 3,  /* TGeoParaboloid dpar: rlo,rhi,dz */
  RBG_SHAPE_PGON = 4,        /* TGeoPgon      dpar: phi1,dphi(deg),nedges,nz, nz x (z,rmin,rmax) */
  RBG_SHAPE_PCON = 5,        /* TGeoPcon      dpar: phi1,dphi(deg),nz, nz x (z,rmin,rmax) */
  RBG_SHAPE_ASPHERE = 6,     /* AGeoAsphericDisk (src/AGeoAsphericDisk.cxx) dpar:
                                z1,z2,curve1,curve2,kappa1,kappa2,rmin,rmax,npol1,npol2,
                                bbox_origin_z,bbox_dz, K1[npol1], K2[npol2] */
  RBG_SHAPE_WINSTON2D = 7,   /* AGeoWinstonCone2D dpar: r1,r2,dy (r1>r2>0) */
  RBG_SHAPE_WINSTONPOLY = 8, /* AGeoWinstonConePoly dpar: r1,r2,npoly */
  RBG_SHAPE_UNION = 9,       /* TGeoCompositeShape / TGeoUnion        left,right,lmat,rmat */
  RBG_SHAPE_INTERSECTION = 10,
  RBG_SHAPE_SUBTRACTION = 11,
  RBG_SHAPE_ARB8 = 12,       /* TGeoArb8      dpar: dz, 8 x (x,y): vertices 0-3 at -dz, 4-7 at +dz (tutorials/AshraOptics.C:264,
                                src/AGeoUtil.cxx:47-82); may be twisted, vertices may coincide */
  RBG_SHAPE_XTRU = 13        /* TGeoXtru      dpar: nvert,nz, nvert x (x,y), nz x (z,x0,y0,scale) (tutorials/AshraOptics.C:791,
                                src/AGeoUtil.cxx:84-125) */
};

/* refractive-index kinds, reference include/ARefractiveIndex.h:36-65 and the formula classes */
enum {
  RBG_INDEX_GRAPH = 0,     /* n(λ), k(λ) by TGraph::Eval (1 point = constant) */
  RBG_INDEX_SELLMEIER = 1, /* src/ASellmeierFormula.cxx:46-54  par = B1,B2,B3,C1,C2,C3 */
  RBG_INDEX_SCHOTT = 2,    /* src/ASchottFormula.cxx:43-55     par = A0..A5 */
  RBG_INDEX_CAUCHY = 3,    /* src/ACauchyFormula.cxx:40-46     par = A,B,C */
  RBG_INDEX_MIXED = 4      /* include/AMixedRefractiveIndex.h:36-45 */
};

/* ---------------------------------------------------------------- flat scene tables (POD) */
typedef struct {
  int32_t type;         /* RBG_SHAPE_* */
  int32_t ipar, npar;   /* slice of dpar[] */
  int32_t left, right;  /* boolean operands (shape ids) or -1 */
  int32_t lmat, rmat;   /* operand placements (matrix ids) or -1 = identity */
} rbg_shape;

typedef struct {        /* TGeoMatrix: master = rot * local + tr */
  double rot[9];        /* row major */
  double tr[3];
} rbg_matrix;

typedef struct {        /* a placed daughter (TGeoNode) inside its mother volume */
  int32_t volume;
  int32_t matrix;       /* -1 = identity */
  int32_t copy_no;
  int32_t overlap;      /* AddNodeOverlap */
} rbg_node;

typedef struct {        /* TGeoVolume / AOpticalComponent */
  int32_t type;         /* RBG_LENS ... RBG_OTHER */
  int32_t shape;
  int32_t index;        /* lens: refractive index id or -1 (n=1,k=0)  src/ALens.cxx:36-60 */
  int32_t mirror;       /* mirror: reflectance id or -1 (R=1)         src/AMirror.cxx:39-60 */
  int32_t focal;        /* focal surface: QE id or -1 (QE=1)          src/AFocalSurface.cxx:35-52 */
  int32_t first_node, nnodes;     /* daughters: slice of nodes[] in AddNode order */
  int32_t first_border, nborders; /* slice of borders[] registered on this volume, insertion order */
  int32_t name;         /* offset into names[] (NUL terminated) */
} rbg_volume;

typedef struct {        /* ABorderSurfaceCondition (include/ABorderSurfaceCondition.h:24-48) */
  int32_t vol2;         /* component2, -1 = null */
  int32_t multilayer;   /* -1 = none */
  int32_t lambertian;
  int32_t pad;
  double sigma;         /* Gaussian roughness (rad), stored as |sigma| */
} rbg_border;

typedef struct { int32_t first, n; } rbg_graph; /* TGraph: slice of gx[],gy[] sorted by x */

typedef struct {        /* TH2D with uniform bins; content(i,j) = th2v[first + i + nx*j], i,j 0-based */
  int32_t first, nx, ny, pad;
  double xmin, xmax, ymin, ymax;
} rbg_th2;

typedef struct {
  int32_t kind;            /* RBG_INDEX_* */
  int32_t ngraph, kgraph;  /* graph ids or -1 (n=1 / k=0) */
  int32_t mix_a, mix_b;    /* RBG_INDEX_MIXED operands */
  int32_t pad;
  double par[6];
  double frac_a, frac_b;
} rbg_index;

typedef struct {        /* AMirror reflectance, priority graph2d > th2 > graph1d > constant */
  double constant;
  int32_t graph1d, th2, graph2d, pad;
} rbg_mirror;

typedef struct { int32_t qe_lambda, qe_angle; } rbg_focal; /* graph ids or -1 */

typedef struct {        /* AMultilayer: layers[first .. first+n) top(0) ... bottom(n-1) */
  int32_t first, n;
  int32_t table_r, table_t; /* PreCalculateCoherentTMM tables (th2 ids) or -1 */
} rbg_multilayer;

/* thickness cm, +inf for the two ends; incoherent != 0 marks a layer as incoherent for IncoherentTMM (the ends always are) */
typedef struct { int32_t index; int32_t incoherent; double thickness; } rbg_layer;

typedef struct {        /* TGraph2D baked to a Delaunay triangle list (x,y,z per vertex) */
  int32_t first_tri, ntri; /* slice of tri[] (3 vertex ids each, into g2x/g2y/g2z) */
} rbg_graph2d;

typedef struct rbg_scene_desc {
  int32_t abi_version;
  int32_t top_volume;
  int32_t nshapes, nmatrices, nnodes, nvolumes, nborders, ngraphs, nth2, nindices, nmirrors,
      nfocals, nmultilayers, nlayers, ngraph2d;
  int32_t ndpar, ngpts, nth2v, nnames, ntri, ng2pts;
  const rbg_shape* shapes;
  const double* dpar;
  const rbg_matrix* matrices;
  const rbg_node* nodes;
  const rbg_volume* volumes;
  const rbg_border* borders;
  const rbg_graph* graphs;
  const double* gx;
  const double* gy;
  const rbg_th2* th2;
  const double* th2v;
  const rbg_index* indices;
  const rbg_mirror* mirrors;
  const rbg_focal* focals;
  const rbg_multilayer* multilayers;
  const rbg_layer* layers;
  const rbg_graph2d* graph2d;
  const int32_t* tri;
  const double* g2x;
  const double* g2y;
  const double* g2z;
  const char* names;
} rbg_scene_desc;

/* ---------------------------------------------------------------- trace options */
/* quirks: ROOT/ROBAST behaviours reproduced by default */
#define RBG_QUIRK_STEPBACK 1u     /* reflection vertex recorded 2e-6 cm short, AOpticsManager.cxx:238-245 */
#define RBG_QUIRK_BOUNDARY_PUSH 2u /* TGeoNavigator 1e-10 cm push when starting on a boundary */
#define RBG_QUIRKS_DEFAULT (RBG_QUIRK_STEPBACK | RBG_QUIRK_BOUNDARY_PUSH)

typedef struct {
  int32_t limit;            /* fLimit (SetLimit), default 100; <=0 -> 100 */
  int32_t disable_fresnel;  /* DisableFresnelReflection */
  uint32_t quirks;          /* RBG_QUIRK_* */
  int32_t steps_per_launch; /* wavefront granularity: boundary steps per bounce kernel; 0 = auto, <0 = until done (one launch) */
  uint64_t seed;            /* Philox key */
  uint64_t ray_id_offset;   /* global index of ray 0 (multi-GPU sharding keeps streams independent of G) */
} rbg_trace_opts;

/* SoA ray batch.  All pointers are either all host or all device (see `on_device`).
 * Inputs: x,y,z,t,dx,dy,dz,lambda (64 B/ray).  Outputs (same order as inputs, 68 B/ray):
 * ox,oy,oz,ot (last point), odx,ody,odz (final direction), status, last_node (placed-node id,
 * -1 none), npoints.  Output pointers may alias the inputs (in-place). */
typedef struct {
  int64_t n;
  int32_t on_device;
  int32_t pad;
  const double *x, *y, *z, *t, *dx, *dy, *dz, *lambda;
  double *ox, *oy, *oz, *ot, *odx, *ody, *odz;
  int32_t *status, *last_node, *npoints;
} rbg_rays;

typedef struct rbg_scene rbg_scene;

/* ---------------------------------------------------------------- library */
int rbg_abi_version(void);
const char* rbg_last_error(void);
/* number of CUDA devices visible (0 when none; never fails) */
int rbg_device_count(void);

/* Deep-copies the flat tables, flattens placed nodes into physical paths, builds the BVH and
 * uploads everything to `device`.  Replaces per-thread TGeoNavigator setup
 * (src/AOpticsManager.cxx:336-344). */
int rbg_scene_create(const rbg_scene_desc* desc, int device, rbg_scene** out);
int rbg_scene_destroy(rbg_scene* scene);
/* physical (flattened) node table: count, and "<volname>_<copyNo>" name of node i (ROOT naming) */
int rbg_scene_num_nodes(const rbg_scene* scene);
/* name of the k_trace instantiation selected for this scene (scene-specialised or generic_dN) */
const char* rbg_scene_kernel_variant(const rbg_scene* scene);
const char* rbg_scene_node_name(const rbg_scene* scene, int node);

/* TraceNonSequential over a ray batch (src/AOpticsManager.cxx:335-520,523-587).
 * `stream` is a cudaStream_t (NULL = default stream). With host pointers the call performs the
 * H2D/D2H copies itself, chunked and overlapped, and returns after completion.  With device
 * pointers the whole trace — every bounce, the compactions between them, the scratch it needs
 * (stream-ordered allocation) — is enqueued on `stream` and the call returns without waiting for the
 * device: the survivor counts that size each bounce stay in device memory.  Several calls on the same
 * scene may be in flight on different streams.  (One exception: the very first large batch a scene sees,
 * when the scene has 32 or more daughters under its top volume, waits once for the probe that tells a
 * coherent beam from a scattered one.) */
int rbg_trace(rbg_scene* scene, const rbg_trace_opts* opts, const rbg_rays* rays, void* stream);

/* Optional polyline record of every ray (ARay's TGeoTrack points and node history, reference
 * include/ARay.h:24-68, src/ARay.cxx:66-71): point k of ray i is written to h?[k*n + i], k = 0 is the
 * start point; hnode holds the placed-node id recorded by ARay::AddNode with point k (-1 for k = 0).
 * Rays keep at most max_points points (rays->npoints still counts all of them); entries k >= npoints are
 * left untouched.  Same residency as the ray arrays (host or device).  A trace with a history runs as one
 * per-ray-loop launch per chunk (no wavefront compaction); it is opt-in and not part of the rays/s metric. */
typedef struct {
  int32_t max_points;
  int32_t pad;
  double *hx, *hy, *hz, *ht;
  int32_t* hnode;
} rbg_history;
int rbg_trace_history(rbg_scene* scene, const rbg_trace_opts* opts, const rbg_rays* rays,
                      const rbg_history* history, void* stream);

/* kernel-launch counter for the calling process (bench.py reports it as gpu_launches) */
int64_t rbg_launch_count(void);
/* timing of the dominant kernel: enable, then read accumulated ms and launch count of bounce kernels */
int rbg_profile_enable(int on);
int rbg_profile_read(double* bounce_ms, int64_t* bounce_launches, double* compact_ms, int64_t* compact_launches);

/* ARayShooter on device (src/ARayShooter.cxx:122-460): fills device SoA x..lambda for n rays.
 * kind: 0 Rectangle(grid nx*ny, x-major), 1 RandomRectangle, 2 RandomCircle, 3 Circle(nr,nphi),
 * 4 RandomCone (point source at tr aimed at the disc of radius dx at z=dy, rotated), 5 RandomSphere
 * (isotropic point source at tr), 6 RandomSphericalCone (point source, uniform within dx DEGREES of rot*z).
 * rot (9, row major) and tr (3) may be NULL; dir (3) NULL = (0,0,1).  lambda_min==lambda_max
 * gives a monochromatic beam, else uniform per ray.  `first` is the global index of the first
 * ray generated (sharding / chunking). */
typedef struct {
  int32_t kind;
  int32_t nx, ny;
  int32_t pad;
  double dx, dy;          /* full widths (Rectangle) or rmax in dx (circle kinds) */
  double lambda_min, lambda_max;
  double rot[9];
  double tr[3];
  double dir[3];
  uint64_t seed;
} rbg_shoot_desc;
int rbg_shoot(const rbg_shoot_desc* d, int64_t first, int64_t n, double* x, double* y, double* z,
              double* t, double* dx, double* dy, double* dz, double* lambda, int device, void* stream);

/* ---------------------------------------------------------------- several GPUs of one box, one process
 * The reference fans a batch out over threads inside TraceNonSequential (src/AOpticsManager.cxx:529-568: contiguous chunks,
 * one TGeoNavigator per thread, ordered Merge).  Here the chunks go to GPUs: the geometry is replicated on every device,
 * device k traces rays [k n/G, (k+1) n/G) (the last one takes the remainder, :533-541), Philox ray ids are global, so the
 * result is bit-identical to the single-GPU one whatever G is. */
typedef struct rbg_multi rbg_multi;
/* devices == NULL: devices 0 .. ndev-1 */
int rbg_multi_create(const rbg_scene_desc* desc, int ndev, const int* devices, rbg_multi** out);
int rbg_multi_destroy(rbg_multi* m);
int rbg_multi_num_devices(const rbg_multi* m);
/* TraceNonSequential over a HOST batch, sharded over the devices (one host thread per device, each running the chunked
 * H2D -> trace -> D2H pipeline of rbg_trace on its range) */
int rbg_multi_trace(rbg_multi* m, const rbg_trace_opts* opts, const rbg_rays* host_rays);
/* The 1e9-ray pattern of BASELINE configs[4]: generate (rbg_shoot), trace and reduce on the devices — n_total rays of `shoot`
 * split over the devices, at most `batch` rays resident per device at a time, nothing but the reducers ever leaves a GPU.
 * hist (nx*ny uint64, bin (i,j) at i + nx*j) is the PSF histogram of the rays with status `sel`; it lives on the first device
 * and every device's histogram kernel adds its block-private bins straight into it through peer memory (NVLink atomics: the
 * cross-GPU reduction is part of the histogram kernel, there is no separate collective; without peer access the partial
 * histograms are copied over and added).  moments: the 8 doubles of rbg_moments; counts: 6 status counters.  All three are
 * HOST pointers, written when the call returns. */
int rbg_multi_shoot_trace_reduce(rbg_multi* m, const rbg_trace_opts* opts, const rbg_shoot_desc* shoot, int64_t n_total,
                                 int64_t batch, int32_t sel, int32_t nx, double xmin, double xmax, int32_t ny, double ymin,
                                 double ymax, unsigned long long* hist, double* moments, long long* counts);

/* CORSIKA IACT photon bunches -> rays on the device: ACorsikaIACTFile::GetRayArray (src/ACorsikaIACTFile.cxx:71-133) for the
 * bunches of one telescope.  Bunch arrays are HOST pointers in CORSIKA units (x,y cm at the observation level relative to the
 * telescope, time ns, direction cosines cx,cy,cz with cz < 0, lambda nm or 0 = undetermined, photons = bunch size); the rays are
 * written to DEVICE arrays in the tracer's units (cm, s).  Bunch i yields the rays j = 0,1,.. while j < photons[i] (the reference's
 * loop, i.e. ceil(photons) rays), all starting at (x - d cx, y - d cy, z) with d = (z - telescope_z) * (-1/cz) at time
 * t*ns - d/(c/n).  A bunch with lambda == 0 draws 1/lambda uniformly between 1/lambda_min and 1/lambda_max (Philox stream of the
 * global ray index, draw 0).  NB the reference passes fMaxPhotonBunches where fMaxWavelength is meant (:123-125); the caller
 * decides what to pass as lambda_max_nm.  rbg_bunch_rays counts the rays; rbg_shoot_bunches fills rays [first, first+n). */
typedef struct {
  int64_t nbunches;
  const float *x, *y, *time, *cx, *cy, *cz, *lambda, *photons;
  double z;                 /* start height of the rays above the CORSIKA observation level (cm) */
  double telescope_z;       /* ACorsikaIACTFile::GetTelescopeZ(telNo) (cm) */
  double refractive_index;  /* of the air above the telescope */
  double lambda_min_nm, lambda_max_nm;
  uint64_t seed;
} rbg_bunches;
int rbg_bunch_rays(const rbg_bunches* b, int64_t* nrays);
int rbg_shoot_bunches(const rbg_bunches* b, int64_t first, int64_t n, double* x, double* y, double* z, double* t, double* dx,
                      double* dy, double* dz, double* lambda, int device, void* stream);

/* On-device reducers replacing the user-side TH2D fill + GetMean/GetRMS loops
 * (tutorials/SimpleParabolicTelescope.C:114-172).  Only rays with status==sel are used.
 * hist: nx*ny uint64 bins (device), under/overflow dropped.  moments: 8 doubles (device):
 * n, Σx, Σy, Σx², Σy², Σt, Σt², unused.  counts: 6 int64 status counters (device). */
int rbg_hist2d(int64_t n, const double* x, const double* y, const int32_t* status, int32_t sel,
               int32_t nx, double xmin, double xmax, int32_t ny, double ymin, double ymax,
               unsigned long long* hist, int device, void* stream);
/* rbg_hist2d of (x - x0, y - y0) that also accumulates the TH2 statistics of the in-range fills into
 * stats[5] (device): Σw, Σx', Σy', Σx'², Σy'² of the shifted coordinates — what TH2::GetMean/GetStdDev use.
 * stats may be NULL.  The origin shift lets the PSFs of several field angles share one binning. */
int rbg_hist2d_stats(int64_t n, const double* x, const double* y, const int32_t* status, int32_t sel,
                     double x0, double y0, int32_t nx, double xmin, double xmax, int32_t ny, double ymin,
                     double ymax, unsigned long long* hist, double* stats, int device, void* stream);
/* AGeoUtil::ContainmentRadius (reference src/AGeoUtil.cxx:18-42,198-308; D80 of the tutorials, e.g.
 * tutorials/MST.C:231-281) on `nhist` device-resident histograms of identical binning: hist
 * (nhist x nx*ny uint64, bin (i,j) at i + nx*j, as written by rbg_hist2d), stats (nhist x 5 doubles
 * from rbg_hist2d_stats), out (nhist x 3 doubles: radius, centre x, centre y). One block per histogram. */
int rbg_containment_radius(int32_t nhist, const unsigned long long* hist, int32_t nx, double xmin,
                           double xmax, int32_t ny, double ymin, double ymax, const double* stats,
                           double fraction, double* out, int device, void* stream);
/* same for ONE host histogram with double bin contents (what AGeoUtil::ContainmentRadius(TH2*) binds);
 * bins, stats and out are host pointers; copies inside, synchronous */
int rbg_containment_radius_host(const double* bins, int32_t nx, double xmin, double xmax, int32_t ny,
                                double ymin, double ymax, const double* stats, double fraction,
                                double* out, int device);
int rbg_moments(int64_t n, const double* x, const double* y, const double* t, const int32_t* status,
                int32_t sel, double* moments, long long* counts, int device, void* stream);

/* AMultilayer::CoherentTMMMixed on device for arrays of (theta, lambda): one thread per pair
 * (include/AMultilayer.h:114-132,243-262).  Pointers are device pointers. */
int rbg_tmm(rbg_scene* scene, int multilayer, int64_t n, const double* theta, const double* lambda,
            double* refl, double* trans, void* stream);
/* AMultilayer::CoherentTMM (one polarisation, complex incidence angle, optionally the reversed stack;
 * src/AMultilayer.cxx:240-481) and AMultilayer::IncoherentTMM (partly coherent stacks, :484-731) for host arrays
 * of n (theta_re, theta_im, lambda) triples.  mode: 0 coherent, 1 incoherent.  pol: 0 = s, 1 = p, 2 = mean of both.
 * `reverse` applies to mode 0 only.  Copies inside, synchronous. */
int rbg_tmm_general_host(rbg_scene* scene, int multilayer, int mode, int pol, int reverse, int64_t n,
                         const double* theta_re, const double* theta_im, const double* lambda,
                         double* refl, double* trans);
/* same with host arrays (copies inside, synchronous) — what AMultilayer::CoherentTMMMixed binds */
int rbg_tmm_host(rbg_scene* scene, int multilayer, int64_t n, const double* theta, const double* lambda,
                 double* refl, double* trans);

#ifdef __cplusplus
}
#endif
#endif /* ROBAST_B200_H */
