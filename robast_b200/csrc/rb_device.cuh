// rb_device.cuh — per-ray fp64 math of the B200 tracer: analytic shapes (ROOT TGeo primitives +
// ROBAST AGeo* shapes + boolean composites), flattened navigation with a threaded BVH, surface
// physics (Snell/Fresnel, mirrors, Lambertian, Gaussian roughness, multilayer TMM, absorption, QE)
// and the Philox counter RNG.  One thread owns one ray; everything lives in registers.
//
// Reference behaviour being reproduced (file:line under /root/reference):
//   state machine            src/AOpticsManager.cxx:335-520   
//   DoFresnel / DoReflection src/AOpticsManager.cxx:52-247, GetFacetNormal :250-301
//   AGeoAsphericDisk         src/AGeoAsphericDisk.cxx:96-153,246-347,363-684
//   AGeoWinstonCone2D/Poly   src/AGeoWinstonCone2D.cxx:59-98,120-428 ; src/AGeoWinstonConePoly.cxx:67-226,293-309
//   AMultilayer              src/AMultilayer.cxx:26-56,120-209,240-481 ; include/AMultilayer.h:114-132
//   indices / mirrors / QE   src/A*Formula.cxx, include/ARefractiveIndex.h:36-65, src/AMirror.cxx:39-60,
//                            src/AFocalSurface.cxx:35-52, src/AOpticalComponent.cxx:51-65
//   ROOT TGeoNavigator / TGeo shapes / TGraph / TH2: external dependency, SURVEY.md Appendix B.
// All functions are RB_HD so that the same source also builds for the host in tests/emul (a
// debugging aid on the GPU-less build box; never part of the product library).
#ifndef RB_DEVICE_CUH
#define RB_DEVICE_CUH

#include <math.h>

#include "rb_scene.h"

// compile-time feature masks: a scene-specialised instantiation only carries the shapes / physics it needs
// (the generic kernel overflows the instruction cache, profiles/r1a_k_trace_ncu_summary.md)
#define RB_SBIT(t) (1u << (t))
#define RB_SHAPES_ALL 0xffffu
#define RB_PH_LENS 1u        /* ALens: Snell/Fresnel, bulk absorption, n(lambda) */
#define RB_PH_MULTILAYER 2u  /* AMultilayer on a border (TMM or table) */
#define RB_PH_ROUGH 4u       /* Gaussian facet roughness */
#define RB_PH_LAMBERT 8u     /* Lambertian border */
#define RB_PH_QE 16u         /* AFocalSurface QE graphs */
#define RB_PH_MIRROR_TABLE 32u /* AMirror reflectance != constant */
#define RB_PH_OVERLAP 64u     /* AddNodeOverlap ("MANY") nodes: overlap-cluster point location, sister candidates */
#define RB_PH_ALL 0xffu
template <class... Cs> struct Combos;
template <int D, unsigned S, unsigned P, int THREADS = 128, int MINB = 4, class CL = Combos<>> struct TraceCfg {
  typedef CL combos;  // two-primitive booleans evaluated by typed inline code (Bool2), e.g. Combos<B2<RBG_SHAPE_INTERSECTION, RBG_SHAPE_SPHERE, RBG_SHAPE_PGON>>
  static constexpr int depth = D;
  static constexpr unsigned shapes = S;
  static constexpr unsigned phys = P;
  static constexpr int threads = THREADS;   // k_trace: __launch_bounds__(THREADS, MINB) — block size and register cap
  static constexpr int min_blocks = MINB;
};

#define RB_BIG 1e30
#define RB_TOL 1e-10
#define RB_PI 3.14159265358979323846
#define RB_C_CM 2.99792458e10 /* TMath::C()*m() in cm/s */

// ------------------------------------------------------------------ small vector helpers
struct V3 {
  double x, y, z;
};
RB_HD inline V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RB_HD inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RB_HD inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RB_HD inline V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
RB_HD inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RB_HD inline V3 along(V3 p, V3 d, double t) { return v3(p.x + t * d.x, p.y + t * d.y, p.z + t * d.z); }
RB_HD inline double sqr(double v) { return v * v; }
RB_HD inline double rb_min(double a, double b) { return a < b ? a : b; }
RB_HD inline double rb_max(double a, double b) { return a > b ? a : b; }
RB_HD inline double rb_atan2(double y, double x) {  // TMath::ATan2
  if (x != 0) return atan2(y, x);
  if (y == 0) return 0;
  return y > 0 ? RB_PI / 2 : -RB_PI / 2;
}
// a / b, bit for bit.  The compiler's inline fp64 division handles a zero (or denormal) numerator through an out-of-line routine
// of ~90 instructions and ~20 local-memory accesses; numerators that are exactly zero are structural in this code (imaginary
// parts of lossless layers, flat surfaces, prisms of constant radius): 30 such calls per ray in the multilayer reflectance, a
// tenth of all instructions of the interaction kernel on BASELINE config 5 (profiles/r2_summary.md).  IEEE gives 0 / b = +-0 with
// the sign of a xor b for every b but 0 and NaN.
RB_HD inline double rb_div(double a, double b) { return (a == 0. && b != 0. && b == b) ? a * copysign(1., b) : a / b; }
RB_HD inline double rb_acos(double x) { return x < -1. ? RB_PI : (x > 1. ? 0 : acos(x)); }
RB_HD inline double rb_asin(double x) { return x < -1. ? -RB_PI / 2 : (x > 1. ? RB_PI / 2 : asin(x)); }

RB_HD inline V3 to_local(const DMat& m, V3 p) {
  double a = p.x - m.t[0], b = p.y - m.t[1], c = p.z - m.t[2];
  return v3(a * m.r[0] + b * m.r[3] + c * m.r[6], a * m.r[1] + b * m.r[4] + c * m.r[7], a * m.r[2] + b * m.r[5] + c * m.r[8]);
}
RB_HD inline V3 to_local_vec(const DMat& m, V3 d) {
  return v3(d.x * m.r[0] + d.y * m.r[3] + d.z * m.r[6], d.x * m.r[1] + d.y * m.r[4] + d.z * m.r[7], d.x * m.r[2] + d.y * m.r[5] + d.z * m.r[8]);
}
RB_HD inline V3 to_master_vec(const DMat& m, V3 l) {
  return v3(l.x * m.r[0] + l.y * m.r[1] + l.z * m.r[2], l.x * m.r[3] + l.y * m.r[4] + l.z * m.r[5], l.x * m.r[6] + l.y * m.r[7] + l.z * m.r[8]);
}

// ------------------------------------------------------------------ Philox4x32-10, key=(seed), ctr=(ray id, draw#)
struct Philox {
  uint32_t k0, k1, id0, id1, ndraw;
};
RB_HD inline uint32_t rb_mulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
RB_HD inline void philox_block(Philox& g, uint32_t o[4]) {
  uint32_t c0 = g.id0, c1 = g.id1, c2 = g.ndraw++, c3 = 0u, k0 = g.k0, k1 = g.k1;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = rb_mulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0, h1 = rb_mulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
RB_HD inline double philox_u53(uint32_t a, uint32_t b) { return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) / 9007199254740992.0; }
RB_HD inline double rng_uniform(Philox& g) {
  uint32_t o[4];
  philox_block(g, o);
  return philox_u53(o[0], o[1]);
}
RB_HD inline double rng_gaus(Philox& g, double mean, double sigma) {
  uint32_t o[4];
  philox_block(g, o);
  double u1 = philox_u53(o[0], o[1]), u2 = philox_u53(o[2], o[3]);
  return mean + sigma * sqrt(-2. * log(u1)) * cos(2 * RB_PI * u2);
}

// ================================================================== tables: TGraph::Eval, TH2::Interpolate
RB_HD inline int rb_min_i(int a, int b) { return a < b ? a : b; }
RB_HD inline double graph_eval(const DScene& sc, int g, double x) {
  const rbg_graph gr = sc.graphs[g];
  const double *X = sc.gx + gr.first, *Y = sc.gy + gr.first;
  int n = gr.n;
  if (n == 0) return 0;
  if (n == 1) return Y[0];
  // points are sorted by x at scene build: bracket by binary search, extrapolate from the end pairs
  int lo = 0, hi = n - 1;
  if (x <= X[0]) { if (x == X[0]) return Y[0]; lo = 0; hi = 1; }
  else if (x >= X[n - 1]) { if (x == X[n - 1]) return Y[n - 1]; lo = n - 2; hi = n - 1; }
  else {
    // tabulated optical constants sit on a (nearly) uniform grid: try the bracket a uniform grid would give before bisecting —
    // the bracket is unique (largest lo with X[lo] <= x), so the value is the same either way, after 2 loads instead of log2(n)
    // dependent ones
    const double gf = (x - X[0]) / (X[n - 1] - X[0]) * (double)(n - 1);
    const int g0 = gf >= 0. ? rb_min_i(n - 2, (int)gf) : 0;  // (a NaN abscissa takes the first bracket, as the bisection would)
    if (X[g0] <= x && x < X[g0 + 1]) { lo = g0; hi = g0 + 1; }
    else {
      if (X[g0] <= x) lo = g0; else if (g0 > 0) hi = g0;
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (X[mid] <= x) lo = mid; else hi = mid;
      }
    }
    if (X[lo] == x) return Y[lo];
  }
  if (X[lo] == X[hi]) return Y[lo];
  return Y[hi] + (x - X[hi]) * (Y[lo] - Y[hi]) / (X[lo] - X[hi]);
}

RB_HD inline int th2_findbin(double x, double lo, double hi, int n) { return x < lo ? 0 : (!(x < hi) ? n + 1 : 1 + int(n * (x - lo) / (hi - lo))); }
RB_HD inline double th2_interp(const DScene& sc, int h, double x, double y) {
  const rbg_th2 H = sc.th2[h];
  const double* v = sc.th2v + H.first;
  double wx = (H.xmax - H.xmin) / H.nx, wy = (H.ymax - H.ymin) / H.ny;
  int bx = th2_findbin(x, H.xmin, H.xmax, H.nx), by = th2_findbin(y, H.ymin, H.ymax, H.ny);
  if (bx < 1 || bx > H.nx || by < 1 || by > H.ny) return 0;
  double ddx = (H.xmin + bx * wx) - x, ddy = (H.ymin + by * wy) - y;
  int ix1 = ddx <= wx / 2 ? bx : bx - 1, iy1 = ddy <= wy / 2 ? by : by - 1;
  double x1 = H.xmin + (ix1 - 0.5) * wx, x2 = H.xmin + (ix1 + 0.5) * wx, y1 = H.ymin + (iy1 - 0.5) * wy, y2 = H.ymin + (iy1 + 0.5) * wy;
  int bx1 = ix1 < 1 ? 1 : ix1, bx2 = ix1 + 1 > H.nx ? H.nx : ix1 + 1, by1 = iy1 < 1 ? 1 : iy1, by2 = iy1 + 1 > H.ny ? H.ny : iy1 + 1;
  double q11 = v[(bx1 - 1) + H.nx * (by1 - 1)], q12 = v[(bx1 - 1) + H.nx * (by2 - 1)], q21 = v[(bx2 - 1) + H.nx * (by1 - 1)],
         q22 = v[(bx2 - 1) + H.nx * (by2 - 1)];
  double dd = 1.0 * (x2 - x1) * (y2 - y1);
  return 1.0 * q11 / dd * (x2 - x) * (y2 - y) + 1.0 * q21 / dd * (x - x1) * (y2 - y) + 1.0 * q12 / dd * (x2 - x) * (y - y1) +
         1.0 * q22 / dd * (x - x1) * (y - y1);
}

// TGraph2D::Interpolate on the Delaunay triangle list baked by the exporter: the first triangle (list order) whose
// barycentric coordinates are all >= -1e-9 interpolates linearly; outside the convex hull the value is 0 (fZout)
RB_HD inline double graph2d_interp(const DScene& sc, int g, double x, double y) {
  const rbg_graph2d G = sc.graph2d[g];
  for (int k = 0; k < G.ntri; k++) {
    const int32_t* t = sc.tri + 3 * (G.first_tri + k);
    double x0 = sc.g2x[t[0]], y0 = sc.g2y[t[0]], x1 = sc.g2x[t[1]], y1 = sc.g2y[t[1]], x2 = sc.g2x[t[2]], y2 = sc.g2y[t[2]];
    double den = (y1 - y2) * (x0 - x2) + (x2 - x1) * (y0 - y2);
    if (den == 0) continue;
    double l0 = ((y1 - y2) * (x - x2) + (x2 - x1) * (y - y2)) / den, l1 = ((y2 - y0) * (x - x2) + (x0 - x2) * (y - y2)) / den, l2 = 1. - l0 - l1;
    if (l0 < -1e-9 || l1 < -1e-9 || l2 < -1e-9) continue;
    return l0 * sc.g2z[t[0]] + l1 * sc.g2z[t[1]] + l2 * sc.g2z[t[2]];
  }
  return 0.;
}

// ================================================================== refractive index n(λ), k(λ)
RB_HD inline double index_k1(const DScene& sc, int id, double lambda) {
  if (id < 0) return 0.;
  int kg = sc.indices[id].kgraph;
  return kg >= 0 ? graph_eval(sc, kg, lambda) : 0.;
}
RB_HD inline double index_n1(const DScene& sc, int id, double lambda) {
  if (id < 0) return 1.;
  const rbg_index& x = sc.indices[id];
  const double* p = x.par;
  double l = lambda / 1e-4;  // cm -> µm
  if (x.kind == RBG_INDEX_SELLMEIER) {
    double l2 = l * l;
    return sqrt(1 + p[0] * l2 / (l2 - p[3]) + p[1] * l2 / (l2 - p[4]) + p[2] * l2 / (l2 - p[5]));
  }
  if (x.kind == RBG_INDEX_SCHOTT) {
    double l2 = l * l, i2 = 1. / l2, i4 = i2 * i2;
    return sqrt(p[0] + p[1] * l2 + p[2] * i2 + p[3] * i4 + p[4] * i4 * i2 + p[5] * i4 * i4);
  }
  if (x.kind == RBG_INDEX_CAUCHY) {
    double i2 = 1. / (l * l);
    return p[0] + p[1] * i2 + p[2] * i2 * i2;
  }
  return x.ngraph >= 0 ? graph_eval(sc, x.ngraph, lambda) : 1.;
}
RB_HD inline double index_n(const DScene& sc, int id, double lambda) {
  if (id >= 0 && sc.indices[id].kind == RBG_INDEX_MIXED) {
    const rbg_index& x = sc.indices[id];
    return index_n1(sc, x.mix_a, lambda) * x.frac_a + index_n1(sc, x.mix_b, lambda) * x.frac_b;  // one mixing level
  }
  return index_n1(sc, id, lambda);
}
RB_HD inline double index_k(const DScene& sc, int id, double lambda) {
  if (id >= 0 && sc.indices[id].kind == RBG_INDEX_MIXED) {
    const rbg_index& x = sc.indices[id];
    return index_k1(sc, x.mix_a, lambda) * x.frac_a + index_k1(sc, x.mix_b, lambda) * x.frac_b;
  }
  return index_k1(sc, id, lambda);
}

// ================================================================== complex helpers + coherent TMM
struct Cx {
  double re, im;
};
RB_HD inline Cx cx(double r, double i) { Cx c; c.re = r; c.im = i; return c; }
RB_HD inline Cx operator+(Cx a, Cx b) { return cx(a.re + b.re, a.im + b.im); }
RB_HD inline Cx operator-(Cx a, Cx b) { return cx(a.re - b.re, a.im - b.im); }
RB_HD inline Cx operator*(Cx a, Cx b) { return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
RB_HD inline Cx operator*(double s, Cx a) { return cx(s * a.re, s * a.im); }
RB_HD inline Cx operator/(Cx a, Cx b) {  // Smith's algorithm
  if (fabs(b.re) >= fabs(b.im)) {
    double r = rb_div(b.im, b.re), den = b.re + b.im * r;
    return cx(rb_div(a.re + a.im * r, den), rb_div(a.im - a.re * r, den));
  }
  double r = rb_div(b.re, b.im), den = b.re * r + b.im;
  return cx(rb_div(a.re * r + a.im, den), rb_div(a.im * r - a.re, den));
}
RB_HD inline double cabs2(Cx a) { return a.re * a.re + a.im * a.im; }
RB_HD inline Cx cconj(Cx a) { return cx(a.re, -a.im); }
RB_HD inline Cx csqrt_(Cx z) {  // principal branch
  double m = hypot(z.re, z.im);
  if (m == 0) return cx(0, 0);
  double s = sqrt(0.5 * (m + fabs(z.re)));
  double o = rb_div(z.im, 2 * s);
  if (z.re >= 0) return cx(s, o);
  return cx(fabs(o), z.im >= 0 ? s : -s);
}
// the transcendental helpers are not inlined: each is hundreds of instructions and the TMM calls them from several places
RB_HD inline RB_NOINLINE Cx cexp_(Cx z) {
  double e = exp(z.re);
  return cx(e * cos(z.im), e * sin(z.im));
}
RB_HD inline Cx clog_(Cx z) { return cx(log(hypot(z.re, z.im)), atan2(z.im, z.re)); }
RB_HD inline RB_NOINLINE Cx ccos_(Cx z) { return cx(cos(z.re) * cosh(z.im), -sin(z.re) * sinh(z.im)); }
RB_HD inline Cx csin_(Cx z) { return cx(sin(z.re) * cosh(z.im), cos(z.re) * sinh(z.im)); }
RB_HD inline RB_NOINLINE Cx casin_(Cx z) {  // asin z = -i log(i z + sqrt(1 - z^2))
  Cx w = csqrt_(cx(1, 0) - z * z);
  Cx l = clog_(cx(-z.im + w.re, z.re + w.im));
  return cx(l.im, -l.re);
}
RB_HD inline bool tmm_is_forward(Cx n, Cx theta) {
  Cx nc = n * ccos_(theta);
  if (fabs(nc.im) > 100 * 2.220446049250313e-16) return nc.im > 0;
  return nc.re > 0;
}
// AMultilayer::CoherentTMM (src/AMultilayer.cxx:240-481) for one polarisation (pol 0 = s, 1 = p) over the stack made of
// layers a..b (inclusive) of sc.layers[first ...], entered through layer a, or through layer b when `reverse`; th0 is the
// (possibly complex) angle in the entrance medium.
// NP = 1: polarisation `pol0`; NP = 2: both (index 0 = s, 1 = p) in one pass — refractive indices, Snell angles, their cosines
// and the layer phases do not depend on the polarisation, only the interface coefficients and the matrix products do, so the
// unpolarised CoherentTMMMixed costs one set of transcendentals instead of two.  Per polarisation the arithmetic is the same
// sequence of operations either way.
template <int NP> RB_HD inline void tmm_coherent_multi(const DScene& sc, int first, int a, int b, bool reverse, int pol0, Cx th0, double lam, double* R, double* T) {
  // r = M10 / M00 and t = 1 / M00 need only the first column of M = B_01 M_1 ... M_{N-2}: the product is applied to the unit
  // vector from the last layer towards the first — two complex numbers of state per polarisation instead of a 2x2 matrix plus
  // the first interface's coefficients (the kernels run at 64 registers; the matrix form lived in local memory).
  const int N = b - a + 1;
  const rbg_layer L0 = sc.layers[first + (reverse ? b : a)];
  const Cx n0 = cx(index_n(sc, L0.index, lam), index_k(sc, L0.index, lam));
  const Cx n0s = th0.im == 0 ? cx(n0.re * sin(th0.re), n0.im * sin(th0.re)) : n0 * csin_(th0);
  // Snell's law (AMultilayer::ListSnell): the reference takes theta_i = asin(n0 sin(theta0) / n_i), corrects the first and the
  // last layer to the forward-travelling solution (theta -> pi - theta) and then only ever uses cos(theta_i).  On the principal
  // branches cos(asin z) = sqrt(1 - z^2), and theta -> pi - theta flips the sign of the cosine: the same numbers without a
  // complex asin (sqrt, log, atan2) and a complex cos (cos, sin, cosh, sinh) per layer.
  auto snell_cos = [&](Cx ni, bool end) {
    const Cx zs = n0s / ni;
    Cx c = csqrt_(cx(1, 0) - zs * zs);
    if (end) {
      const Cx nc = ni * c;  // IsForwardAngle (src/AMultilayer.cxx:26-56)
      const bool forward = fabs(nc.im) > 100 * 2.220446049250313e-16 ? nc.im > 0 : nc.re > 0;
      if (!forward) c = cx(-c.re, -c.im);
    }
    return c;
  };
  const rbg_layer LN = sc.layers[first + (reverse ? a : b)];
  const Cx n_last = cx(index_n(sc, LN.index, lam), index_k(sc, LN.index, lam));
  const Cx c_last = snell_cos(n_last, true);
  Cx nf = n_last, cf = c_last;  // the layer behind the interface being added
  Cx v0[NP], v1[NP];
#pragma unroll
  for (int q = 0; q < NP; q++) { v0[q] = cx(1, 0); v1[q] = cx(0, 0); }
  for (int i = N - 2; i >= 0; i--) {
    const rbg_layer L = sc.layers[first + (reverse ? b - i : a + i)];
    const Cx ni = i == 0 ? n0 : cx(index_n(sc, L.index, lam), index_k(sc, L.index, lam));
    const Cx ci = snell_cos(ni, i == 0);
    const Cx ii = ni * ci;
    Cx em = cx(1, 0), ep = cx(1, 0);
    if (i > 0) {
      // an inner layer: M_i = (1/t) diag(e^{-i delta}, e^{i delta}) [[1,r],[r,1]]
      Cx delta = cx((2 * RB_PI) * ii.re / lam * L.thickness, rb_div((2 * RB_PI) * ii.im, lam) * L.thickness);
      if (delta.im > 35) delta.im = 35;
      em = cexp_(cx(delta.im, -delta.re));  // exp(-i delta)
      ep = cexp_(cx(-delta.im, delta.re));  // exp(i delta)
    }
#pragma unroll
    for (int q = 0; q < NP; q++) {
      const int pol = NP == 1 ? pol0 : q;
      Cx r, s;  // interface i -> i+1: reflection coefficient and 1 / transmission coefficient
      if (pol == 0) {
        const Cx ff = nf * cf;
        r = (ii - ff) / (ii + ff);
        s = (ii + ff) / (2. * ii);
      } else {
        const Cx fi = nf * ci, i_f = ni * cf;
        r = (fi - i_f) / (fi + i_f);
        s = (fi + i_f) / (2. * ii);
      }
      const Cx w0 = v0[q] + r * v1[q], w1 = r * v0[q] + v1[q];
      v0[q] = (s * em) * w0;
      v1[q] = (s * ep) * w1;
    }
    nf = ni;
    cf = ci;
  }
  const Cx c0 = th0.im == 0 ? cx(cos(th0.re), 0) : ccos_(th0);
#pragma unroll
  for (int q = 0; q < NP; q++) {
    const int pol = NP == 1 ? pol0 : q;
    const Cx r = v1[q] / v0[q], t = cx(1, 0) / v0[q];
    R[q] = cabs2(r);
    const double tt = cabs2(t);  // |t*t| = |t|^2
    if (pol == 0) T[q] = tt * ((n_last * c_last).re / (n0 * c0).re);
    else T[q] = tt * ((n_last * cconj(c_last)).re / (n0 * cconj(c0)).re);
  }
}
RB_HD inline void tmm_coherent_sub(const DScene& sc, int first, int a, int b, bool reverse, int pol, Cx th0, double lam, double& R, double& T) {
  tmm_coherent_multi<1>(sc, first, a, b, reverse, pol, th0, lam, &R, &T);
}
RB_HD inline void tmm_coherent(const DScene& sc, int ml, int pol, double th0, double lam, double& R, double& T) {
  const rbg_multilayer M = sc.multilayers[ml];
  tmm_coherent_sub(sc, M.first, 0, M.n - 1, false, pol, cx(th0, 0), lam, R, T);
}
// power reflectance / transmittance of a single interface (tmm.interface_R / interface_T)
RB_HD inline void tmm_interface_RT(int pol, Cx ni, Cx nf, Cx thi, Cx thf, double& R, double& T) {
  Cx ci = ccos_(thi), cf = ccos_(thf), r, t;
  if (pol == 0) {
    r = (ni * ci - nf * cf) / (ni * ci + nf * cf);
    t = (2. * (ni * ci)) / (ni * ci + nf * cf);
    T = cabs2(t) * ((nf * cf).re / (ni * ci).re);
  } else {
    r = (nf * ci - ni * cf) / (nf * ci + ni * cf);
    t = (2. * (ni * ci)) / (nf * ci + ni * cf);
    T = cabs2(t) * ((nf * cconj(cf)).re / (ni * cconj(ci)).re);
  }
  R = cabs2(r);
}
// AMultilayer::IncoherentTMM (src/AMultilayer.cxx:484-731, tmm.inc_tmm): layers flagged incoherent (and the two
// semi-infinite ends) exchange power, runs of coherent layers between them are condensed by the coherent TMM in both
// directions; the 2x2 power transfer matrices are multiplied on the fly from the top incoherent layer down.
RB_HD inline void tmm_incoherent(const DScene& sc, int ml, int pol, Cx th0, double lam, double& R, double& T) {
  const rbg_multilayer M = sc.multilayers[ml];
  const int N = M.n;
  const rbg_layer L0 = sc.layers[M.first];
  const Cx n0 = cx(index_n(sc, L0.index, lam), index_k(sc, L0.index, lam));
  const Cx n0s = n0 * csin_(th0);
  auto n_of = [&](int i) {
    const rbg_layer L = sc.layers[M.first + i];
    return cx(index_n(sc, L.index, lam), index_k(sc, L.index, lam));
  };
  auto th_of = [&](int i) {  // ListSnell: forward-angle correction for the first and the last layer only
    Cx ni = n_of(i), th = casin_(n0s / ni);
    if ((i == 0 || i == N - 1) && !tmm_is_forward(ni, th)) th = cx(RB_PI - th.re, -th.im);
    return th;
  };
  double l00 = 1, l01 = 0, l10 = 0, l11 = 1;  // Ltilde
  int prev = 0;                                // previous incoherent layer (index into the full list)
  bool first_pair = true;
  for (int i = 1; i < N; i++) {
    const bool inc = i == N - 1 || sc.layers[M.first + i].incoherent != 0;
    if (!inc) continue;
    double Rf, Tf, Rb, Tb;  // prev -> i and i -> prev
    if (i == prev + 1) {
      tmm_interface_RT(pol, n_of(prev), n_of(i), th_of(prev), th_of(i), Rf, Tf);
      tmm_interface_RT(pol, n_of(i), n_of(prev), th_of(i), th_of(prev), Rb, Tb);
    } else {
      tmm_coherent_sub(sc, M.first, prev, i, false, pol, th_of(prev), lam, Rf, Tf);
      tmm_coherent_sub(sc, M.first, prev, i, true, pol, th_of(i), lam, Rb, Tb);
    }
    double a00 = 1, a01 = -Rb, a10 = Rf, a11 = Tb * Tf - Rb * Rf;
    if (first_pair) {
      l00 = a00 / Tf; l01 = a01 / Tf; l10 = a10 / Tf; l11 = a11 / Tf;
      first_pair = false;
    } else {
      // L = diag(1/P, P) * [[1,-Rb],[Rf, Tb Tf - Rb Rf]] / Tf, P = single-pass transmission of incoherent layer `prev`
      Cx np = n_of(prev);
      double P = exp(-4 * RB_PI * sc.layers[M.first + prev].thickness * (np * ccos_(th_of(prev))).im / lam);
      if (P < 1e-30) P = 1e-30;
      double b00 = a00 / P / Tf, b01 = a01 / P / Tf, b10 = a10 * P / Tf, b11 = a11 * P / Tf;
      double q00 = l00 * b00 + l01 * b10, q01 = l00 * b01 + l01 * b11, q10 = l10 * b00 + l11 * b10, q11 = l10 * b01 + l11 * b11;
      l00 = q00; l01 = q01; l10 = q10; l11 = q11;
    }
    prev = i;
  }
  T = 1 / l00;
  R = l10 / l00;
}
RB_HD inline RB_NOINLINE void tmm_mixed(const DScene& sc, int ml, double th, double lam, double& R, double& T) {  // one copy: two call sites
  const rbg_multilayer M = sc.multilayers[ml];
  if (M.table_r >= 0 && M.table_t >= 0) {
    R = th2_interp(sc, M.table_r, lam, th);
    T = th2_interp(sc, M.table_t, lam, th);
    return;
  }
  double r2[2], t2[2];  // 0 = s, 1 = p
  tmm_coherent_multi<2>(sc, M.first, 0, M.n - 1, false, 0, cx(th, 0), lam, r2, t2);
  const double rp = r2[1], tp = t2[1], rs = r2[0], ts = t2[0];
  R = (rp + rs) / 2.;
  T = (tp + ts) / 2.;
}

// ================================================================== primitive shapes
// Crossing parameters of all bounding surfaces of a non-convex primitive; the intervals between consecutive
// candidates are classified by Contains().  The candidates stay in registers (fixed slots, RB_BIG = unused) and are
// visited in ascending order by repeated selection of the smallest one above the last — no sorted array in local memory.
RB_HD inline double cand_ok(double t) { return (t > 1e-11 && t <= 1e29) ? t : RB_BIG; }
RB_HD inline void cand_quadratic(double A, double B, double C, double& t0, double& t1) {
  t0 = t1 = RB_BIG;
  if (fabs(A) < 1e-300 || fabs(A) < 1e-14 * fabs(B)) {
    if (B != 0) t0 = cand_ok(-C / B);
    return;
  }
  double disc = B * B - 4 * A * C;
  if (disc < 0) return;
  double s = sqrt(disc), q = -0.5 * (B + (B >= 0 ? s : -s));
  t0 = cand_ok(q / A);
  if (q != 0) t1 = cand_ok(C / q);
}

// ---- TGeoBBox  P: dx,dy,dz,ox,oy,oz
RB_HD inline bool bbox_contains(const double* P, V3 p) { return !(fabs(p.x - P[3]) > P[0] || fabs(p.y - P[4]) > P[1] || fabs(p.z - P[5]) > P[2]); }
RB_HD inline double bbox_dist_in(const double* P, V3 p, V3 d) {
  double np[3] = {p.x - P[3], p.y - P[4], p.z - P[5]}, dd[3] = {d.x, d.y, d.z}, smin = RB_BIG;
#pragma unroll
  for (int i = 0; i < 3; i++)
    if (dd[i] != 0) {
      double s = dd[i] > 0 ? (P[i] - np[i]) / dd[i] : -(P[i] + np[i]) / dd[i];
      if (s < 0) return 0.0;
      if (s < smin) smin = s;
    }
  return smin;
}
RB_HD inline double bbox_dist_out(const double* P, V3 p, V3 d, double step) {
  double np[3] = {p.x - P[3], p.y - P[4], p.z - P[5]}, dd[3] = {d.x, d.y, d.z}, saf[3];
  bool in = true;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    saf[i] = fabs(np[i]) - P[i];
    if (saf[i] >= step) return RB_BIG;
    if (in && saf[i] > 0) in = false;
  }
  if (in) {
    int j = 0;
    double ss = saf[0];
    if (saf[1] > ss) { ss = saf[1]; j = 1; }
    if (saf[2] > ss) j = 2;
    if (np[j] * dd[j] > 0) return RB_BIG;
    return 0.0;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (saf[i] < 0) continue;
    if (np[i] * dd[i] >= 0) continue;
    double snxt = saf[i] / fabs(dd[i]);
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 3; j++)
      if (j != i && fabs(np[j] + snxt * dd[j]) > P[j]) ok = false;
    if (ok) return snxt;
  }
  return RB_BIG;
}
RB_HD inline V3 bbox_normal(const double* P, V3 p, V3 d) {
  double s0 = fabs(P[0] - fabs(p.x - P[3])), s1 = fabs(P[1] - fabs(p.y - P[4])), s2 = fabs(P[2] - fabs(p.z - P[5]));
  int i = s1 < s0 ? 1 : 0;
  if (s2 < (i ? s1 : s0)) i = 2;
  if (i == 0) return v3(d.x > 0 ? 1 : -1, 0, 0);
  if (i == 1) return v3(0, d.y > 0 ? 1 : -1, 0);
  return v3(0, 0, d.z > 0 ? 1 : -1);
}

// ---- TGeoTube  P: rmin,rmax,dz
RB_HD inline void tube_roots(double rsq, double nsq, double rdotn, double radius, double& b, double& delta) {
  double inv = 1. / nsq;
  b = inv * rdotn;
  double c = inv * (rsq - radius * radius);
  delta = b * b - c;
  delta = delta > 0 ? sqrt(delta) : -1;
}
RB_HD inline bool tube_contains(const double* P, V3 p) {
  if (fabs(p.z) > P[2]) return false;
  double r2 = p.x * p.x + p.y * p.y;
  return !(r2 < P[0] * P[0] || r2 > P[1] * P[1]);
}
RB_HD inline double tube_dist_in(double rmin, double rmax, double dz, V3 p, V3 d) {
  double sz = RB_BIG;
  if (d.z != 0) {
    sz = ((d.z >= 0 ? dz : -dz) - p.z) / d.z;
    if (sz <= 0) return 0.0;
  }
  double nsq = d.x * d.x + d.y * d.y;
  if (fabs(nsq) < RB_TOL) return sz;
  double rsq = p.x * p.x + p.y * p.y, rdotn = p.x * d.x + p.y * d.y, b, dl;
  if (rmin > 0) {
    if (rsq <= rmin * rmin + RB_TOL) {
      if (rdotn < 0) return 0.0;
    } else if (rdotn < 0) {
      tube_roots(rsq, nsq, rdotn, rmin, b, dl);
      if (dl > 0) {
        double sr = -b - dl;
        if (sr > 0) return rb_min(sz, sr);
      }
    }
  }
  if (rsq >= rmax * rmax - RB_TOL && rdotn >= 0) return 0.0;
  tube_roots(rsq, nsq, rdotn, rmax, b, dl);
  if (dl > 0) {
    double sr = -b + dl;
    if (sr > 0) return rb_min(sz, sr);
  }
  return 0.;
}
RB_HD inline double tube_dist_out(double rmin, double rmax, double dz, V3 p, V3 d) {
  double rmaxsq = rmax * rmax, rminsq = rmin * rmin, zi = dz - fabs(p.z);
  bool inz = !(zi < 0);
  if (!inz) {
    if (p.z * d.z >= 0) return RB_BIG;
    double s = -zi / fabs(d.z), xi = p.x + s * d.x, yi = p.y + s * d.y, r2 = xi * xi + yi * yi;
    if (rminsq <= r2 && r2 <= rmaxsq) return s;
  }
  double rsq = p.x * p.x + p.y * p.y, nsq = d.x * d.x + d.y * d.y, rdotn = p.x * d.x + p.y * d.y, b, dl;
  bool inrmax = rsq <= rmaxsq + RB_TOL, inrmin = rsq >= rminsq - RB_TOL;
  if (inz && inrmin && inrmax) {  // on a boundary within machine precision
    double r = sqrt(rsq);
    if (zi < rmax - r && (fabs(rmin) < RB_TOL || zi < r - rmin)) return p.z * d.z < 0 ? 0.0 : RB_BIG;
    if ((rmaxsq - rsq) < (rsq - rminsq)) return rdotn >= 0 ? RB_BIG : 0.0;
    if (fabs(rmin) < RB_TOL) return 0.0;
    if (rdotn >= 0) return 0.0;
    if (fabs(nsq) < RB_TOL) return RB_BIG;
    tube_roots(rsq, nsq, rdotn, rmin, b, dl);
    if (dl > 0) {
      double s = -b + dl;
      if (s > 0 && fabs(p.z + s * d.z) <= dz) return s;
    }
    return RB_BIG;
  }
  if (fabs(nsq) < RB_TOL) return RB_BIG;
  if (!inrmax) {
    tube_roots(rsq, nsq, rdotn, rmax, b, dl);
    if (dl > 0) {
      double s = -b - dl;
      if (s > 0 && fabs(p.z + s * d.z) <= dz) return s;
    }
  }
  if (rmin > 0) {
    tube_roots(rsq, nsq, rdotn, rmin, b, dl);
    if (dl > 0) {
      double s = -b + dl;
      if (s > 0 && fabs(p.z + s * d.z) <= dz) return s;
    }
  }
  return RB_BIG;
}
RB_HD inline V3 tube_normal(const double* P, V3 p, V3 d) {
  double r = sqrt(p.x * p.x + p.y * p.y);
  double s0 = fabs(P[2] - fabs(p.z)), s1 = P[0] > 1e-10 ? fabs(r - P[0]) : RB_BIG, s2 = fabs(P[1] - r);
  if (s0 <= s1 && s0 <= s2) return v3(0, 0, d.z >= 0 ? 1 : -1);
  double nx = r > 0 ? p.x / r : 1., ny = r > 0 ? p.y / r : 0.;  // (cos phi, sin phi) of phi = atan2(y, x)
  if (nx * d.x + ny * d.y < 0) { nx = -nx; ny = -ny; }
  return v3(nx, ny, 0);
}

// ---- TGeoParaboloid  P: rlo,rhi,dz,a,b   (z = a r^2 + b)
RB_HD inline bool para_contains(const double* P, V3 p) {
  if (fabs(p.z) > P[2]) return false;
  double aa = P[3] * (p.z - P[4]);
  if (aa < 0) return false;
  return !(aa < P[3] * P[3] * (p.x * p.x + p.y * p.y));
}
RB_HD inline double para_surface(const double* P, V3 p, V3 d, bool in) {
  double rsq = p.x * p.x + p.y * p.y, fa = P[3];
  double a = fa * (d.x * d.x + d.y * d.y), b = 2. * fa * (p.x * d.x + p.y * d.y) - d.z, c = fa * rsq + P[4] - p.z;
  if (fabs(a) < RB_TOL) {
    if (fabs(b) < RB_TOL) return RB_BIG;
    double dist = -c / b;
    return dist < 0 ? RB_BIG : dist;
  }
  double ainv = 1. / a, sum = -b * ainv, prod = c * ainv, delta = sum * sum - 4. * prod;
  if (delta < 0) return RB_BIG;
  delta = sqrt(delta);
  double sone = ainv >= 0 ? 1. : -1.;
  for (int i = -1; i < 2; i += 2) {
    double dist = 0.5 * (sum + i * sone * delta);
    if (dist < 0) continue;
    if (dist < 1.E-8) {
      double rr = sqrt(rsq), talf = -2. * fa * rr;
      double ndotd = (rr > 0 ? talf * (p.x * d.x + p.y * d.y) / rr : talf * d.x) + d.z;
      if (!in) ndotd = -ndotd;
      if (ndotd < 0) return dist;
    } else return dist;
  }
  return RB_BIG;
}
RB_HD inline double para_dist_in(const double* P, V3 p, V3 d) {
  double dz = RB_BIG;
  if (d.z < 0) dz = -(p.z + P[2]) / d.z;
  else if (d.z > 0) dz = (P[2] - p.z) / d.z;
  return rb_min(dz, para_surface(P, p, d, true));
}
RB_HD inline double para_dist_out(const double* P, V3 p, V3 d) {
  if (p.z <= -P[2]) {
    if (d.z <= 0) return RB_BIG;
    double s = -(P[2] + p.z) / d.z, xn = p.x + s * d.x, yn = p.y + s * d.y;
    if (xn * xn + yn * yn <= P[0] * P[0]) return s;
  } else if (p.z >= P[2]) {
    if (d.z >= 0) return RB_BIG;
    double s = (P[2] - p.z) / d.z, xn = p.x + s * d.x, yn = p.y + s * d.y;
    if (xn * xn + yn * yn <= P[1] * P[1]) return s;
  }
  double s = para_surface(P, p, d, false);
  if (s > 1E20) return s;
  return fabs(p.z + s * d.z) <= P[2] ? s : RB_BIG;
}
RB_HD inline V3 para_normal(const double* P, V3 p, V3 d) {
  if ((fabs(p.z) - P[2]) > -1E-5) return v3(0, 0, d.z >= 0 ? 1. : -1.);
  double safz = P[2] - fabs(p.z), r = sqrt(p.x * p.x + p.y * p.y), safr = fabs(r - sqrt((p.z - P[4]) / P[3]));
  if (safz < safr) return v3(0, 0, d.z >= 0 ? 1. : -1.);
  double talf = -2. * P[3] * r, calf = 1. / sqrt(1. + talf * talf), salf = talf * calf;
  double cph = r > 0 ? p.x / r : 1., sph = r > 0 ? p.y / r : 0.;
  V3 n = v3(salf * cph, salf * sph, calf);
  if (dot(n, d) < 0) n = v3(-n.x, -n.y, -n.z);
  return n;
}

// ---- TGeoSphere  P: rmin,rmax,th1,th2,ph1,ph2(deg), c1,s1,c2,s2, cp1,sp1,cp2,sp2, flags
static RB_HD RB_NOINLINE bool sphere_phi_outside(const double* P, V3 p) {  // azimuthal segment (rare): kept out of line, atan2 is large
  double phi = rb_atan2(p.y, p.x) * 180. / RB_PI;
  while (phi < P[4]) phi += 360.;
  return phi - P[4] > P[5] - P[4];
}
RB_HD inline bool sphere_contains(const double* P, V3 p) {
  double r2 = dot(p, p);
  if (P[0] > 0 && r2 < P[0] * P[0]) return false;
  if (r2 > P[1] * P[1]) return false;
  if (r2 < 1E-20) return true;
  int flags = (int)P[14];
  if ((flags & 4) && sphere_phi_outside(P, p)) return false;
  if (flags & 3) {
    double ct = p.z / sqrt(r2);  // theta >= th1 <=> cos(theta) <= cos(th1)
    if ((flags & 1) && ct > P[6]) return false;
    if ((flags & 2) && ct < P[8]) return false;
  }
  return true;
}
RB_HD inline double sphere_dist(const double* P, V3 p, V3 d, bool from_inside, double limit = RB_BIG) {
  double c[10];
#pragma unroll
  for (int k = 0; k < 10; k++) c[k] = RB_BIG;
  double a = dot(d, d), b = 2 * dot(p, d), pp = dot(p, p);
  if (!from_inside && P[0] > 0 && pp < P[0] * P[0] && ((int)P[14] & 4) == 0) {
    // Start point in the hollow of a shell (every ray that meets a mirror facet from its concave side, and every ray leaving
    // one): the solid lies between the two spheres, so it can only be entered where the ray leaves the inner sphere, at E, or —
    // if E is outside the polar range — through a cone while the ray crosses the shell between E and its exit from the outer
    // sphere at X; after X it never comes back.  One quadratic settles the first case, two more the second; same root
    // formulas as the general candidate scan below, which takes whatever is left.
    const int flags = (int)P[14];
    double e0, e1;
    cand_quadratic(a, b, pp - P[0] * P[0], e0, e1);
    const double tE = rb_min(e0, e1);
    if (tE > limit && tE < 1e29) return RB_BIG;  // the shell starts beyond the caller's limit (the step already found)
    if (tE < 1e29) {
      const double ct = (p.z + tE * d.z) / P[0];
      if (!(((flags & 1) && ct > P[6]) || ((flags & 2) && ct < P[8]))) {
        // E is in the polar range with a margin? (on the edge of the range the interval classification of the scan decides)
        const double m = 1e-9;
        if (!(((flags & 1) && ct > P[6] - m) || ((flags & 2) && ct < P[8] + m))) return tE;
      } else {
        double x0, x1;
        cand_quadratic(a, b, pp - P[1] * P[1], x0, x1);
        const double tX = rb_min(x0, x1);
        bool cone_in_window = false;
        const double dxy = d.x * d.x + d.y * d.y, pdxy = p.x * d.x + p.y * d.y, pxy = p.x * p.x + p.y * p.y;
#pragma unroll
        for (int k = 0; k < 2; k++) {
          if (!(flags & (1 << k))) continue;
          const double c2 = P[6 + 2 * k] * P[6 + 2 * k], s2 = P[7 + 2 * k] * P[7 + 2 * k];
          double t0, t1;
          cand_quadratic(dxy * c2 - d.z * d.z * s2, 2 * (pdxy * c2 - p.z * d.z * s2), pxy * c2 - p.z * p.z * s2, t0, t1);
          const double lo = tE - 1e-9, hi = tX + 1e-9;
          cone_in_window = cone_in_window || (t0 >= lo && t0 <= hi) || (t1 >= lo && t1 <= hi);
        }
        if (!cone_in_window && tX < 1e29) return RB_BIG;
      }
    }
  }
  if (P[0] > 0) cand_quadratic(a, b, pp - P[0] * P[0], c[0], c[1]);
  cand_quadratic(a, b, pp - P[1] * P[1], c[2], c[3]);
  int flags = (int)P[14];
  double dxy = d.x * d.x + d.y * d.y, pdxy = p.x * d.x + p.y * d.y, pxy = p.x * p.x + p.y * p.y;
  if (flags & 1) {
    double c2 = P[6] * P[6], s2 = P[7] * P[7];
    cand_quadratic(dxy * c2 - d.z * d.z * s2, 2 * (pdxy * c2 - p.z * d.z * s2), pxy * c2 - p.z * p.z * s2, c[4], c[5]);
  }
  if (flags & 2) {
    double c2 = P[8] * P[8], s2 = P[9] * P[9];
    cand_quadratic(dxy * c2 - d.z * d.z * s2, 2 * (pdxy * c2 - p.z * d.z * s2), pxy * c2 - p.z * p.z * s2, c[6], c[7]);
  }
  if (flags & 4) {
    double den = d.y * P[10] - d.x * P[11];
    if (den != 0) c[8] = cand_ok(-(p.y * P[10] - p.x * P[11]) / den);
    den = d.y * P[12] - d.x * P[13];
    if (den != 0) c[9] = cand_ok(-(p.y * P[12] - p.x * P[13]) / den);
  }
  double prev = 0, last = 0;
  for (int it = 0; it < 10; it++) {
    double t = RB_BIG;
#pragma unroll
    for (int k = 0; k < 10; k++) t = (c[k] > last && c[k] < t) ? c[k] : t;
    if (t > 1e29) break;
    last = t;
    if (t - prev < 1e-12) { prev = t; continue; }
    bool in = sphere_contains(P, along(p, d, 0.5 * (prev + t)));
    if (from_inside ? !in : in) return prev;
    prev = t;
  }
  return from_inside ? prev : RB_BIG;
}
RB_HD inline V3 sphere_normal(const double* P, V3 p, V3 d) {
  double r = sqrt(dot(p, p)), rxy = sqrt(p.x * p.x + p.y * p.y);
  int flags = (int)P[14];
  double best = P[0] > 0 ? fabs(r - P[0]) : RB_BIG;
  int which = 0;
  double s = fabs(P[1] - r);
  if (s < best) { best = s; which = 1; }
  // distance to a theta cone: r * |sin(theta - th)|, with sin(theta - th) = (rxy*cos(th) - z*sin(th))/r
  if (flags & 1) { s = fabs(rxy * P[6] - p.z * P[7]); if (s < best) { best = s; which = 2; } }
  if (flags & 2) { s = fabs(rxy * P[8] - p.z * P[9]); if (s < best) { best = s; which = 3; } }
  if (flags & 4) {
    s = fabs(p.y * P[10] - p.x * P[11]); if (s < best) { best = s; which = 4; }
    s = fabs(p.y * P[12] - p.x * P[13]); if (s < best) { best = s; which = 5; }
  }
  V3 n;
  if (which < 2) n = r > 0 ? v3(p.x / r, p.y / r, p.z / r) : v3(0, 0, 1);
  else if (which < 4) {
    double cth = which == 2 ? P[6] : P[8], sth = which == 2 ? P[7] : P[9];
    double cph = rxy > 0 ? p.x / rxy : 1, sph = rxy > 0 ? p.y / rxy : 0;
    n = v3(cth * cph, cth * sph, -sth);
  } else n = which == 4 ? v3(-P[11], P[10], 0) : v3(-P[13], P[12], 0);
  if (dot(n, d) < 0) n = v3(-n.x, -n.y, -n.z);
  return n;
}

// ---- TGeoPgon / TGeoPcon (rmin == 0, full 360 deg; enforced at scene build)
// P: phi1,dphi,nedges,nz, nz x (z,rmin,rmax), nedges x (cos,sin) of the edge-centre azimuths; a polycone is stored
// with nedges = 0 (CONE = true: the radial measure is sqrt(x^2+y^2) instead of the largest edge projection).
// The solid is a stack of convex slabs: each slab is clipped analytically (2 z planes + nedges side planes, or one
// cone frustum); no trigonometry per ray.
template <bool CONE> RB_HD inline double poly_radius(const double* P, V3 p, int& eb) {
  eb = 0;
  if (CONE) return sqrt(p.x * p.x + p.y * p.y);
  int ne = (int)P[2], nz = (int)P[3];
  const double* cs = P + 4 + 3 * nz;
  double m = -RB_BIG;
  for (int e = 0; e < ne; e++) {
    double pr = p.x * cs[2 * e] + p.y * cs[2 * e + 1];
    if (pr > m) { m = pr; eb = e; }
  }
  return m;
}
template <bool CONE> RB_HD inline bool poly_contains(const double* P, V3 p) {
  int nz = (int)P[3];
  const double* sec = P + 4;
  if (p.z < sec[0] || p.z > sec[3 * (nz - 1)]) return false;
  int eb;
  double r = poly_radius<CONE>(P, p, eb);
  int iz = 0, hi = nz;  // last section with z <= p.z (sections ascend; bisection instead of a scan over up to 100 of them)
  while (hi - iz > 1) {
    int mid = (iz + hi) >> 1;
    if (sec[3 * mid] <= p.z) iz = mid; else hi = mid;
  }
  if (iz == nz - 1) return !(r > sec[3 * iz + 2]);
  double dz = sec[3 * (iz + 1)] - sec[3 * iz];
  if (dz < 1E-8) return !(r > rb_max(sec[3 * iz + 2], sec[3 * (iz + 1) + 2]));
  double rmax = sec[3 * iz + 2] + (p.z - sec[3 * iz]) / dz * (sec[3 * (iz + 1) + 2] - sec[3 * iz + 2]);
  return !(r > rmax);
}
// ray interval inside slab k; returns false if empty
template <bool CONE> RB_HD inline bool poly_slab(const double* P, int k, V3 p, V3 d, double& tin, double& tout) {
  int ne = (int)P[2], nz = (int)P[3];
  const double* sec = P + 4;
  const double* cs = P + 4 + 3 * nz;
  double z0 = sec[3 * k], z1 = sec[3 * (k + 1)], r0 = sec[3 * k + 2], r1 = sec[3 * (k + 1) + 2], dz = z1 - z0;
  if (dz < 1E-8) return false;
  tin = -RB_BIG;
  tout = RB_BIG;
  if (d.z != 0) {
    double ta = (z0 - p.z) / d.z, tb = (z1 - p.z) / d.z;
    tin = rb_min(ta, tb);
    tout = rb_max(ta, tb);
  } else if (p.z < z0 || p.z > z1) return false;
  if (tout <= 1e-11) return false;  // the whole slab lies behind the ray: no caller uses such an interval
  double s = rb_div(r1 - r0, dz), base = r0 + (p.z - z0) * s;
  if constexpr (CONE) {
    // inside the frustum: f(t) = |xy(t)|^2 - (base + s dz t)^2 <= 0 on the nappe with non-negative radius, which is
    // the only one the z slab can reach (r0, r1 >= 0)
    double g = s * d.z, A = d.x * d.x + d.y * d.y - g * g, B = 2 * (p.x * d.x + p.y * d.y - base * g), C = p.x * p.x + p.y * p.y - base * base;
    if (fabs(A) < 1e-14 * (d.x * d.x + d.y * d.y + g * g)) {  // parallel to a generator: f is linear
      if (B > 0) tout = rb_min(tout, -C / B);
      else if (B < 0) tin = rb_max(tin, -C / B);
      else if (C > 0) return false;
      return tin < tout;
    }
    double disc = B * B - 4 * A * C;
    if (disc < 0) return false;  // (A < 0 cannot have disc < 0: the line would cross the axis plane inside both nappes)
    double sq = sqrt(disc), q = -0.5 * (B + (B >= 0 ? sq : -sq));
    double ta = q / A, tb = q != 0 ? C / q : ta;
    double t1 = rb_min(ta, tb), t2 = rb_max(ta, tb);
    if (A > 0) { tin = rb_max(tin, t1); tout = rb_min(tout, t2); }
    else if (g > 0) tin = rb_max(tin, t2);
    else tout = rb_min(tout, t1);
    return tin < tout;
  } else {
    for (int e = 0; e < ne; e++) {
      // half-space: x c + y s - (r0 + (z - z0) s) <= 0
      double f0 = p.x * cs[2 * e] + p.y * cs[2 * e + 1] - base, fd = d.x * cs[2 * e] + d.y * cs[2 * e + 1] - s * d.z;
      if (fd > 0) tout = rb_min(tout, -f0 / fd);
      else if (fd < 0) tin = rb_max(tin, -f0 / fd);
      else if (f0 > 0) return false;
      if (!(tin < tout)) return false;  // already empty: the remaining edges can only shrink it further
    }
    return true;
  }
}
template <bool CONE> RB_HD inline double poly_dist_out(const double* P, V3 p, V3 d) {
  int nz = (int)P[3];
  const double* sec = P + 4;
  if (d.z == 0) {  // the ray stays in its z slab(s)
    double best = RB_BIG;
    for (int k = 0; k + 1 < nz; k++) {
      double tin, tout;
      if (!poly_slab<CONE>(P, k, p, d, tin, tout)) continue;
      if (tout <= 1e-11) continue;
      double t = tin > 0 ? tin : 0.0;
      if (tout - t < 1e-12) continue;
      if (t < best) best = t;
    }
    return best;
  }
  // The slabs are stacked in z and z is monotonic along the ray: visited in the order the ray meets them, the first slab with
  // a valid interval holds the smallest entry parameter (a Bezier profile has 100 sections: no need to clip against all).
  const bool up = d.z > 0;
  for (int j = 0; j + 1 < nz; j++) {
    const int k = up ? j : nz - 2 - j;
    if (up ? sec[3 * (k + 1)] <= p.z : sec[3 * k] >= p.z) continue;  // wholly behind the start point
    double tin, tout;
    if (!poly_slab<CONE>(P, k, p, d, tin, tout)) continue;
    if (tout <= 1e-11) continue;
    double t = tin > 0 ? tin : 0.0;
    if (tout - t < 1e-12) continue;
    return t;
  }
  return RB_BIG;
}
// From inside: leave slab after slab until the exit point is not inside a neighbouring slab any more.  The slabs are stacked
// in z, so the slab that continues the path is the next non-degenerate one in the direction of d.z: the march is linear in
// the number of slabs crossed (a Bezier profile has 100 sections).
template <bool CONE> RB_HD inline double poly_dist_in(const double* P, V3 p, V3 d) {
  int nz = (int)P[3];
  const double* sec = P + 4;
  double cur = 0;
  // slabs whose z range holds the start point (two when it sits on a section plane)
  int k = -1;
  for (int j = 0; j + 1 < nz; j++) {
    if (p.z < sec[3 * j] - 1e-9 || p.z > sec[3 * (j + 1)] + 1e-9) continue;
    double tin, tout;
    if (!poly_slab<CONE>(P, j, p, d, tin, tout)) continue;
    if (tin <= cur + 1e-9 && tout > cur + 1e-9) { cur = tout; k = j; }
  }
  if (k < 0) return cur;
  const int dir = d.z > 0 ? 1 : (d.z < 0 ? -1 : 0);
  while (dir != 0) {
    int j = k + dir;
    while (j >= 0 && j + 1 < nz && sec[3 * (j + 1)] - sec[3 * j] < 1E-8) j += dir;  // radius steps have no volume
    if (j < 0 || j + 1 >= nz) break;
    double tin, tout;
    if (!poly_slab<CONE>(P, j, p, d, tin, tout)) break;
    if (!(tin <= cur + 1e-9 && tout > cur + 1e-9)) break;
    cur = tout;
    k = j;
  }
  return cur;
}
template <bool CONE> RB_HD inline V3 poly_normal(const double* P, V3 p, V3 d) {
  int nz = (int)P[3];
  const double* sec = P + 4;
  const double* cs = P + 4 + 3 * nz;
  int eb = 0;
  double r = poly_radius<CONE>(P, p, eb);
  double ux, uy;  // outward radial unit vector of the lateral face
  if (CONE) { ux = r > 0 ? p.x / r : 1.; uy = r > 0 ? p.y / r : 0.; }
  else { ux = cs[2 * eb]; uy = cs[2 * eb + 1]; }
  double best = RB_BIG;
  V3 n = v3(0, 0, 1);
  for (int i = 0; i < nz; i++) {
    bool cap = i == 0 || i == nz - 1;
    bool step = (i + 1 < nz && sec[3 * (i + 1)] - sec[3 * i] < 1e-8) || (i > 0 && sec[3 * i] - sec[3 * (i - 1)] < 1e-8);
    if (!cap && !step) continue;
    double s = fabs(p.z - sec[3 * i]);
    if (s < best) { best = s; n = v3(0, 0, 1); }
  }
  for (int k = 0; k + 1 < nz; k++) {
    double z0 = sec[3 * k], z1 = sec[3 * (k + 1)], dz = z1 - z0;
    if (dz < 1e-8 || p.z < z0 - 1e-6 || p.z > z1 + 1e-6) continue;
    double s = (sec[3 * (k + 1) + 2] - sec[3 * k + 2]) / dz, rr = sec[3 * k + 2] + (p.z - z0) * s, nn = sqrt(1 + s * s);
    double dist = fabs(r - rr) / nn;
    if (dist < best) { best = dist; n = v3(ux / nn, uy / nn, -s / nn); }
  }
  if (dot(n, d) < 0) n = v3(-n.x, -n.y, -n.z);
  return n;
}

// ---- general TGeoPgon / TGeoPcon: hollow sections (rmin > 0) and/or an azimuthal range (dphi < 360 deg).
// After the edge table the parameter block carries: general flag, cos/sin(phi1), cos/sin(phi1 + dphi).
// Not convex any more, so the solid is handled like TGeoSphere: every crossing parameter of every bounding surface
// (z planes, outer and inner lateral faces of each z slab, the two phi planes) is a candidate; the candidates are
// visited in ascending order (selection of the smallest one above the last: no candidate array) and the interval
// between two consecutive candidates is classified by Contains() at its midpoint.
RB_HD inline const double* polyg_tail(const double* P) { return P + 4 + 3 * (int)P[3] + 2 * (int)P[2]; }
RB_HD inline bool poly_is_general(const double* P) { return polyg_tail(P)[0] != 0.; }
template <bool CONE> RB_HD RB_NOINLINE bool polyg_contains(const double* P, V3 p) {
  int ne = (int)P[2], nz = (int)P[3];
  const double* sec = P + 4;
  const double* cs = P + 4 + 3 * nz;
  if (p.z < sec[0] || p.z > sec[3 * (nz - 1)]) return false;
  double r;
  const bool seg = fabs(P[1] - 360.) > 1e-9;
  if (!CONE) {
    double divphi = P[1] / ne, phi = rb_atan2(p.y, p.x) * 180. / RB_PI;
    while (phi < P[0]) phi += 360.0;
    double ddp = phi - P[0];
    if (ddp > P[1]) return false;
    int ipsec = (int)(ddp / divphi);
    if (ipsec > ne - 1) ipsec = ne - 1;
    r = p.x * cs[2 * ipsec] + p.y * cs[2 * ipsec + 1];
  } else {
    r = sqrt(p.x * p.x + p.y * p.y);
    if (seg) {
      double phi = rb_atan2(p.y, p.x) * 180. / RB_PI;
      while (phi < P[0]) phi += 360.0;
      if (phi - P[0] > P[1]) return false;
    }
  }
  int iz = 0;
  for (int i = 0; i < nz; i++)
    if (sec[3 * i] <= p.z) iz = i;
  if (iz == nz - 1) return !(r < sec[3 * iz + 1] || r > sec[3 * iz + 2]);
  double dz = sec[3 * (iz + 1)] - sec[3 * iz];
  if (dz < 1E-8) {
    double rmin = rb_min(sec[3 * iz + 1], sec[3 * (iz + 1) + 1]), rmax = rb_max(sec[3 * iz + 2], sec[3 * (iz + 1) + 2]);
    return !(r < rmin || r > rmax);
  }
  double dzrat = (p.z - sec[3 * iz]) / dz;
  double rmin = sec[3 * iz + 1] + dzrat * (sec[3 * (iz + 1) + 1] - sec[3 * iz + 1]);
  if (r < rmin) return false;
  double rmax = sec[3 * iz + 2] + dzrat * (sec[3 * (iz + 1) + 2] - sec[3 * iz + 2]);
  return !(r > rmax);
}
// smallest valid candidate strictly above `last` (RB_BIG if none)
template <bool CONE> RB_HD inline double polyg_next(const double* P, V3 p, V3 d, double last) {
  int ne = (int)P[2], nz = (int)P[3];
  const double* sec = P + 4;
  const double* cs = P + 4 + 3 * nz;
  const double* tail = polyg_tail(P);
  double best = RB_BIG;
  auto offer = [&](double t) {
    if (t > 1e-11 && t <= 1e29 && t > last && t < best) best = t;
  };
  if (d.z != 0)
    for (int i = 0; i < nz; i++) offer((sec[3 * i] - p.z) / d.z);
  const double dxy = d.x * d.x + d.y * d.y, pdxy = p.x * d.x + p.y * d.y, pxy = p.x * p.x + p.y * p.y;
  for (int k = 0; k + 1 < nz; k++) {
    double z0 = sec[3 * k], z1 = sec[3 * (k + 1)], dz = z1 - z0;
    if (dz < 1E-8) continue;
    for (int w = 0; w < 2; w++) {
      double r0 = sec[3 * k + 1 + w], r1 = sec[3 * (k + 1) + 1 + w];
      if (!w && r0 <= 0 && r1 <= 0) continue;
      double s = (r1 - r0) / dz;
      auto offer_in_slab = [&](double t) {  // keep only crossings inside this z slab
        double zz = p.z + t * d.z;
        if (zz >= z0 - 1e-9 && zz <= z1 + 1e-9) offer(t);
      };
      if (!CONE) {
        for (int e = 0; e < ne; e++) {
          double ux = cs[2 * e], uy = cs[2 * e + 1], den = d.x * ux + d.y * uy - s * d.z;
          if (den != 0) offer_in_slab((r0 + (p.z - z0) * s - (p.x * ux + p.y * uy)) / den);
        }
      } else {
        double a0 = r0 + (p.z - z0) * s, b0 = s * d.z, t0, t1;
        cand_quadratic(dxy - b0 * b0, 2 * (pdxy - a0 * b0), pxy - a0 * a0, t0, t1);
        offer_in_slab(t0);
        offer_in_slab(t1);
      }
    }
  }
  if (fabs(P[1] - 360.) > 1e-9)
    for (int k = 0; k < 2; k++) {
      double co = tail[1 + 2 * k], si = tail[2 + 2 * k], den = d.y * co - d.x * si;
      if (den != 0) offer(-(p.y * co - p.x * si) / den);
    }
  return best;
}
template <bool CONE> RB_HD RB_NOINLINE double polyg_dist(const double* P, V3 p, V3 d, bool from_inside) {
  double prev = 0, last = 0;
  for (int it = 0; it < 4096; it++) {
    double t = polyg_next<CONE>(P, p, d, last);
    if (t > 1e29) break;
    last = t;
    if (t - prev < 1e-12) { prev = t; continue; }
    bool in = polyg_contains<CONE>(P, along(p, d, 0.5 * (prev + t)));
    if (from_inside ? !in : in) return prev;
    prev = t;
  }
  return from_inside ? prev : RB_BIG;
}
template <bool CONE> RB_HD RB_NOINLINE V3 polyg_normal(const double* P, V3 p, V3 d) {
  int ne = (int)P[2], nz = (int)P[3];
  const double* sec = P + 4;
  const double* cs = P + 4 + 3 * nz;
  const double* tail = polyg_tail(P);
  double best = RB_BIG, ux = 0, uy = 0, r;
  V3 n = v3(0, 0, 1);
  if (!CONE) {
    double divphi = P[1] / ne, phi = rb_atan2(p.y, p.x) * 180. / RB_PI;
    while (phi < P[0]) phi += 360.0;
    int ipsec = (int)((phi - P[0]) / divphi);
    ipsec = ipsec < 0 ? 0 : (ipsec > ne - 1 ? ne - 1 : ipsec);
    ux = cs[2 * ipsec]; uy = cs[2 * ipsec + 1];
    r = p.x * ux + p.y * uy;
  } else {
    r = sqrt(p.x * p.x + p.y * p.y);
    ux = r > 0 ? p.x / r : 1; uy = r > 0 ? p.y / r : 0;
  }
  for (int i = 0; i < nz; i++) {
    bool cap = i == 0 || i == nz - 1;
    bool step = (i + 1 < nz && sec[3 * (i + 1)] - sec[3 * i] < 1e-8) || (i > 0 && sec[3 * i] - sec[3 * (i - 1)] < 1e-8);
    if (!cap && !step) continue;
    double s = fabs(p.z - sec[3 * i]);
    if (s < best) { best = s; n = v3(0, 0, 1); }
  }
  for (int k = 0; k + 1 < nz; k++) {
    double z0 = sec[3 * k], z1 = sec[3 * (k + 1)], dz = z1 - z0;
    if (dz < 1e-8 || p.z < z0 - 1e-6 || p.z > z1 + 1e-6) continue;
    for (int w = 0; w < 2; w++) {
      double r0 = sec[3 * k + 1 + w], r1 = sec[3 * (k + 1) + 1 + w];
      if (!w && r0 <= 0 && r1 <= 0) continue;
      double s = (r1 - r0) / dz, rr = r0 + (p.z - z0) * s, nn = sqrt(1 + s * s), dist = fabs(r - rr) / nn;
      if (dist < best) { best = dist; n = v3(ux / nn, uy / nn, -s / nn); }
    }
  }
  if (fabs(P[1] - 360.) > 1e-9)
    for (int k = 0; k < 2; k++) {  // the two phi planes (only the half plane on the solid's side)
      double co = tail[1 + 2 * k], si = tail[2 + 2 * k];
      if (p.x * co + p.y * si < 0) continue;
      double dist = fabs(p.y * co - p.x * si);
      if (dist < best) { best = dist; n = v3(-si, co, 0); }
    }
  if (dot(n, d) < 0) n = v3(-n.x, -n.y, -n.z);
  return n;
}

// ---- AGeoAsphericDisk  P: z1,z2,c1,c2,k1,k2,rmin,rmax,n1,n2,oz,dz,K1[],K2[]
RB_HD inline bool asph_F(const double* P, int s, double r, double& out) {
  double c = P[1 + s], kap = P[3 + s], z0 = P[s - 1];
  int n = (int)P[7 + s];
  const double* K = s == 1 ? P + 12 : P + 12 + (int)P[8];
  double r2 = r * r, pp = r2 * c * c * kap;
  if (1 - pp < 0) return false;
  double poly = 0;
  for (int i = n - 1; i >= 0; i--) poly = (poly + K[i]) * r2;  // sum K_i r^(2(i+1)), Horner in r^2
  out = z0 + rb_div(r2 * c, 1 + sqrt(1 - pp)) + poly;
  return true;
}
RB_HD inline bool asph_dF(const double* P, int s, double r, double& out) {
  double c = P[1 + s], kap = P[3 + s];
  int n = (int)P[7 + s];
  const double* K = s == 1 ? P + 12 : P + 12 + (int)P[8];
  double r2 = r * r, pp = r2 * c * c * kap;
  if (1 - pp <= 0) return false;
  double poly = 0;
  for (int i = n - 1; i >= 0; i--) poly = poly * r2 + 2 * (i + 1) * K[i];  // sum 2(i+1) K_i r^(2i), then * r
  out = rb_div(r * c, sqrt(1 - pp)) + poly * r;
  return true;
}
RB_HD inline bool asph_contains(const double* P, V3 p) {
  double r = sqrt(p.x * p.x + p.y * p.y);
  if (r > P[7] || r < P[6]) return false;
  double f1, f2;
  if (!asph_F(P, 1, r, f1) || !asph_F(P, 2, r, f2)) return false;
  return !(p.z < f1 || f2 < p.z);
}
// ray / even-asphere intersection: conic closed-form start, then Newton on the sag (cap 100, |e|<1e-10)
RB_HD inline double asph_surface(const double* P, int s, V3 pt, V3 dir) {
  double d = P[s - 1], curve = P[1 + s], kappa = P[3 + s];
  int npol = (int)P[7 + s];
  const double* K = s == 1 ? P + 12 : P + 12 + (int)P[8];
  double H2 = pt.x * pt.x + pt.y * pt.y, zr = pt.z - d;
  double p = -(zr * dir.z + pt.x * dir.x + pt.y * dir.y);
  double M = p * dir.z + zr, M2 = zr * zr + H2 - p * p;
  double w = (M2 * curve - 2 * M);
  double check = 1 - rb_div(rb_div(w * curve, dir.z), dir.z);
  if (check < 0) return RB_BIG;
  double q = p + rb_div(w, dir.z * (1 + sqrt(check)));
  double nx = pt.x + q * dir.x, ny = pt.y + q * dir.y, nz = pt.z + q * dir.z - d;
  double ck = curve * kappa, cc = kappa * curve * curve;
  for (int i = 0;; i++) {
    if (i > 100) return RB_BIG;
    H2 = nx * nx + ny * ny;
    check = 1 - cc * H2;
    if (check < 0) return RB_BIG;
    double l = sqrt(check), x = 0, v = 0;
    for (int j = npol - 1; j >= 0; j--) {
      x = (x + K[j]) * H2;
      v = v * H2 + 2 * (j + 1) * K[j];
    }
    if (curve != 0) x += (1 - l) / curve / kappa;
    v = ck + l * v;
    double m = -nx * v, n = -ny * v, inv = 1. / sqrt(l * l + m * m + n * n);
    l *= inv; m *= inv; n *= inv;
    check = dir.z * l + dir.x * m + dir.y * n;
    if (check == 0) return RB_BIG;
    double e = rb_div(l * (x - nz), check);
    nx += e * dir.x; ny += e * dir.y; nz += e * dir.z;
    if (fabs(e) < 1e-10) break;
  }
  nz += d;
  double ex = nx - pt.x, ey = ny - pt.y, ez = nz - pt.z;
  if (dir.x * ex + dir.y * ey + dir.z * ez < 0) return RB_BIG;
  double rho = sqrt(nx * nx + ny * ny);
  if (rho < P[6] || rho > P[7]) return RB_BIG;
  return sqrt(ex * ex + ey * ey + ez * ez);
}
RB_HD inline double asph_cylinder(const double* P, double R, V3 pt, V3 dir) {
  double rsq = pt.x * pt.x + pt.y * pt.y, nsq = dir.x * dir.x + dir.y * dir.y;
  if (sqrt(nsq) < RB_TOL) return RB_BIG;
  double rdotn = pt.x * dir.x + pt.y * dir.y, b, delta;
  tube_roots(rsq, nsq, rdotn, R, b, delta);
  if (delta < 0) return RB_BIG;
  double t1 = -b + delta, t2 = -b - delta;
  if (t1 < 0 && t2 < 0) return RB_BIG;
  double zmin, zmax;
  if (!asph_F(P, 1, R, zmin) || !asph_F(P, 2, R, zmax)) return RB_BIG;
  if (t2 > 0) {
    double z1 = t1 * dir.z + pt.z, z2 = t2 * dir.z + pt.z;
    if (z1 < zmin || zmax < z1) t1 = RB_BIG;
    if (z2 < zmin || zmax < z2) t2 = RB_BIG;
    return t1 < t2 ? t1 : t2;
  }
  bool zin = zmin <= pt.z && pt.z <= zmax;
  if (t2 == 0) {
    if (t1 > 0) {
      if (zin) return 0;
      double z1 = t1 * dir.z + pt.z;
      if (zmin <= z1 && z1 <= zmax) return t1;
    } else if (t1 == 0 && zin) return 0;
    return RB_BIG;
  }
  if (t1 > 0) {
    double z1 = t1 * dir.z + pt.z;
    if (zmin <= z1 && z1 <= zmax) return t1;
  } else if (t1 == 0 && zin) return 0;
  return RB_BIG;
}
RB_HD inline double asph_dist4(const double* P, V3 p, V3 d) {
  double m = asph_surface(P, 1, p, d);
  m = rb_min(m, asph_surface(P, 2, p, d));
  if (P[6] > 0) m = rb_min(m, asph_cylinder(P, P[6], p, d));
  return rb_min(m, asph_cylinder(P, P[7], p, d));
}
RB_HD inline double asph_dist_out(const double* P, V3 p, V3 d, double step) {
  if (tube_dist_out(P[6], P[7], P[11], v3(p.x, p.y, p.z - P[10]), d) >= step) return RB_BIG;
  return asph_dist4(P, p, d);
}
RB_HD inline V3 asph_normal(const double* P, V3 p, V3 d) {
  double r = sqrt(p.x * p.x + p.y * p.y);
  double best = P[6] > 0 ? fabs(r - P[6]) : RB_BIG, f, df = 0;
  int which = 0;
  double s = fabs(r - P[7]);
  if (s < best) { best = s; which = 1; }
  double dfs = 0;
  for (int k = 1; k <= 2; k++) {
    double saf = RB_BIG;
    if (asph_F(P, k, r, f) && asph_dF(P, k, r, df)) saf = rb_div(fabs(f - p.z), sqrt(1 + df * df));
    if (saf < best) { best = saf; which = 1 + k; dfs = df; }
  }
  double nx = 0, nz = 0;
  if (which < 2) nx = 1;
  else if (dfs == 0) nz = 1;
  else { double inv = 1. / sqrt(1 + dfs * dfs); nx = dfs * inv; nz = -inv; }
  V3 n = r > 0 ? v3(rb_div(nx * p.x, r), rb_div(nx * p.y, r), nz) : v3(nx, 0, nz);  // RotateZ(atan2(y, x))
  if (dot(n, d) < 0) n = v3(-n.x, -n.y, -n.z);
  return n;
}

// ---- AGeoWinstonCone2D / Poly  P: r1,r2,(dy|npoly),theta,dz,f,cos(theta),sin(theta)
RB_HD inline bool win_R(const double* P, double z, double& out) {
  if (fabs(z) > P[4] + 1e-10) return false;
  double sint = P[7], cost = P[6], f = P[5], t = z + P[4];
  double a0 = t * t * sint * sint - 4. * f * (t * cost + f), a1 = 2. * t * sint * cost + 4. * f * sint, a2 = cost * cost;
  out = (-a1 + sqrt(a1 * a1 - 4. * a0 * a2)) / (2 * a2) - P[1];
  return true;
}
RB_HD inline bool win_dRdZ(const double* P, double z, double& out) {
  if (fabs(z) > P[4] + 1e-10) return false;
  double sint = P[7], cost = P[6], f = P[5], t = z + P[4];
  double a0 = t * t * sint * sint - 4. * f * (t * cost + f), a1 = 2. * t * sint * cost + 4. * f * sint, a2 = cost * cost;
  double da0 = 2 * t * sint * sint - 4 * f * cost, da1 = 2 * sint * cost;
  out = (-da1 + (a1 * da1 - 2 * da0 * a2) / sqrt(a1 * a1 - 4 * a0 * a2)) / (2 * a2);
  return true;
}
// P (device layout): r1,r2,(dy|npoly),theta,dz,f,cos(theta),sin(theta),tan(theta),tan(pi/npoly), npoly x (cos,sin) of the face azimuths
// Intersection with the tilted parabola of the face at azimuth phi (cphi,sphi).  Same algebra as the reference's
// DistToParabola, with the trigonometry removed: tan(atan2(pz,px) - theta) = (pz - px tan(theta)) / (px + pz tan(theta)),
// and the azimuth window |atan2(yc,xc)| <= open/2 becomes |yc| <= xc tan(open/2) (xc >= r2 > 0; for open = pi it is
// always true).  wtan < 0 means "no window".
RB_HD inline double win_parabola(const double* P, V3 pt, V3 dir, double cphi, double sphi, double wtan) {
  double x = cphi * pt.x + sphi * pt.y, y = -sphi * pt.x + cphi * pt.y, z = pt.z;
  double px = cphi * dir.x + sphi * dir.y, py = -sphi * dir.x + cphi * dir.y, pz = dir.z;
  if (px == 0 && pz == 0) return RB_BIG;
  double r1 = P[0], r2 = P[1], DZ = P[4], f = P[5], cost = P[6], sint = P[7], tant = P[8];
  double X = cost * (x + r2) + (z + DZ) * sint, Z = -sint * (x + r2) + (z + DZ) * cost + f;
  double tanA = (pz - px * tant) / (px + pz * tant);
  double tmp = tanA * tanA - (X * tanA - Z) / f;
  if (tmp < 0) return RB_BIG;
  double Xc[2];
  if (DZ * 2 / fabs(tanA) < RB_TOL) { Xc[0] = X; Xc[1] = X; }
  else { double sq = sqrt(tmp); Xc[0] = 2 * f * (tanA + sq); Xc[1] = 2 * f * (tanA - sq); }
  // the transverse coordinate follows the better conditioned of py/pz, py/px (reference :375-394)
  double apx = fabs(px), apy = fabs(py), apz = fabs(pz);
  bool use_z = (apx <= apz && apy <= apz) ? true : ((apy <= apx && apz <= apx) ? false : apx < 1e-5);
  double slope = use_z ? py / pz : py / px;
  double best = RB_BIG;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    double Zc = Xc[k] * Xc[k] / 4. / f;
    double xc = cost * Xc[k] - sint * (Zc - f) - r2, zc = sint * Xc[k] + cost * (Zc - f) - DZ;
    double ddx = xc - x, ddz = zc - z;
    double ddy = (use_z ? ddz : ddx) * slope, yc = y + ddy;
    if (xc < r2 || r1 < xc || zc < -DZ || DZ < zc || ddx * px + ddz * pz < 0) continue;
    if (wtan < 0 || fabs(yc) <= xc * wtan) best = rb_min(best, sqrt(ddx * ddx + ddy * ddy + ddz * ddz));
  }
  return best;
}
// largest projection of (x,y) on the face normals = rho cos(folded azimuth) of the reference's InsidePolygon
RB_HD inline double win_face_proj(const double* P, double x, double y, int& kbest) {
  int n = (int)P[2];
  const double* cs = P + 10;
  double m = -RB_BIG;
  kbest = 0;
  for (int k = 0; k < n; k++) {
    double pr = x * cs[2 * k] + y * cs[2 * k + 1];
    if (pr > m) { m = pr; kbest = k; }
  }
  return m;
}
RB_HD inline bool win_inside_polygon(const double* P, double x, double y, double r) {
  int k;
  return !(win_face_proj(P, x, y, k) > r);
}
RB_HD inline bool win_contains(const double* P, bool poly, V3 p) {
  double r;
  if (poly) {
    if (fabs(p.z) > P[4]) return false;
    if (!win_R(P, p.z, r)) return false;
    return win_inside_polygon(P, p.x, p.y, r);
  }
  if (fabs(p.y) > P[2] || fabs(p.z) > P[4]) return false;
  if (!win_R(P, p.z, r)) return false;
  return !(fabs(p.x) > r);
}
RB_HD inline double win_dist_in(const double* P, bool poly, V3 p, V3 d) {
  double best = RB_BIG;
  if (d.z < 0) best = -(p.z + P[4]) / d.z;
  else if (d.z > 0) best = (P[4] - p.z) / d.z;
  if (poly) {
    int n = (int)P[2];
    const double* cs = P + 10;
    for (int i = 0; i < n; i++) best = rb_min(best, win_parabola(P, p, d, cs[2 * i], cs[2 * i + 1], -1.));
    return best;
  }
  if (d.y < 0) best = rb_min(best, -(p.y + P[2]) / d.y);
  else if (d.y > 0) best = rb_min(best, (P[2] - p.y) / d.y);
  best = rb_min(best, win_parabola(P, p, d, 1., 0., -1.));
  return rb_min(best, win_parabola(P, p, d, -1., 0., -1.));
}
RB_HD inline double win_dist_out(const double* P, bool poly, V3 p, V3 d) {
  double DZ = P[4];
  if (poly) {
    int n = (int)P[2];
    const double* cs = P + 10;
    if (p.z <= -DZ) {
      if (d.z <= 0) return RB_BIG;
      double s = -(DZ + p.z) / d.z;
      if (win_inside_polygon(P, p.x + s * d.x, p.y + s * d.y, P[1])) return s;
    } else if (p.z >= DZ) {
      if (d.z >= 0) return RB_BIG;
      double s = (DZ - p.z) / d.z;
      if (win_inside_polygon(P, p.x + s * d.x, p.y + s * d.y, P[0])) return s;
    }
    double best = RB_BIG;
    for (int i = 0; i < n; i++) best = rb_min(best, win_parabola(P, p, d, cs[2 * i], cs[2 * i + 1], P[9]));
    return best;
  }
  double DY = P[2], r;
  if (p.z <= -DZ) {
    if (d.z <= 0) return RB_BIG;
    double s = -(DZ + p.z) / d.z;
    if (fabs(p.x + s * d.x) <= P[1] && fabs(p.y + s * d.y) <= DY) return s;
  } else if (p.z >= DZ) {
    if (d.z >= 0) return RB_BIG;
    double s = (DZ - p.z) / d.z;
    if (fabs(p.x + s * d.x) <= P[0] && fabs(p.y + s * d.y) <= DY) return s;
  }
  if (p.y <= -DY) {
    if (d.y <= 0) return RB_BIG;
    double s = -(DY + p.y) / d.y, xn = p.x + s * d.x, zn = p.z + s * d.z;
    if (fabs(zn) <= DZ && win_R(P, zn, r) && fabs(xn) <= r) return s;
  } else if (p.y >= DY) {
    if (d.y >= 0) return RB_BIG;
    double s = (DY - p.y) / d.y, xn = p.x + s * d.x, zn = p.z + s * d.z;
    if (fabs(zn) <= DZ && win_R(P, zn, r) && fabs(xn) <= r) return s;
  }
  double s0 = win_parabola(P, p, d, 1., 0., -1.);
  if (!(fabs(p.y + s0 * d.y) <= DY)) s0 = RB_BIG;
  double s1 = win_parabola(P, p, d, -1., 0., -1.);
  if (!(fabs(p.y + s1 * d.y) <= DY)) s1 = RB_BIG;
  return rb_min(s0, s1);
}
RB_HD inline V3 win_normal(const double* P, bool poly, V3 p, V3 d) {
  double r, dr = 0;
  V3 n;
  if (poly) {
    const double* cs = P + 10;
    int k;
    double proj = win_face_proj(P, p.x, p.y, k);
    double s0 = fabs(fabs(P[4]) - fabs(p.z));
    double s1 = win_R(P, p.z, r) ? fabs(r - proj) : RB_BIG;
    if (!(s1 < s0)) n = v3(0, 0, 1);
    else {
      win_dRdZ(P, p.z, dr);
      n = v3(cs[2 * k], cs[2 * k + 1], -dr);
    }
  } else {
    double s0 = fabs(fabs(P[2]) - fabs(p.y)), s1 = fabs(fabs(P[4]) - fabs(p.z)), s2 = win_R(P, p.z, r) ? fabs(r - fabs(p.x)) : RB_BIG;
    if (s0 <= s1 && s0 <= s2) n = v3(0, 1, 0);
    else if (s1 <= s2) n = v3(0, 0, 1);
    else {
      win_dRdZ(P, p.z, dr);
      n = v3(1, 0, p.x > 0 ? -dr : dr);
    }
  }
  double inv = 1. / sqrt(dot(n, n));
  n = inv * n;
  if (dot(n, d) < 0) n = v3(-n.x, -n.y, -n.z);
  return n;
}

// ---- TGeoArb8  P: dz, 8 x (x,y): vertices 0-3 at z = -dz, 4-7 at z = +dz, clockwise seen from +z (TGeoArb8 convention;
// tutorials/AshraOptics.C:264-284,403-441, src/AGeoUtil.cxx:47-82).  Vertices may coincide (triangular / pyramidal solids) and
// a lateral face may be twisted (its lower and upper edges not parallel: a ruled bilinear patch).  At height z the section is
// the quadrilateral of the interpolated vertices; inside means cross((B-A),(P-A)) >= 0 for its four edges (ROOT's
// TGeoArb8::Contains / InsidePolygon).  Along a ray each of these cross products is a quadratic in the ray parameter, so every
// boundary crossing is a root of one of four quadratics or of the two z planes: the solid is evaluated like TGeoSphere
// (register-resident candidate roots, intervals classified by Contains at their midpoints).
RB_HD inline void arb8_edge(const double* P, int i, double s, double& ax, double& ay, double& bx, double& by) {
  const double* v = P + 1;
  int j = (i + 1) & 3;
  ax = v[2 * i] + s * (v[2 * i + 8] - v[2 * i]);
  ay = v[2 * i + 1] + s * (v[2 * i + 9] - v[2 * i + 1]);
  bx = v[2 * j] + s * (v[2 * j + 8] - v[2 * j]);
  by = v[2 * j + 1] + s * (v[2 * j + 9] - v[2 * j + 1]);
}
RB_HD inline bool arb8_contains(const double* P, V3 p) {
  double dz = P[0];
  if (fabs(p.z) > dz) return false;
  double s = 0.5 * (p.z + dz) / dz;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double ax, ay, bx, by;
    arb8_edge(P, i, s, ax, ay, bx, by);
    if ((p.x - ax) * (by - ay) - (p.y - ay) * (bx - ax) < 0) return false;
  }
  return true;
}
RB_HD inline double arb8_dist(const double* P, V3 p, V3 d, bool from_inside) {
  double c[10];
#pragma unroll
  for (int k = 0; k < 10; k++) c[k] = RB_BIG;
  const double dz = P[0];
  const double* v = P + 1;
  if (d.z != 0) {
    c[8] = cand_ok((dz - p.z) / d.z);
    c[9] = cand_ok((-dz - p.z) / d.z);
  }
  const double s0 = 0.5 * (p.z + dz) / dz, s1 = 0.5 * d.z / dz;  // height fraction along the ray: s0 + s1 t
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int j = (i + 1) & 3;
    double eax = v[2 * i + 8] - v[2 * i], eay = v[2 * i + 9] - v[2 * i + 1], ebx = v[2 * j + 8] - v[2 * j], eby = v[2 * j + 9] - v[2 * j + 1];
    double a0x = v[2 * i] + s0 * eax, a0y = v[2 * i + 1] + s0 * eay, b0x = v[2 * j] + s0 * ebx, b0y = v[2 * j + 1] + s0 * eby;
    // U = P - A = u0 + u1 t, W = B - A = w0 + w1 t
    double u0x = p.x - a0x, u0y = p.y - a0y, u1x = d.x - s1 * eax, u1y = d.y - s1 * eay;
    double w0x = b0x - a0x, w0y = b0y - a0y, w1x = s1 * (ebx - eax), w1y = s1 * (eby - eay);
    cand_quadratic(u1x * w1y - u1y * w1x, u0x * w1y + u1x * w0y - u0y * w1x - u1y * w0x, u0x * w0y - u0y * w0x, c[2 * i], c[2 * i + 1]);
  }
  double prev = 0, last = 0;
  for (int it = 0; it < 10; it++) {
    double t = RB_BIG;
#pragma unroll
    for (int k = 0; k < 10; k++) t = (c[k] > last && c[k] < t) ? c[k] : t;
    if (t > 1e29) break;
    last = t;
    if (t - prev < 1e-12) { prev = t; continue; }
    bool in = arb8_contains(P, along(p, d, 0.5 * (prev + t)));
    if (from_inside ? !in : in) return prev;
    prev = t;
  }
  return from_inside ? prev : RB_BIG;
}
// TGeoArb8::ComputeNormal: a z face within 10 tolerances, else the lateral face of the closest edge of the section at the
// point's height; its normal is (edge direction) x (ruling direction at the foot point), exact on a twisted face too.
RB_HD inline V3 arb8_normal(const double* P, V3 p, V3 d) {
  const double dz = P[0];
  const double* v = P + 1;
  if (dz - fabs(p.z) < 10. * RB_TOL) return v3(0, 0, d.z >= 0 ? 1. : -1.);
  double s = 0.5 * (p.z + dz) / dz;
  s = s < 0 ? 0 : (s > 1 ? 1 : s);
  double best = RB_BIG, frac = 0;
  int iseg = 0;
  for (int i = 0; i < 4; i++) {
    double ax, ay, bx, by;
    arb8_edge(P, i, s, ax, ay, bx, by);
    double ex = bx - ax, ey = by - ay, len2 = ex * ex + ey * ey, ux = p.x - ax, uy = p.y - ay, f = 0, dist2;
    if (len2 < 1e-20) dist2 = ux * ux + uy * uy;
    else {
      f = (ux * ex + uy * ey) / len2;
      f = f < 0 ? 0 : (f > 1 ? 1 : f);
      double qx = ux - f * ex, qy = uy - f * ey;
      dist2 = qx * qx + qy * qy;
      // a degenerate neighbour edge (coinciding vertices) ties with this one at the shared corner: prefer the real edge
    }
    if (dist2 < best - (len2 < 1e-20 ? 0. : 1e-24)) { best = dist2; iseg = i; frac = f; }
  }
  int j = (iseg + 1) & 3;
  double ax, ay, bx, by;
  arb8_edge(P, iseg, s, ax, ay, bx, by);
  double ex = bx - ax, ey = by - ay;
  double rx = (1 - frac) * (v[2 * iseg + 8] - v[2 * iseg]) + frac * (v[2 * j + 8] - v[2 * j]);
  double ry = (1 - frac) * (v[2 * iseg + 9] - v[2 * iseg + 1]) + frac * (v[2 * j + 9] - v[2 * j + 1]);
  double rz = 2 * dz;
  if (ex * ex + ey * ey < 1e-20) {  // the section collapses to a point here (apex): use the edge of the opposite z face
    arb8_edge(P, iseg, s < 0.5 ? 1. : 0., ax, ay, bx, by);
    ex = bx - ax; ey = by - ay;
  }
  V3 n = v3(ey * rz, -ex * rz, ex * ry - ey * rx);  // (e,0) x r
  double mag = sqrt(dot(n, n));
  if (!(mag > 0)) return v3(0, 0, d.z >= 0 ? 1. : -1.);
  n = (1. / mag) * n;
  if (dot(n, d) < 0) n = v3(-n.x, -n.y, -n.z);
  return n;
}

// ---- TGeoXtru  P: nvert, nz, nvert x (x,y), nz x (z,x0,y0,scale)   (tutorials/AshraOptics.C:791-1021, src/AGeoUtil.cxx:84-125)
// A (possibly concave) polygon extruded along z; each section places the polygon at (x0,y0) with a scale, linearly interpolated
// between sections (TGeoXtru::SetCurrentZ).  A scaled and shifted copy of an edge stays parallel to it, so every lateral face
// is a planar trapezoid: one linear crossing per (slab, edge).  Evaluated like the general TGeoPgon: the next crossing above the
// last one is recomputed on demand (no candidate array), intervals are classified by Contains at their midpoints.
RB_HD inline bool xtru_contains(const double* P, V3 p) {
  int nv = (int)P[0], nz = (int)P[1];
  const double* V = P + 2;
  const double* sec = P + 2 + 2 * nv;
  if (p.z < sec[0] || p.z > sec[4 * (nz - 1)]) return false;
  int iz = 0;
  for (int i = 0; i + 1 < nz; i++)
    if (sec[4 * i] <= p.z) iz = i;
  double z0 = sec[4 * iz], dzs = sec[4 * (iz + 1)] - z0, f = dzs > 1e-8 ? (p.z - z0) / dzs : 0.;
  double x0 = sec[4 * iz + 1] + f * (sec[4 * (iz + 1) + 1] - sec[4 * iz + 1]), y0 = sec[4 * iz + 2] + f * (sec[4 * (iz + 1) + 2] - sec[4 * iz + 2]);
  double sc = sec[4 * iz + 3] + f * (sec[4 * (iz + 1) + 3] - sec[4 * iz + 3]);
  if (!(sc > 0)) return false;
  double x = (p.x - x0) / sc, y = (p.y - y0) / sc;
  bool in = false;  // crossing number
  for (int k = 0, j = nv - 1; k < nv; j = k++) {
    double xk = V[2 * k], yk = V[2 * k + 1], xj = V[2 * j], yj = V[2 * j + 1];
    if ((yk > y) != (yj > y) && x < (xj - xk) * (y - yk) / (yj - yk) + xk) in = !in;
  }
  return in;
}
RB_HD inline double xtru_next(const double* P, V3 p, V3 d, double last) {
  int nv = (int)P[0], nz = (int)P[1];
  const double* V = P + 2;
  const double* sec = P + 2 + 2 * nv;
  double best = RB_BIG;
  auto offer = [&](double t) {
    if (t > 1e-11 && t <= 1e29 && t > last && t < best) best = t;
  };
  if (d.z != 0)
    for (int i = 0; i < nz; i++) offer((sec[4 * i] - p.z) / d.z);
  for (int s = 0; s + 1 < nz; s++) {
    double z0 = sec[4 * s], z1 = sec[4 * (s + 1)], dzs = z1 - z0;
    if (dzs < 1e-8) continue;
    double f0 = (p.z - z0) / dzs, f1 = d.z / dzs;
    double dox = sec[4 * (s + 1) + 1] - sec[4 * s + 1], doy = sec[4 * (s + 1) + 2] - sec[4 * s + 2], dsc = sec[4 * (s + 1) + 3] - sec[4 * s + 3];
    double sc0 = sec[4 * s + 3] + f0 * dsc, sc1 = f1 * dsc;
    double bx = p.x - sec[4 * s + 1] - f0 * dox, by = p.y - sec[4 * s + 2] - f0 * doy, mx = d.x - f1 * dox, my = d.y - f1 * doy;
    for (int k = 0, j = nv - 1; k < nv; j = k++) {  // edge j -> k
      double ex = V[2 * k] - V[2 * j], ey = V[2 * k + 1] - V[2 * j + 1];
      double q0x = bx - sc0 * V[2 * j], q0y = by - sc0 * V[2 * j + 1], q1x = mx - sc1 * V[2 * j], q1y = my - sc1 * V[2 * j + 1];
      double den = q1x * ey - q1y * ex;
      if (den == 0) continue;
      double t = -(q0x * ey - q0y * ex) / den, zz = p.z + t * d.z;
      if (zz >= z0 - 1e-9 && zz <= z1 + 1e-9) offer(t);
    }
  }
  return best;
}
RB_HD inline double xtru_dist(const double* P, V3 p, V3 d, bool from_inside) {
  double prev = 0, last = 0;
  for (int it = 0; it < 4096; it++) {
    double t = xtru_next(P, p, d, last);
    if (t > 1e29) break;
    last = t;
    if (t - prev < 1e-12) { prev = t; continue; }
    bool in = xtru_contains(P, along(p, d, 0.5 * (prev + t)));
    if (from_inside ? !in : in) return prev;
    prev = t;
  }
  return from_inside ? prev : RB_BIG;
}
// nearest of: the two end planes, section planes where the outline jumps, the lateral trapezoids of the point's slab
RB_HD inline V3 xtru_normal(const double* P, V3 p, V3 d) {
  int nv = (int)P[0], nz = (int)P[1];
  const double* V = P + 2;
  const double* sec = P + 2 + 2 * nv;
  double best = RB_BIG;
  V3 n = v3(0, 0, 1);
  for (int i = 0; i < nz; i++) {
    bool cap = i == 0 || i == nz - 1;
    bool step = (i + 1 < nz && sec[4 * (i + 1)] - sec[4 * i] < 1e-8) || (i > 0 && sec[4 * i] - sec[4 * (i - 1)] < 1e-8);
    if (!cap && !step) continue;
    double s = fabs(p.z - sec[4 * i]);
    if (s < best) { best = s; n = v3(0, 0, 1); }
  }
  for (int s = 0; s + 1 < nz; s++) {
    double z0 = sec[4 * s], z1 = sec[4 * (s + 1)], dzs = z1 - z0;
    if (dzs < 1e-8 || p.z < z0 - 1e-6 || p.z > z1 + 1e-6) continue;
    double f = (p.z - z0) / dzs;
    double dox = (sec[4 * (s + 1) + 1] - sec[4 * s + 1]) / dzs, doy = (sec[4 * (s + 1) + 2] - sec[4 * s + 2]) / dzs, dsc = (sec[4 * (s + 1) + 3] - sec[4 * s + 3]) / dzs;
    double x0 = sec[4 * s + 1] + f * dzs * dox, y0 = sec[4 * s + 2] + f * dzs * doy, sc = sec[4 * s + 3] + f * dzs * dsc;
    for (int k = 0, j = nv - 1; k < nv; j = k++) {
      double ax = x0 + sc * V[2 * j], ay = y0 + sc * V[2 * j + 1], ex = sc * (V[2 * k] - V[2 * j]), ey = sc * (V[2 * k + 1] - V[2 * j + 1]);
      double len2 = ex * ex + ey * ey;
      if (len2 < 1e-20) continue;
      double ux = p.x - ax, uy = p.y - ay, fr = (ux * ex + uy * ey) / len2;
      fr = fr < 0 ? 0 : (fr > 1 ? 1 : fr);
      double qx = ux - fr * ex, qy = uy - fr * ey;
      // the face contains the edge direction (ex,ey,0) and the path of vertex j along z: (dox + dsc Vx, doy + dsc Vy, 1)
      double rx = dox + dsc * V[2 * j], ry = doy + dsc * V[2 * j + 1];
      V3 fn = v3(ey, -ex, ex * ry - ey * rx);
      double mag = sqrt(dot(fn, fn)), dist = sqrt(qx * qx + qy * qy) * sqrt(len2) / mag;  // in-plane foot distance -> distance to the face
      if (dist < best) { best = dist; n = (1. / mag) * fn; }
    }
  }
  if (dot(n, d) < 0) n = v3(-n.x, -n.y, -n.z);
  return n;
}

// ================================================================== shape dispatch, boolean composites
// DEPTH = remaining boolean nesting levels compiled in (scene build picks the instantiation).
RB_HD inline bool rb_is_bool(int type) { return type >= RBG_SHAPE_UNION && type <= RBG_SHAPE_SUBTRACTION; }
template <int DEPTH, unsigned SM> struct Csg {
  static RB_HD RB_NOINLINE bool contains(const DScene& sc, int sh, V3 p);
  static RB_HD RB_NOINLINE double dist_in(const DScene& sc, int sh, V3 p, V3 d, int& sel);
  static RB_HD RB_NOINLINE double dist_out(const DScene& sc, int sh, V3 p, V3 d, double step, int& sel);
  static RB_HD RB_NOINLINE V3 normal(const DScene& sc, int sh, V3 p, V3 d, int sel);
};

template <unsigned SM> RB_HD inline bool prim_contains(const DScene& sc, const DShape& s, V3 p) {
  const double* P = sc.dpar + s.ipar;
  switch (s.type) {
    case RBG_SHAPE_BBOX: if constexpr ((SM & RB_SBIT(RBG_SHAPE_BBOX)) != 0) return bbox_contains(P, p); else break;
    case RBG_SHAPE_TUBE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_TUBE)) != 0) return tube_contains(P, p); else break;
    case RBG_SHAPE_SPHERE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_SPHERE)) != 0) return sphere_contains(P, p); else break;
    case RBG_SHAPE_PARABOLOID: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PARABOLOID)) != 0) return para_contains(P, p); else break;
    case RBG_SHAPE_PGON: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PGON)) != 0) return poly_is_general(P) ? polyg_contains<false>(P, p) : poly_contains<false>(P, p); else break;
    case RBG_SHAPE_PCON: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PCON)) != 0) return poly_is_general(P) ? polyg_contains<true>(P, p) : poly_contains<true>(P, p); else break;
    case RBG_SHAPE_ASPHERE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_ASPHERE)) != 0) return asph_contains(P, p); else break;
    case RBG_SHAPE_WINSTON2D: if constexpr ((SM & RB_SBIT(RBG_SHAPE_WINSTON2D)) != 0) return win_contains(P, false, p); else break;
    case RBG_SHAPE_WINSTONPOLY: if constexpr ((SM & RB_SBIT(RBG_SHAPE_WINSTONPOLY)) != 0) return win_contains(P, true, p); else break;
    case RBG_SHAPE_ARB8: if constexpr ((SM & RB_SBIT(RBG_SHAPE_ARB8)) != 0) return arb8_contains(P, p); else break;
    case RBG_SHAPE_XTRU: if constexpr ((SM & RB_SBIT(RBG_SHAPE_XTRU)) != 0) return xtru_contains(P, p); else break;
  }
  return false;
}
template <unsigned SM> RB_HD inline double prim_dist_in(const DScene& sc, const DShape& s, V3 p, V3 d) {
  const double* P = sc.dpar + s.ipar;
  switch (s.type) {
    case RBG_SHAPE_BBOX: if constexpr ((SM & RB_SBIT(RBG_SHAPE_BBOX)) != 0) return bbox_dist_in(P, p, d); else break;
    case RBG_SHAPE_TUBE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_TUBE)) != 0) return tube_dist_in(P[0], P[1], P[2], p, d); else break;
    case RBG_SHAPE_SPHERE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_SPHERE)) != 0) return sphere_dist(P, p, d, true); else break;
    case RBG_SHAPE_PARABOLOID: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PARABOLOID)) != 0) return para_dist_in(P, p, d); else break;
    case RBG_SHAPE_PGON: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PGON)) != 0) return poly_is_general(P) ? polyg_dist<false>(P, p, d, true) : poly_dist_in<false>(P, p, d); else break;
    case RBG_SHAPE_PCON: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PCON)) != 0) return poly_is_general(P) ? polyg_dist<true>(P, p, d, true) : poly_dist_in<true>(P, p, d); else break;
    case RBG_SHAPE_ASPHERE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_ASPHERE)) != 0) return asph_dist4(P, p, d); else break;
    case RBG_SHAPE_WINSTON2D: if constexpr ((SM & RB_SBIT(RBG_SHAPE_WINSTON2D)) != 0) return win_dist_in(P, false, p, d); else break;
    case RBG_SHAPE_WINSTONPOLY: if constexpr ((SM & RB_SBIT(RBG_SHAPE_WINSTONPOLY)) != 0) return win_dist_in(P, true, p, d); else break;
    case RBG_SHAPE_ARB8: if constexpr ((SM & RB_SBIT(RBG_SHAPE_ARB8)) != 0) return arb8_dist(P, p, d, true); else break;
    case RBG_SHAPE_XTRU: if constexpr ((SM & RB_SBIT(RBG_SHAPE_XTRU)) != 0) return xtru_dist(P, p, d, true); else break;
  }
  return RB_BIG;
}
template <unsigned SM> RB_HD inline double prim_dist_out(const DScene& sc, const DShape& s, V3 p, V3 d, double step) {
  const double* P = sc.dpar + s.ipar;
  switch (s.type) {
    case RBG_SHAPE_BBOX: if constexpr ((SM & RB_SBIT(RBG_SHAPE_BBOX)) != 0) return bbox_dist_out(P, p, d, step); else break;
    case RBG_SHAPE_TUBE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_TUBE)) != 0) return tube_dist_out(P[0], P[1], P[2], p, d); else break;
    case RBG_SHAPE_SPHERE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_SPHERE)) != 0) return sphere_dist(P, p, d, false); else break;
    case RBG_SHAPE_PARABOLOID: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PARABOLOID)) != 0) return para_dist_out(P, p, d); else break;
    case RBG_SHAPE_PGON: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PGON)) != 0) return poly_is_general(P) ? polyg_dist<false>(P, p, d, false) : poly_dist_out<false>(P, p, d); else break;
    case RBG_SHAPE_PCON: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PCON)) != 0) return poly_is_general(P) ? polyg_dist<true>(P, p, d, false) : poly_dist_out<true>(P, p, d); else break;
    case RBG_SHAPE_ASPHERE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_ASPHERE)) != 0) return asph_dist_out(P, p, d, step); else break;
    case RBG_SHAPE_WINSTON2D: if constexpr ((SM & RB_SBIT(RBG_SHAPE_WINSTON2D)) != 0) return win_dist_out(P, false, p, d); else break;
    case RBG_SHAPE_WINSTONPOLY: if constexpr ((SM & RB_SBIT(RBG_SHAPE_WINSTONPOLY)) != 0) return win_dist_out(P, true, p, d); else break;
    case RBG_SHAPE_ARB8: if constexpr ((SM & RB_SBIT(RBG_SHAPE_ARB8)) != 0) return arb8_dist(P, p, d, false); else break;
    case RBG_SHAPE_XTRU: if constexpr ((SM & RB_SBIT(RBG_SHAPE_XTRU)) != 0) return xtru_dist(P, p, d, false); else break;
  }
  return RB_BIG;
}
template <unsigned SM> RB_HD inline V3 prim_normal(const DScene& sc, const DShape& s, V3 p, V3 d) {
  const double* P = sc.dpar + s.ipar;
  switch (s.type) {
    case RBG_SHAPE_BBOX: if constexpr ((SM & RB_SBIT(RBG_SHAPE_BBOX)) != 0) return bbox_normal(P, p, d); else break;
    case RBG_SHAPE_TUBE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_TUBE)) != 0) return tube_normal(P, p, d); else break;
    case RBG_SHAPE_SPHERE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_SPHERE)) != 0) return sphere_normal(P, p, d); else break;
    case RBG_SHAPE_PARABOLOID: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PARABOLOID)) != 0) return para_normal(P, p, d); else break;
    case RBG_SHAPE_PGON: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PGON)) != 0) return poly_is_general(P) ? polyg_normal<false>(P, p, d) : poly_normal<false>(P, p, d); else break;
    case RBG_SHAPE_PCON: if constexpr ((SM & RB_SBIT(RBG_SHAPE_PCON)) != 0) return poly_is_general(P) ? polyg_normal<true>(P, p, d) : poly_normal<true>(P, p, d); else break;
    case RBG_SHAPE_ASPHERE: if constexpr ((SM & RB_SBIT(RBG_SHAPE_ASPHERE)) != 0) return asph_normal(P, p, d); else break;
    case RBG_SHAPE_WINSTON2D: if constexpr ((SM & RB_SBIT(RBG_SHAPE_WINSTON2D)) != 0) return win_normal(P, false, p, d); else break;
    case RBG_SHAPE_WINSTONPOLY: if constexpr ((SM & RB_SBIT(RBG_SHAPE_WINSTONPOLY)) != 0) return win_normal(P, true, p, d); else break;
    case RBG_SHAPE_ARB8: if constexpr ((SM & RB_SBIT(RBG_SHAPE_ARB8)) != 0) return arb8_normal(P, p, d); else break;
    case RBG_SHAPE_XTRU: if constexpr ((SM & RB_SBIT(RBG_SHAPE_XTRU)) != 0) return xtru_normal(P, p, d); else break;
  }
  return v3(0, 0, 1);
}

// RB_PRIM_CALL: a translation unit may define it empty (before including this header) to inline the primitive
// dispatch into its callers — worthwhile only for instantiations with very few shape types.
#ifndef RB_PRIM_CALL
#define RB_PRIM_CALL RB_NOINLINE
#endif
template <unsigned SM> struct Csg<0, SM> {
  static RB_HD RB_PRIM_CALL bool contains(const DScene& sc, int sh, V3 p) { return prim_contains<SM>(sc, sc.shapes[sh], p); }
  static RB_HD RB_PRIM_CALL double dist_in(const DScene& sc, int sh, V3 p, V3 d, int& sel) { sel = 0; return prim_dist_in<SM>(sc, sc.shapes[sh], p, d); }
  static RB_HD RB_PRIM_CALL double dist_out(const DScene& sc, int sh, V3 p, V3 d, double step, int& sel) {
    sel = 0;
    return prim_dist_out<SM>(sc, sc.shapes[sh], p, d, step);
  }
  static RB_HD RB_PRIM_CALL V3 normal(const DScene& sc, int sh, V3 p, V3 d, int) { return prim_normal<SM>(sc, sc.shapes[sh], p, d); }
};

RB_HD inline V3 op_point(const DScene& sc, int m, V3 p) {
  if (m < 0) return p;
  if (m & RB_MAT_TRANS) {
    const double* t = sc.mats[m & ~RB_MAT_TRANS].t;
    return v3(p.x - t[0], p.y - t[1], p.z - t[2]);
  }
  return to_local(sc.mats[m], p);
}
RB_HD inline V3 op_vec(const DScene& sc, int m, V3 d) { return (m < 0 || (m & RB_MAT_TRANS)) ? d : to_local_vec(sc.mats[m], d); }

template <int DEPTH, unsigned SM> RB_HD RB_NOINLINE bool Csg<DEPTH, SM>::contains(const DScene& sc, int sh, V3 p) {
  const DShape s = sc.shapes[sh];
  if (!rb_is_bool(s.type)) return Csg<0, SM>::contains(sc, sh, p);  // one shared copy of the primitive code
  typedef Csg<DEPTH - 1, SM> Sub;
  bool l = Sub::contains(sc, s.left, op_point(sc, s.lmat, p));
  if ((SM & RB_SBIT(RBG_SHAPE_UNION)) != 0 && s.type == RBG_SHAPE_UNION) return l || Sub::contains(sc, s.right, op_point(sc, s.rmat, p));
  if (!l) return false;
  bool r = Sub::contains(sc, s.right, op_point(sc, s.rmat, p));
  return s.type == RBG_SHAPE_INTERSECTION ? r : !r;
}

template <int DEPTH, unsigned SM> RB_HD RB_NOINLINE double Csg<DEPTH, SM>::dist_in(const DScene& sc, int sh, V3 p, V3 d, int& sel) {
  const DShape s = sc.shapes[sh];
  sel = 0;
  if (!rb_is_bool(s.type)) return Csg<0, SM>::dist_in(sc, sh, p, d, sel);
  typedef Csg<DEPTH - 1, SM> Sub;
  V3 lp = op_point(sc, s.lmat, p), rp = op_point(sc, s.rmat, p), ld = op_vec(sc, s.lmat, d), rd = op_vec(sc, s.rmat, d);
  int s1 = 0, s2 = 0;
  if ((SM & RB_SBIT(RBG_SHAPE_UNION)) == 0 || s.type != RBG_SHAPE_UNION) {  // TGeoIntersection / TGeoSubtraction :: DistFromInside
    double d1 = Sub::dist_in(sc, s.left, lp, ld, s1);
    double d2 = s.type == RBG_SHAPE_INTERSECTION ? Sub::dist_in(sc, s.right, rp, rd, s2) : Sub::dist_out(sc, s.right, rp, rd, RB_BIG, s2);
    if (d1 < d2) { sel = 1 | (s1 << 2); return d1; }
    sel = 2 | (s2 << 2);
    return d2;
  }
  // TGeoUnion::DistFromInside: leave whichever operand holds the point, continue through the other
  bool in1 = Sub::contains(sc, s.left, lp), in2 = Sub::contains(sc, s.right, rp);
  double d1 = 0, d2 = 0, snxt = 0;
  if (in1) d1 = Sub::dist_in(sc, s.left, lp, ld, s1);
  if (in2) d2 = Sub::dist_in(sc, s.right, rp, rd, s2);
  if (!(in1 || in2)) {
    d1 = Sub::dist_out(sc, s.left, lp, ld, RB_BIG, s1);
    if (d1 < 2. * RB_TOL) {
      double eps = d1 + RB_TOL;
      in1 = true;
      d1 = Sub::dist_in(sc, s.left, along(lp, ld, eps), ld, s1) + eps;
    } else {
      d2 = Sub::dist_out(sc, s.right, rp, rd, RB_BIG, s2);
      if (d2 < 2. * RB_TOL) {
        double eps = d2 + RB_TOL;
        in2 = true;
        d2 = Sub::dist_in(sc, s.right, along(rp, rd, eps), rd, s2) + eps;
      }
    }
  }
  V3 master = p;
  for (int guard = 0; guard < 64 && (in1 || in2); guard++) {
    if (in1 && (!in2 || d1 < d2)) {
      snxt += d1;
      sel = 1 | (s1 << 2);
      in1 = false;
      master = along(master, d, d1);
      V3 q = op_point(sc, s.rmat, along(master, d, (1. + d1) * RB_TOL));
      in2 = Sub::contains(sc, s.right, q);
      if (!in2) return snxt;
      d2 = Sub::dist_in(sc, s.right, q, rd, s2);
      if (d2 < RB_TOL) return snxt;
      d2 += (1. + d1) * RB_TOL;
    } else {
      snxt += d2;
      sel = 2 | (s2 << 2);
      in2 = false;
      master = along(master, d, d2);
      V3 q = op_point(sc, s.lmat, along(master, d, (1. + d2) * RB_TOL));
      in1 = Sub::contains(sc, s.left, q);
      if (!in1) return snxt;
      d1 = Sub::dist_in(sc, s.left, q, ld, s1);
      if (d1 < RB_TOL) return snxt;
      d1 += (1. + d2) * RB_TOL;
    }
  }
  return snxt;
}

template <int DEPTH, unsigned SM> RB_HD RB_NOINLINE double Csg<DEPTH, SM>::dist_out(const DScene& sc, int sh, V3 p, V3 d, double step, int& sel) {
  const DShape s = sc.shapes[sh];
  sel = 0;
  if (!rb_is_bool(s.type)) return Csg<0, SM>::dist_out(sc, sh, p, d, step, sel);
  typedef Csg<DEPTH - 1, SM> Sub;
  V3 ld = op_vec(sc, s.lmat, d), rd = op_vec(sc, s.rmat, d);
  V3 lp = op_point(sc, s.lmat, p), rp = op_point(sc, s.rmat, p);
  int s1 = 0, s2 = 0;
  if ((SM & RB_SBIT(RBG_SHAPE_UNION)) != 0 && s.type == RBG_SHAPE_UNION) {
    double d1 = Sub::dist_out(sc, s.left, lp, ld, step, s1), d2 = Sub::dist_out(sc, s.right, rp, rd, step, s2);
    if (d1 < d2) { sel = 1 | (s1 << 2); return d1; }
    sel = 2 | (s2 << 2);
    return d2;
  }
  V3 master = p;
  if ((SM & RB_SBIT(RBG_SHAPE_INTERSECTION)) != 0 && s.type == RBG_SHAPE_INTERSECTION) {
    bool inl = Sub::contains(sc, s.left, lp), inr = Sub::contains(sc, s.right, rp);
    double snext = 0.0, d1, d2;
    if (inl && inr) {
      d1 = Sub::dist_in(sc, s.left, lp, ld, s1);
      d2 = Sub::dist_in(sc, s.right, rp, rd, s2);
      if (d1 < 1.E-3) inl = false;
      if (d2 < 1.E-3) inr = false;
      if (inl && inr) return snext;
    }
    // either operand missing means no hit: try the cheaper primitive first (same result, fewer instructions)
    const bool right_first = sc.shapes[s.right].type < sc.shapes[s.left].type && !rb_is_bool(sc.shapes[s.left].type) ? false : sc.shapes[s.right].type != RBG_SHAPE_SPHERE;
    for (int guard = 0; guard < 64; guard++) {
      d1 = d2 = 0;
      if (right_first && !inr) {
        d2 = rb_max(Sub::dist_out(sc, s.right, rp, rd, RB_BIG, s2), RB_TOL);
        if (d2 > 1E20) return RB_BIG;
      }
      if (!inl) {
        d1 = rb_max(Sub::dist_out(sc, s.left, lp, ld, RB_BIG, s1), RB_TOL);
        if (d1 > 1E20) return RB_BIG;
      }
      if (!right_first && !inr) {
        d2 = rb_max(Sub::dist_out(sc, s.right, rp, rd, RB_BIG, s2), RB_TOL);
        if (d2 > 1E20) return RB_BIG;
      }
      if (d1 > d2) {
        snext += d1;
        sel = 1 | (s1 << 2);
        inl = true;
        master = along(master, d, d1);
        lp = op_point(sc, s.lmat, master);
        rp = op_point(sc, s.rmat, master);
        inr = Sub::contains(sc, s.right, along(rp, rd, RB_TOL));
        if (inr) return snext;
      } else {
        snext += d2;
        sel = 2 | (s2 << 2);
        inr = true;
        master = along(master, d, d2);
        lp = op_point(sc, s.lmat, master);
        rp = op_point(sc, s.rmat, master);
        inl = Sub::contains(sc, s.left, along(lp, ld, RB_TOL));
        if (inl) return snext;
      }
    }
    return RB_BIG;
  }
  // TGeoSubtraction::DistFromOutside
  if ((SM & RB_SBIT(RBG_SHAPE_SUBTRACTION)) == 0) return RB_BIG;
  bool inside = Sub::contains(sc, s.right, rp);
  double snxt = 0., epsil = 0.;
  for (int guard = 0; guard < 64; guard++) {
    if (inside) {
      double d1 = Sub::dist_in(sc, s.right, rp, rd, s2);
      sel = 2 | (s2 << 2);
      snxt += d1 + epsil;
      master = along(master, d, d1 + 1E-8);
      epsil = 1.E-8;
      if (Sub::contains(sc, s.left, op_point(sc, s.lmat, master))) return snxt;
    }
    lp = op_point(sc, s.lmat, master);
    double d2 = Sub::dist_out(sc, s.left, lp, ld, RB_BIG, s1);
    if (d2 > 1E20) return RB_BIG;
    rp = op_point(sc, s.rmat, master);
    int s2b = 0;
    double d1 = Sub::dist_out(sc, s.right, rp, rd, RB_BIG, s2b);
    if (d2 < d1 - RB_TOL) {
      sel = 1 | (s1 << 2);
      return snxt + d2 + epsil;
    }
    snxt += d1 + epsil;
    master = along(master, d, d1 + 1E-8);
    epsil = 1.E-8;
    rp = op_point(sc, s.rmat, master);
    inside = true;
  }
  return RB_BIG;
}

template <int DEPTH, unsigned SM> RB_HD RB_NOINLINE V3 Csg<DEPTH, SM>::normal(const DScene& sc, int sh, V3 p, V3 d, int sel) {
  const DShape s = sc.shapes[sh];
  if (!rb_is_bool(s.type)) return Csg<0, SM>::normal(sc, sh, p, d, 0);
  typedef Csg<DEPTH - 1, SM> Sub;
  int side = sel & 3;
  if (side == 0) {
    bool inl = Sub::contains(sc, s.left, op_point(sc, s.lmat, p)), inr = Sub::contains(sc, s.right, op_point(sc, s.rmat, p));
    if (s.type == RBG_SHAPE_SUBTRACTION) side = inr ? 2 : 1;
    else if (s.type == RBG_SHAPE_UNION) side = inl ? 1 : 2;
    else side = inl ? 2 : 1;
  }
  int m = side == 1 ? s.lmat : s.rmat;
  V3 ln = Sub::normal(sc, side == 1 ? s.left : s.right, op_point(sc, m, p), op_vec(sc, m, d), sel >> 2);
  return (m < 0 || (m & RB_MAT_TRANS)) ? ln : to_master_vec(sc.mats[m], ln);
}

// ================================================================== flat leaf evaluators
// The shapes of real telescopes are almost always a primitive, a boolean of two primitives (mirror facet = sphere * polygon,
// light guide = polygon - Winston cone, camera housing = box - box) or a chain of unions (spider = box + box + box + tube).
// Leaf<K> evaluates these without the generic walk's recursion of non-inlined calls:
//   * a primitive is dispatched to its code directly;
//   * a boolean of two primitives whose (operation, left type, right type) is in the instantiation's list K::combos is
//     evaluated by Bool2<...>: TGeoIntersection / TGeoSubtraction / TGeoUnion with both operands inlined and their types known
//     at compile time (one copy of each primitive routine, no dispatch, everything in registers);
//   * a chain of unions is evaluated as a flat loop over its primitives.
// Everything else goes to Csg<DEPTH>.  The flat paths perform exactly the arithmetic of the generic walk (same operations in the
// same order, TGeoBoolNode algorithms included), so both give identical bits; SceneBuilder::leaf_kind classifies the nodes.
template <int T> struct PrimT;  // primitive routines by compile-time type (P = the shape's parameter block)
#define RB_PRIMT(T, CONTAINS, DIN, DOUT, NORMAL)                                                          \
  template <> struct PrimT<T> {                                                                            \
    static RB_HD inline bool contains(const double* P, V3 p) { return CONTAINS; }                          \
    static RB_HD inline double dist_in(const double* P, V3 p, V3 d) { return DIN; }                        \
    static RB_HD inline double dist_out(const double* P, V3 p, V3 d, double step) { (void)step; return DOUT; } \
    static RB_HD inline V3 normal(const double* P, V3 p, V3 d) { return NORMAL; }                          \
  };
RB_PRIMT(RBG_SHAPE_BBOX, bbox_contains(P, p), bbox_dist_in(P, p, d), bbox_dist_out(P, p, d, step), bbox_normal(P, p, d))
RB_PRIMT(RBG_SHAPE_TUBE, tube_contains(P, p), tube_dist_in(P[0], P[1], P[2], p, d), tube_dist_out(P[0], P[1], P[2], p, d), tube_normal(P, p, d))
RB_PRIMT(RBG_SHAPE_SPHERE, sphere_contains(P, p), sphere_dist(P, p, d, true), sphere_dist(P, p, d, false, step), sphere_normal(P, p, d))
RB_PRIMT(RBG_SHAPE_PARABOLOID, para_contains(P, p), para_dist_in(P, p, d), para_dist_out(P, p, d), para_normal(P, p, d))
// polygons / polycones reach the typed path only in their convex form (rmin = 0, full azimuth): leaf_kind sees to that
RB_PRIMT(RBG_SHAPE_PGON, poly_contains<false>(P, p), poly_dist_in<false>(P, p, d), poly_dist_out<false>(P, p, d), poly_normal<false>(P, p, d))
RB_PRIMT(RBG_SHAPE_PCON, poly_contains<true>(P, p), poly_dist_in<true>(P, p, d), poly_dist_out<true>(P, p, d), poly_normal<true>(P, p, d))
RB_PRIMT(RBG_SHAPE_ASPHERE, asph_contains(P, p), asph_dist4(P, p, d), asph_dist_out(P, p, d, step), asph_normal(P, p, d))
RB_PRIMT(RBG_SHAPE_WINSTON2D, win_contains(P, false, p), win_dist_in(P, false, p, d), win_dist_out(P, false, p, d), win_normal(P, false, p, d))
RB_PRIMT(RBG_SHAPE_WINSTONPOLY, win_contains(P, true, p), win_dist_in(P, true, p, d), win_dist_out(P, true, p, d), win_normal(P, true, p, d))
#undef RB_PRIMT

// leaf code of a placed node: class in the low 4 bits; for RB_LEAF_BOOL2 the operation and the operand types above them
#define RB_LEAF_CODE(OP, TL, TR) (RB_LEAF_BOOL2 | ((OP) << 4) | ((TL) << 8) | ((TR) << 12))
template <int OP_, int TL_, int TR_> struct B2 {
  static constexpr int OP = OP_, TL = TL_, TR = TR_, code = RB_LEAF_CODE(OP_, TL_, TR_);
};
template <class... Cs> struct Combos {};

// boolean of two primitives of known types: ROOT's TGeoBoolNode algorithms (as in Csg<DEPTH>) on inlined operands
template <class C> struct Bool2 {
  typedef PrimT<C::TL> PL;
  typedef PrimT<C::TR> PR;
  static RB_HD inline bool contains(const DScene& sc, const DShape& s, V3 p) {
    const double *A = sc.dpar + sc.shapes[s.left].ipar, *B = sc.dpar + sc.shapes[s.right].ipar;
    bool l = PL::contains(A, op_point(sc, s.lmat, p));
    if (C::OP == RBG_SHAPE_UNION) return l || PR::contains(B, op_point(sc, s.rmat, p));
    if (!l) return false;
    bool r = PR::contains(B, op_point(sc, s.rmat, p));
    return C::OP == RBG_SHAPE_INTERSECTION ? r : !r;
  }
  // TGeoIntersection / TGeoSubtraction :: DistFromInside (a union's is iterative and rare: left to the generic walk)
  static RB_HD inline double dist_in(const DScene& sc, const DShape& s, V3 p, V3 d, int& sel) {
    const double *A = sc.dpar + sc.shapes[s.left].ipar, *B = sc.dpar + sc.shapes[s.right].ipar;
    V3 lp = op_point(sc, s.lmat, p), rp = op_point(sc, s.rmat, p), ld = op_vec(sc, s.lmat, d), rd = op_vec(sc, s.rmat, d);
    double d1 = PL::dist_in(A, lp, ld);
    double d2 = C::OP == RBG_SHAPE_INTERSECTION ? PR::dist_in(B, rp, rd) : PR::dist_out(B, rp, rd, RB_BIG);
    if (d1 < d2) { sel = 1; return d1; }
    sel = 2;
    return d2;
  }
  // returns false when the case is left to the generic walk (start point inside an intersection: not a navigation case)
  static RB_HD inline bool dist_out(const DScene& sc, const DShape& s, V3 p, V3 d, double step, int& sel, double& out) {
    const double *A = sc.dpar + sc.shapes[s.left].ipar, *B = sc.dpar + sc.shapes[s.right].ipar;
    sel = 0;
    V3 ld = op_vec(sc, s.lmat, d), rd = op_vec(sc, s.rmat, d);
    V3 lp = op_point(sc, s.lmat, p), rp = op_point(sc, s.rmat, p);
    if (C::OP == RBG_SHAPE_UNION) {
      double d1 = PL::dist_out(A, lp, ld, step), d2 = PR::dist_out(B, rp, rd, step);
      if (d1 < d2) { sel = 1; out = d1; }
      else { sel = 2; out = d2; }
      return true;
    }
    V3 master = p;
    if (C::OP == RBG_SHAPE_INTERSECTION) {
      bool inl = PL::contains(A, lp), inr = PR::contains(B, rp);
      if (inl && inr) return false;
      // either operand missing means no hit: the cheaper primitive goes first, in the order the generic walk uses
      constexpr bool right_first = C::TR < C::TL ? false : C::TR != RBG_SHAPE_SPHERE;
      double snext = 0.0;
      out = RB_BIG;
      for (int guard = 0; guard < 64; guard++) {
        double d1 = 0, d2 = 0;
        // (the operands get what is left of the caller's limit: one that starts beyond it reports no crossing, and so does the
        // intersection — the same outcome as the `snext > step` test below, reached earlier)
        const double left = step - snext;
        if (right_first && !inr) {
          d2 = rb_max(PR::dist_out(B, rp, rd, left), RB_TOL);
          if (d2 > 1E20) return true;
        }
        if (!inl) {
          d1 = rb_max(PL::dist_out(A, lp, ld, left), RB_TOL);
          if (d1 > 1E20) return true;
        }
        if (!right_first && !inr) {
          d2 = rb_max(PR::dist_out(B, rp, rd, left), RB_TOL);
          if (d2 > 1E20) return true;
        }
        const bool left_entered = d1 > d2;
        const double adv = left_entered ? d1 : d2;
        snext += adv;
        if (snext > step) return true;  // the solid starts beyond the caller's limit (the step already found): no crossing to report
        sel = left_entered ? 1 : 2;
        master = along(master, d, adv);
        lp = op_point(sc, s.lmat, master);
        rp = op_point(sc, s.rmat, master);
        if (left_entered) {
          inl = true;
          inr = PR::contains(B, along(rp, rd, RB_TOL));
          if (inr) { out = snext; return true; }
        } else {
          inr = true;
          inl = PL::contains(A, along(lp, ld, RB_TOL));
          if (inl) { out = snext; return true; }
        }
      }
      return true;
    }
    // TGeoSubtraction::DistFromOutside
    bool inside = PR::contains(B, rp);
    double snxt = 0., epsil = 0.;
    out = RB_BIG;
    for (int guard = 0; guard < 64; guard++) {
      if (inside) {
        double d1 = PR::dist_in(B, rp, rd);
        sel = 2;
        snxt += d1 + epsil;
        master = along(master, d, d1 + 1E-8);
        epsil = 1.E-8;
        if (PL::contains(A, op_point(sc, s.lmat, master))) { out = snxt; return true; }
      }
      lp = op_point(sc, s.lmat, master);
      double d2 = PL::dist_out(A, lp, ld, RB_BIG);
      if (d2 > 1E20 || snxt + d2 > step) return true;  // misses the left solid, or reaches it beyond the caller's limit
      rp = op_point(sc, s.rmat, master);
      double d1 = PR::dist_out(B, rp, rd, RB_BIG);
      if (d2 < d1 - RB_TOL) {
        sel = 1;
        out = snxt + d2 + epsil;
        return true;
      }
      snxt += d1 + epsil;
      master = along(master, d, d1 + 1E-8);
      epsil = 1.E-8;
      rp = op_point(sc, s.rmat, master);
      inside = true;
    }
    return true;
  }
  static RB_HD inline V3 normal(const DScene& sc, const DShape& s, V3 p, V3 d, int side) {
    const int m = side == 1 ? s.lmat : s.rmat;
    const double* P = sc.dpar + sc.shapes[side == 1 ? s.left : s.right].ipar;
    V3 lp = op_point(sc, m, p), ld = op_vec(sc, m, d);
    V3 ln = side == 1 ? PL::normal(P, lp, ld) : PR::normal(P, lp, ld);
    return (m < 0 || (m & RB_MAT_TRANS)) ? ln : to_master_vec(sc.mats[m], ln);
  }
};

template <class K> struct Leaf {
  typedef Csg<K::depth, K::shapes> G;
  typedef Csg<0, K::shapes> G0;
  static constexpr unsigned SM = K::shapes;

  // ---- chain of unions ((a + b) + c) + d: operands visited from the outermost right operand inwards; on equal distances
  // TGeoUnion keeps the right operand, i.e. the outer one.  `sel` is the selection path the generic walk would record.
  static RB_HD RB_NOINLINE bool unions_contains(const DScene& sc, int sh, V3 p) {
    while (true) {
      const DShape s = sc.shapes[sh];
      if (!rb_is_bool(s.type)) return G0::contains(sc, sh, p);
      if (G0::contains(sc, s.right, op_point(sc, s.rmat, p))) return true;
      p = op_point(sc, s.lmat, p);
      sh = s.left;
    }
  }
  static RB_HD RB_NOINLINE double unions_dist_out(const DScene& sc, int sh, V3 p, V3 d, double step, int& sel) {
    double best = RB_BIG;
    int path = 0, shift = 0, dummy = 0;
    sel = 0;
    bool first = true;
    while (true) {
      const DShape s = sc.shapes[sh];
      if (!rb_is_bool(s.type)) {
        double t = G0::dist_out(sc, sh, p, d, step, dummy);
        if (first || t < best) { best = t; sel = path; }
        return best;
      }
      double t = G0::dist_out(sc, s.right, op_point(sc, s.rmat, p), op_vec(sc, s.rmat, d), step, dummy);
      if (first || t < best) { best = t; sel = path | (2 << shift); }
      first = false;
      path |= 1 << shift;
      shift += 2;
      p = op_point(sc, s.lmat, p);
      d = op_vec(sc, s.lmat, d);
      sh = s.left;
    }
  }

  // ---- typed two-primitive booleans: try every combination of the instantiation's list
  template <class... Cs> static RB_HD inline bool combo_contains(Combos<Cs...>, int code, const DScene& sc, const DShape& s, V3 p, bool& out) {
    return ((code == Cs::code ? (out = Bool2<Cs>::contains(sc, s, p), true) : false) || ...);
  }
  template <class... Cs> static RB_HD inline bool combo_dist_in(Combos<Cs...>, int code, const DScene& sc, const DShape& s, V3 p, V3 d, int& sel, double& out) {
    return ((code == Cs::code && Cs::OP != RBG_SHAPE_UNION ? (out = Bool2<Cs>::dist_in(sc, s, p, d, sel), true) : false) || ...);
  }
  template <class... Cs>
  static RB_HD inline bool combo_dist_out(Combos<Cs...>, int code, const DScene& sc, const DShape& s, V3 p, V3 d, double step, int& sel, double& out) {
    return ((code == Cs::code ? Bool2<Cs>::dist_out(sc, s, p, d, step, sel, out) : false) || ...);
  }
  template <class... Cs> static RB_HD inline bool combo_normal(Combos<Cs...>, int code, const DScene& sc, const DShape& s, V3 p, V3 d, int side, V3& out) {
    return ((code == Cs::code ? (out = Bool2<Cs>::normal(sc, s, p, d, side), true) : false) || ...);
  }

  // ---- dispatch on the node's class
  static RB_HD inline bool contains(const DScene& sc, int leaf, int sh, V3 p) {
    if (leaf == RB_LEAF_PRIM) return prim_contains<SM>(sc, sc.shapes[sh], p);
    if constexpr (K::depth > 0) {
      if ((leaf & 15) == RB_LEAF_BOOL2) {
        bool out;
        if (combo_contains(typename K::combos(), leaf, sc, sc.shapes[sh], p, out)) return out;
      }
      if ((SM & RB_SBIT(RBG_SHAPE_UNION)) != 0 && leaf == RB_LEAF_UNIONS) return unions_contains(sc, sh, p);
    }
    return G::contains(sc, sh, p);
  }
  static RB_HD inline double dist_in(const DScene& sc, int leaf, int sh, V3 p, V3 d, int& sel) {
    sel = 0;
    if (leaf == RB_LEAF_PRIM) return prim_dist_in<SM>(sc, sc.shapes[sh], p, d);
    if constexpr (K::depth > 0) {
      if ((leaf & 15) == RB_LEAF_BOOL2) {
        double out;
        if (combo_dist_in(typename K::combos(), leaf, sc, sc.shapes[sh], p, d, sel, out)) return out;
      }
    }
    return G::dist_in(sc, sh, p, d, sel);
  }
  static RB_HD inline double dist_out(const DScene& sc, int leaf, int sh, V3 p, V3 d, double step, int& sel) {
    sel = 0;
    if (leaf == RB_LEAF_PRIM) return prim_dist_out<SM>(sc, sc.shapes[sh], p, d, step);
    if constexpr (K::depth > 0) {
      if ((leaf & 15) == RB_LEAF_BOOL2) {
        double out;
        if (combo_dist_out(typename K::combos(), leaf, sc, sc.shapes[sh], p, d, step, sel, out)) return out;
      }
      if ((SM & RB_SBIT(RBG_SHAPE_UNION)) != 0 && leaf == RB_LEAF_UNIONS) return unions_dist_out(sc, sh, p, d, step, sel);
    }
    return G::dist_out(sc, sh, p, d, step, sel);
  }
  static RB_HD inline V3 normal(const DScene& sc, int leaf, int sh, V3 p, V3 d, int sel) {
    if (leaf == RB_LEAF_PRIM) return prim_normal<SM>(sc, sc.shapes[sh], p, d);
    if constexpr (K::depth > 0) {
      if ((leaf & 15) == RB_LEAF_BOOL2 && (sel & 3) != 0) {  // side 0 = no boundary recorded (start on the surface): generic
        V3 out;
        if (combo_normal(typename K::combos(), leaf, sc, sc.shapes[sh], p, d, sel & 3, out)) return out;
      }
    }
    return G::normal(sc, sh, p, d, sel);
  }
};

// point test of a placed node (world coordinates); not inlined: it has many call sites (point location, relocation after a
// reflection), none of them inside the distance loops
template <class K> RB_HD RB_NOINLINE bool node_contains(const DScene& sc, int node, V3 q) {
  const DNode& nd = sc.nodes[node];
  return Leaf<K>::contains(sc, nd.leaf, nd.shape, to_local(nd.g, q));
}
// The line through lp along ld against the node's own box (separating axes d x e_k, no division), and the box entirely behind
// the start point.  fp32 on local coordinates: the margins cover the roundings (|lp| relative) on top of the padded widths.
RB_HD inline bool local_box_missed(const DNode& nd, V3 lp, V3 ld) {
  const float px = (float)lp.x - nd.lc[0], py = (float)lp.y - nd.lc[1], pz = (float)lp.z - nd.lc[2];
  const float dx = (float)ld.x, dy = (float)ld.y, dz = (float)ld.z;
  const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
  const float eps = (2e-6f * (fabsf(px) + fabsf(py) + fabsf(pz)) + 1e-3f) * (ax + ay + az);
  const float hx = nd.lh[0], hy = nd.lh[1], hz = nd.lh[2];
  if (fabsf(py * dz - pz * dy) > hy * az + hz * ay + eps) return true;
  if (fabsf(pz * dx - px * dz) > hz * ax + hx * az + eps) return true;
  if (fabsf(px * dy - py * dx) > hx * ay + hy * ax + eps) return true;
  return px * dx + py * dy + pz * dz > hx * ax + hy * ay + hz * az + eps;
}
RB_HD inline bool node_box_holds(const DNode& nd, V3 q) {
  const float x = (float)q.x, y = (float)q.y, z = (float)q.z;  // the boxes are padded for this rounding
  return x >= nd.blo[0] && x <= nd.bhi[0] && y >= nd.blo[1] && y <= nd.bhi[1] && z >= nd.blo[2] && z <= nd.bhi[2];
}

// ================================================================== flattened navigation
struct RayReg {           // register-resident ray state
  V3 p, d;
  double t, lambda;
  int cur;                // physical node containing the point, -1 = outside the top volume
  int status, npoints, last_node;
  uint32_t ndraw;
  int on_boundary;
};

// first (lowest id = daughter order) child of `node` with id > after whose shape contains q; -1 if none
template <class K> RB_HD inline int child_containing(const DScene& sc, int node, V3 q, int skip, int after = -1) {
  const DNode& nd = sc.nodes[node];
  int best = -1;
  int i = nd.bvh_count > 0 ? nd.bvh_first : -1;
  const float qx = (float)q.x, qy = (float)q.y, qz = (float)q.z;  // boxes are padded for this rounding
  while (i >= 0) {
    const DBvh& b = sc.bvh[i];
    bool in = qx >= b.lo[0] && qx <= b.hi[0] && qy >= b.lo[1] && qy <= b.hi[1] && qz >= b.lo[2] && qz <= b.hi[2];
    if (!in) { i = b.skip; continue; }
    if (b.child >= 0) {
      int c = b.child;
      if (c != skip && c > after && (best < 0 || c < best)) {
        if (node_contains<K>(sc, c, q)) best = c;
      }
      i = b.skip;
    } else i = i + 1;
  }
  return best;
}
// Descent of TGeoNavigator::SearchNode(downwards) from `node`.  Nodes placed with AddNodeOverlap ("MANY") may share space with
// their sisters: when the first daughter holding q is such a node, every later daughter holding q joins its cluster
// (GetTouchedCluster) and FindInCluster decides — a member whose branch reaches an ordinary ("ONLY") node wins at once, otherwise
// the member whose branch ends deepest, the first one on ties; the node FindNextBoundary announced (`prefer`, fNextNode) wins like
// an ONLY branch.  LV bounds the nesting of clusters inside cluster members.
template <class K, int LV> struct Locate {
  static RB_HD int down(const DScene& sc, int node, V3 q, int skip, bool& only, int prefer) {
    while (true) {
      int c = child_containing<K>(sc, node, q, skip);
      if (c < 0) return node;
      if constexpr ((K::phys & RB_PH_OVERLAP) != 0 && LV > 0) {
        if (sc.nodes[c].overlap) {
          int best = -1, best_level = -1;
          for (int m = c; m >= 0; m = child_containing<K>(sc, node, q, skip, m)) {
            bool o = !sc.nodes[m].overlap;
            int r = Locate<K, LV - 1>::down(sc, m, q, -1, o, prefer);
            if (o || m == prefer) { only = true; return r; }
            if (sc.nodes[r].level > best_level) { best = r; best_level = sc.nodes[r].level; }
          }
          return best;
        }
      }
      skip = -1;
      only = true;
      node = c;
    }
  }
};
// TGeoNavigator::SearchNode(downwards=false, skip) starting at `node`; returns the deepest node containing q or -1
template <class K> RB_HD inline int search_node(const DScene& sc, int node, V3 q, int skip, bool check_current, int prefer = -1) {
  if (check_current) {
    while (true) {
      if (node < 0) return -1;
      const DNode& nd = sc.nodes[node];
      bool inside = node == skip ? true : node_contains<K>(sc, node, q);
      if constexpr ((K::phys & RB_PH_OVERLAP) != 0) {
        // GotoSafeLevel: the search restarts from the first ordinary node above a run of overlapping ones
        if (inside && nd.overlap && nd.mother >= 0) {
          int up = nd.mother;
          while (sc.nodes[up].overlap && sc.nodes[up].mother >= 0) up = sc.nodes[up].mother;
          node = up;
          continue;
        }
      }
      if (inside) break;
      skip = node;
      node = nd.mother;
    }
  }
  bool only = false;
  return Locate<K, 2>::down(sc, node, q, skip, only, prefer);
}

#ifndef RB_MAXVIS
#define RB_MAXVIS 12
#endif
struct StepOut {
  double step;
  int next;        // node now containing the point (-1 outside)
  int crossed;     // node whose shape boundary was crossed (for FindNormal), -1 none
  int sel;         // boolean-operand selection path of the crossed shape
  int from;        // node the step started in (-1 outside)
  int nvis;        // daughters of `from` whose padded AABB the ray touched before the boundary; -1 = not recorded
  int vis[RB_MAXVIS];     // ... ordered by the parameter at which the ray enters the box
  float tin[RB_MAXVIS];
};

RB_HD inline double locate_extra(const DScene& sc, int node, double step) {
  double trmax = 1.;
  if (node >= 0) {
    const double* tr = sc.nodes[node].g.t;
    trmax += fabs(tr[0]) + fabs(tr[1]) + fabs(tr[2]);
  }
  return 100. * (trmax + step) * RB_TOL;
}

// TGeoNavigator::FindNextBoundaryAndStep(Big) on the flattened scene
// ---- one TGeoNavigator::FindNextBoundaryAndStep, split into phases so that the block-lock-step kernel can
// put a barrier between them (all warps of an SM then execute the same code region: instruction-cache reuse)
struct NavStep {
  StepOut o;
  double extra, best, s_exit;
  int sel_exit, enter, esel;
  int mode;      // 0 = finished in nb_begin, 1 = daughters to examine
  int bvh_next;  // >= 0: traversal stopped because o.vis was full; resume here after evaluating the batch
  int xkind, xnode, xsel;  // overlap extension (nb_many): 0 none, 1 = left the mother `xnode` of an overlapping node, 2 = met its sister `xnode`
  // point location behind the crossed boundary, requested by nb_begin / nb_finish and done at one place (nb_locate):
  // search_node(loc_node, point pushed by locate_extra(loc_node), loc_skip, loc_check, loc_prefer); loc_node = -2: none pending
  int loc_node, loc_skip, loc_check, loc_prefer;
};

// phase A: boundary push, outside-world entry, DistFromInside of the current shape
template <class K> RB_HD inline void nb_begin(const DScene& sc, RayReg& r, bool push_quirk, NavStep& st) {
  StepOut& o = st.o;
  o.crossed = -1;
  o.sel = 0;
  o.from = r.cur;
  o.nvis = -1;
  o.next = -1;
  st.mode = 0;
  st.bvh_next = -1;
  st.enter = -1;
  st.esel = 0;
  st.xkind = 0;
  st.loc_node = -2;
  st.loc_skip = -1;
  st.loc_check = 0;
  st.loc_prefer = -1;
  double extra = (r.on_boundary && push_quirk) ? RB_TOL : 0.0;
  st.extra = extra;
  r.on_boundary = 0;
  r.p = along(r.p, r.d, extra);
  if (r.cur < 0) {
    int sel = 0;
    double s = Leaf<K>::dist_out(sc, sc.top_leaf, sc.top_shape, r.p, r.d, RB_BIG, sel);
    if (s > 1e29) { o.step = RB_BIG; o.next = -1; return; }
    if (s <= 0) { s = 0.0; o.step = 0.0; r.p = along(r.p, r.d, -extra); }
    else o.step = s + extra;
    r.p = along(r.p, r.d, s);
    r.on_boundary = 1;
    o.crossed = 0;
    o.sel = sel;
    st.loc_node = 0;
    return;
  }
  const DNode& cn = sc.nodes[r.cur];
  st.sel_exit = 0;
  st.s_exit = Leaf<K>::dist_in(sc, cn.leaf, cn.shape, to_local(cn.g, r.p), to_local_vec(cn.g, r.d), st.sel_exit);
  if (st.s_exit <= RB_TOL) {
    o.step = RB_TOL;
    r.p = along(r.p, r.d, o.step);
    r.on_boundary = 1;
    o.crossed = r.cur;
    o.sel = st.sel_exit;
    if (cn.mother < 0) { o.next = -1; return; }
    st.loc_node = cn.mother;
    st.loc_skip = r.cur;
    st.loc_check = 1;
    return;
  }
  st.best = RB_BIG;
  if (st.s_exit < st.best - RB_TOL) st.best = st.s_exit;
  st.mode = 1;
  o.nvis = 0;
  st.bvh_next = cn.bvh_count > 0 ? cn.bvh_first : -1;
}

// fp32 slab test of one ray against padded boxes: one FFMA per plane, FMNMX min/max
struct RayBox {
  float ix, iy, iz, ox, oy, oz, fx, fy, fz, bestf;
  bool px, py, pz;
};
RB_HD inline RayBox raybox_prepare(V3 p, V3 d, double best) {
  RayBox q;
  q.ix = (float)(1. / d.x); q.iy = (float)(1. / d.y); q.iz = (float)(1. / d.z);
  q.ox = -(float)p.x * q.ix; q.oy = -(float)p.y * q.iy; q.oz = -(float)p.z * q.iz;
  q.px = isfinite(q.ix) && isfinite(q.ox); q.py = isfinite(q.iy) && isfinite(q.oy); q.pz = isfinite(q.iz) && isfinite(q.oz);
  q.fx = (float)p.x; q.fy = (float)p.y; q.fz = (float)p.z;
  q.bestf = best > 1e29 ? 3.0e38f : (float)best * 1.000001f + 1e-3f;
  return q;
}
RB_HD inline bool raybox_test(const RayBox& q, const float* lo, const float* hi, float& tmin) {
  tmin = 0.f;
  float tmax = q.bestf;
  bool hit = true;
  if (q.px) { float t0 = fmaf(lo[0], q.ix, q.ox), t1 = fmaf(hi[0], q.ix, q.ox); tmin = fmaxf(tmin, fminf(t0, t1)); tmax = fminf(tmax, fmaxf(t0, t1)); }
  else hit = hit && q.fx >= lo[0] && q.fx <= hi[0];
  if (q.py) { float t0 = fmaf(lo[1], q.iy, q.oy), t1 = fmaf(hi[1], q.iy, q.oy); tmin = fmaxf(tmin, fminf(t0, t1)); tmax = fminf(tmax, fmaxf(t0, t1)); }
  else hit = hit && q.fy >= lo[1] && q.fy <= hi[1];
  if (q.pz) { float t0 = fmaf(lo[2], q.iz, q.oz), t1 = fmaf(hi[2], q.iz, q.oz); tmin = fmaxf(tmin, fminf(t0, t1)); tmax = fminf(tmax, fmaxf(t0, t1)); }
  else hit = hit && q.fz >= lo[2] && q.fz <= hi[2];
  return hit && tmin <= tmax * 1.000002f + 1e-4f;
}
// Boxes that hold the start point all enter the list with parameter 0; among them the one the point sits deepest in goes first
// (key = minus the distance to the nearest face).  That is the cell the ray is flying through — a Winston cone's own wall, a
// facet's own shell — while its neighbours' boxes only reach the point with a corner: once the own cell has given the step,
// the neighbours are dropped after their cheapest operand (Bool2::dist_out's step limit) instead of being evaluated in full.
RB_HD inline float cand_key(const RayBox& q, const float* lo, const float* hi, float tmin) {
  if (tmin > 0.f) return tmin;
  const float depth = fminf(fminf(fminf(q.fx - lo[0], hi[0] - q.fx), fminf(q.fy - lo[1], hi[1] - q.fy)), fminf(q.fz - lo[2], hi[2] - q.fz));
  return depth > 0.f ? -depth : 0.f;
}
// candidate list, nearest box first: the step found in the first candidates lets nb_eval drop the farther ones unevaluated
RB_HD inline void cand_insert(StepOut& o, int first, int child, float tmin) {
  int k = o.nvis++;
  while (k > first && o.tin[k - 1] > tmin) { o.vis[k] = o.vis[k - 1]; o.tin[k] = o.tin[k - 1]; k--; }
  o.vis[k] = child;
  o.tin[k] = tmin;
}

// phase B: threaded-BVH walk over the daughters of the current node; only collects candidates (postponed leaf intersection)
template <class K> RB_HD inline void nb_collect(const DScene& sc, const RayReg& r, NavStep& st) {
  StepOut& o = st.o;
  int i = st.bvh_next;
  if (st.mode != 1 || i < 0) { st.bvh_next = -1; return; }
  const RayBox q = raybox_prepare(r.p, r.d, st.best);
  const int first = o.nvis;
  while (i >= 0 && o.nvis < RB_MAXVIS) {
    const DBvh& b = sc.bvh[i];
    float tmin;
    if (!raybox_test(q, b.lo, b.hi, tmin)) { i = b.skip; continue; }
    if (b.child >= 0) {
      cand_insert(o, first, b.child, cand_key(q, b.lo, b.hi, tmin));
      i = b.skip;
    } else i = i + 1;
  }
  st.bvh_next = i;  // >= 0 only if the candidate buffer filled up
}

// phase C: DistFromOutside of candidate k.  TGeo scans daughters in order and keeps the first one within
// tolerance: on (near-)ties the lowest daughter index wins regardless of the visiting order.
#if defined(RB_EMUL_STATS) && !defined(__CUDACC__)
// host-side analysis build (profiles/cand_stats.py): what happens to the candidates of a step, per step number of the ray
struct RbEvalStats { long long steps[8], listed[8], culled[8], boxed[8], evaluated[8], hits[8]; };
inline RbEvalStats g_eval_stats;
inline int g_eval_step = 0;
#define RB_STAT(field) g_eval_stats.field[g_eval_step < 7 ? g_eval_step : 7]++
#else
#define RB_STAT(field) ((void)0)
#endif
template <class K> RB_HD inline void nb_eval(const DScene& sc, const RayReg& r, NavStep& st, int k) {
  int c = st.o.vis[k];
  RB_STAT(listed);
  // the ray enters this daughter's box (padded by >= 2e-3) beyond the step already found: its DistFromOutside cannot be
  // nearer than the best one, nor tie with it within the tolerance
  if (st.o.tin[k] > (float)st.best * 1.000001f + 1e-3f) { RB_STAT(culled); return; }
  const DNode& dn = sc.nodes[c];
  int sel = 0;
  const V3 lp = to_local(dn.g, r.p), ld = to_local_vec(dn.g, r.d);
  if (local_box_missed(dn, lp, ld)) { RB_STAT(boxed); return; }  // DistFromOutside would report no crossing
  double s = Leaf<K>::dist_out(sc, dn.leaf, dn.shape, lp, ld, st.best + 2 * RB_TOL, sel);
  RB_STAT(evaluated);
  if (s < 1e29) RB_STAT(hits);
  if (s < st.best - RB_TOL || (st.enter >= 0 && c < st.enter && s <= st.best + RB_TOL)) { st.best = s; st.enter = c; st.esel = sel; }
}

// TGeoNavigator::FindNextBoundary with overlapping nodes on the current branch (fNmany > 0): for every node A of the branch
// that was placed with AddNodeOverlap, the boundary of A's mother and the sisters of A are candidates too — an ordinary sister
// from outside, an overlapping sister from inside if it holds the point.  (ROOT restricts the sisters to A's overlap list; a
// sister outside that list cannot be nearer than A's own boundary, so examining all of them gives the same step.)
template <class K> RB_HD inline void nb_many(const DScene& sc, const RayReg& r, NavStep& st) {
  typedef Csg<K::depth, K::shapes> G;
  for (int a = r.cur; a >= 0 && sc.nodes[a].mother >= 0; a = sc.nodes[a].mother) {
    if (!sc.nodes[a].overlap) continue;
    const int y = sc.nodes[a].mother;
    const DNode& yn = sc.nodes[y];
    int sel = 0;
    double s = G::dist_in(sc, yn.shape, to_local(yn.g, r.p), to_local_vec(yn.g, r.d), sel);
    if (s < st.best - RB_TOL) { st.best = s; st.xkind = 1; st.xnode = y; st.xsel = sel; }
    for (int i = yn.bvh_count > 0 ? yn.bvh_first : -1; i >= 0;) {
      const DBvh& b = sc.bvh[i];
      if (b.child < 0) { i = i + 1; continue; }  // every leaf is examined: this path is rare
      const int c = b.child;
      i = b.skip;
      if (c == a) continue;
      const DNode& cn = sc.nodes[c];
      V3 lp = to_local(cn.g, r.p), ld = to_local_vec(cn.g, r.d);
      sel = 0;
      if (cn.overlap && G::contains(sc, cn.shape, lp)) s = G::dist_in(sc, cn.shape, lp, ld, sel);
      else s = G::dist_out(sc, cn.shape, lp, ld, st.best + 2 * RB_TOL, sel);
      if (s < st.best - RB_TOL) { st.best = s; st.xkind = 2; st.xnode = c; st.xsel = sel; }
    }
  }
}

// phase D: move to the boundary (nb_arrive) and locate the node behind it (nb_locate) — CrossBoundaryAndLocate
template <class K> RB_HD inline void nb_arrive(const DScene& sc, RayReg& r, NavStep& st) {
  StepOut& o = st.o;
  if (st.mode == 1) {
    const DNode& cn = sc.nodes[r.cur];
    bool done = false;
    if constexpr ((K::phys & RB_PH_OVERLAP) != 0) {
      if (sc.has_many) {
        nb_many<K>(sc, r, st);
        o.nvis = -1;  // relocate_back must not take its sibling shortcut in a scene with overlapping nodes
        if (st.xkind != 0) {
          r.p = along(r.p, r.d, st.best);
          o.step = st.best + st.extra;
          r.on_boundary = 1;
          o.crossed = st.xnode;
          o.sel = st.xsel;
          const int y = sc.nodes[st.xnode].mother;  // left node: relocate above it; met sister: relocate from the common mother
          if (y < 0) o.next = -1;
          else {
            st.loc_node = y;
            st.loc_skip = st.xkind == 1 ? st.xnode : -1;
            st.loc_check = 1;
            st.loc_prefer = st.xkind == 2 ? st.xnode : -1;
          }
          done = true;
        }
      }
    }
    if (!done) {
      r.p = along(r.p, r.d, st.best);
      o.step = st.best + st.extra;
      r.on_boundary = 1;
      if (st.enter >= 0) {
        o.crossed = st.enter;
        o.sel = st.esel;
        st.loc_node = st.enter;
        if constexpr ((K::phys & RB_PH_OVERLAP) != 0) {
          if (sc.nodes[st.enter].overlap) {  // an overlapping daughter: an ordinary sister holding the point has priority
            st.loc_node = r.cur;
            st.loc_check = 1;
            st.loc_prefer = st.enter;
          }
        }
      } else {
        o.crossed = r.cur;
        o.sel = st.sel_exit;
        if (cn.mother < 0) o.next = -1;
        else {
          st.loc_node = cn.mother;
          st.loc_skip = r.cur;
          st.loc_check = 1;
        }
      }
    }
  }
}
// point location behind the crossed boundary, as requested by nb_begin / nb_arrive (loc_node = -2: nothing to do, o.next is set)
template <class K> RB_HD inline int nb_locate(const DScene& sc, V3 p, V3 d, double step, int loc_node, int loc_skip, int loc_check, int loc_prefer, int next) {
  if (loc_node > -2) return search_node<K>(sc, loc_node, along(p, d, locate_extra(sc, loc_node, step)), loc_skip, loc_check != 0, loc_prefer);
  return next;
}
template <class K> RB_HD inline void nb_finish(const DScene& sc, RayReg& r, NavStep& st) {
  nb_arrive<K>(sc, r, st);
  st.o.next = nb_locate<K>(sc, r.p, r.d, st.o.step, st.loc_node, st.loc_skip, st.loc_check, st.loc_prefer, st.o.next);
}

// the phases run back to back for one thread (per-ray loop kernel, host emulation)
template <class K> RB_HD inline void next_boundary(const DScene& sc, RayReg& r, bool push_quirk, NavStep& st) {
  nb_begin<K>(sc, r, push_quirk, st);
  bool overflow = false;
  while (st.mode == 1 && st.bvh_next >= 0) {
    if (st.o.nvis >= RB_MAXVIS) { st.o.nvis = 0; overflow = true; }  // more than RB_MAXVIS candidates: process in batches
    int first = st.o.nvis;
    nb_collect<K>(sc, r, st);
    for (int k = first; k < st.o.nvis; k++) nb_eval<K>(sc, r, st, k);
  }
  if (overflow) st.o.nvis = -1;  // the visited list is incomplete: relocate_back must not rely on it
  nb_finish<K>(sc, r, st);
}

// ================================================================== physics
RB_HD inline int find_border(const DScene& sc, int vol1, int vol2) {
  if (vol1 < 0) return -1;
  const rbg_volume& v = sc.volumes[vol1];
  for (int i = 0; i < v.nborders; i++)
    if (sc.borders[v.first_border + i].vol2 == vol2) return v.first_border + i;
  return -1;
}

// optional per-ray polyline record (ARay's TGeoTrack points + node history, include/ARay.h:24-68): point k of ray
// `idx` lives at [k * stride + idx]; points beyond max_points are dropped.  Null sink = keep the last point only.
struct DHist {
  double *x, *y, *z, *t;
  int32_t* node;
  long long stride;
  int32_t max_points;
};
struct HistSink {
  const DHist* h;
  long long idx;
};
struct Hit {       // context of one boundary interaction
  int cur_vol, next_vol, next_node, border;
  int crossed, sel;
  double step;
  const StepOut* so;
  const HistSink* hs;
};

template <class K> RB_HD inline V3 geometric_normal(const DScene& sc, const RayReg& r, const Hit& h, V3 dir) {
  if (h.crossed < 0) return v3(0, 0, 1);
  const DNode& nd = sc.nodes[h.crossed];
  V3 ln = Leaf<K>::normal(sc, nd.leaf, nd.shape, to_local(nd.g, r.p), to_local_vec(nd.g, dir), h.sel);
  return to_master_vec(nd.g, ln);
}

// AOpticsManager::GetFacetNormal — geometric normal, optionally perturbed by Gaussian micro-facets
template <class K> RB_HD inline V3 facet_normal(const DScene& sc, const RayReg& r, const Hit& h, Philox& g) {
  V3 n = geometric_normal<K>(sc, r, h, r.d);
  if constexpr ((K::phys & RB_PH_ROUGH) == 0) return n;
  if (h.border < 0) return n;
  const rbg_border& c = sc.borders[h.border];
  if (c.lambertian || c.sigma == 0) return n;
  double sigma = c.sigma, f_max = rb_min(1., 4. * sigma);
  V3 fn;
  for (int guard = 0; guard < 100000; guard++) {
    double alpha;
    do {
      alpha = rng_gaus(g, 0, sigma);
    } while (f_max * rng_uniform(g) > sin(alpha) || alpha >= RB_PI / 2);
    double phi = 2 * RB_PI * rng_uniform(g);
    double sa = sin(alpha), ca = cos(alpha), px = sa * cos(phi), py = sa * sin(phi), pz = ca;
    double up = n.x * n.x + n.y * n.y;  // TVector3::RotateUz(n)
    if (up != 0) {
      up = sqrt(up);
      fn = v3((n.x * n.z * px - n.y * py + n.x * up * pz) / up, (n.y * n.z * px + n.x * py + n.y * up * pz) / up, (n.z * n.z * px - px + n.z * up * pz) / up);
    } else if (n.z < 0.) fn = v3(-px, py, -pz);
    else fn = v3(px, py, pz);
    if (dot(r.d, fn) > 0.0) break;
  }
  return fn;
}

RB_HD inline double mirror_reflectance(const DScene& sc, int vol, double lambda, double angle) {
  int mi = sc.volumes[vol].mirror;
  double ret = 1.0;
  if (mi >= 0) {
    const rbg_mirror& m = sc.mirrors[mi];
    if (m.graph2d >= 0) ret = graph2d_interp(sc, m.graph2d, lambda, angle);
    else if (m.th2 >= 0) ret = th2_interp(sc, m.th2, lambda, angle);
    else if (m.graph1d >= 0) ret = graph_eval(sc, m.graph1d, lambda);
    else ret = m.constant;
  }
  return ret > 1 ? 1 : (ret < 0 ? 0 : ret);
}

// ARay::AddPoint + AddNode
RB_HD inline void add_point(RayReg& r, V3 p, double t, int node, const HistSink* hs) {
  r.p = p;
  r.t = t;
  r.last_node = node;
  if (hs && r.npoints < hs->h->max_points) {
    long long o = (long long)r.npoints * hs->h->stride + hs->idx;
    hs->h->x[o] = p.x; hs->h->y[o] = p.y; hs->h->z[o] = p.z; hs->h->t[o] = t;
    hs->h->node[o] = node;
  }
  r.npoints++;
}
RB_HD inline void set_direction(RayReg& r, V3 d2) {
  double mag = sqrt(dot(d2, d2));
  if (mag > 0) r.d = (1. / mag) * d2;
}

// Navigator state after DoReflection's backward Step(): TGeoNavigator::FindNode() from the entered node at
// `back` = boundary - 2e-6 d1.  Common case (a daughter of the start node was entered): `back` lies on the ray
// just before the boundary, so any sibling containing it had its AABB touched by this step's traversal and is in
// so.vis — no second BVH walk is needed.  Everything else takes the generic SearchNode path.
template <class K> RB_HD inline int relocate_back(const DScene& sc, const Hit& h, V3 back) {
  const StepOut& so = *h.so;
  int start = h.next_node < 0 ? 0 : h.next_node;
  bool check = true;
  if (so.nvis >= 0 && so.from >= 0 && so.next >= 0 && so.next == so.crossed && sc.nodes[so.next].mother == so.from && so.step > 4e-6) {
    if (node_contains<K>(sc, so.next, back)) { start = so.next; check = false; }
    else if (node_contains<K>(sc, so.from, back)) {
      int best = -1;
      for (int k = 0; k < so.nvis; k++) {
        int c = so.vis[k];
        if (c == so.next || (best >= 0 && c > best)) continue;
        if (!node_box_holds(sc.nodes[c], back)) continue;  // a point outside the (padded) box is outside the shape
        if (node_contains<K>(sc, c, back)) best = c;
      }
      if (best < 0) return so.from;
      start = best;
      check = false;
    }
  }
  return search_node<K>(sc, start, back, -1, check);
}

// AOpticsManager::DoReflection.  `pos` is the boundary point; r.p still holds the segment start.
// (`n` is the facet normal, drawn by the caller: trace_shade evaluates GetFacetNormal at one place for all its users)
template <class K> RB_HD inline void do_reflection(const DScene& sc, const DTraceParams& tp, RayReg& r, V3& pos, int& loc, const Hit& h, double n1,
                                                    Philox& g, V3 n) {
  V3 d1 = r.d;
  double cos1 = dot(d1, n);
  bool absorbed = false;
  int next_type = h.next_vol < 0 ? RBG_NULL : sc.volumes[h.next_vol].type;
  const rbg_border* c = h.border >= 0 ? &sc.borders[h.border] : nullptr;
  if (next_type == RBG_MIRROR) {
    double ref = 1.0;
    bool have = false;
    if constexpr ((K::phys & RB_PH_MULTILAYER) != 0) {
      if (c && c->multilayer >= 0) {
        double tr;
        tmm_mixed(sc, c->multilayer, rb_acos(cos1), r.lambda, ref, tr);
        have = true;
      }
    }
    if (!have) {
      if constexpr ((K::phys & RB_PH_MIRROR_TABLE) != 0) ref = mirror_reflectance(sc, h.next_vol, r.lambda, rb_acos(cos1));
      else {
        int mi = sc.volumes[h.next_vol].mirror;
        ref = mi >= 0 ? sc.mirrors[mi].constant : 1.0;
        ref = ref > 1 ? 1 : (ref < 0 ? 0 : ref);
      }
    }
    if (ref < rng_uniform(g)) { absorbed = true; r.status = RBG_ABSORB; }  // the reference always draws here
  }
  V3 d2;
  bool lamb = false;
  if constexpr ((K::phys & RB_PH_LAMBERT) != 0) lamb = c && c->lambertian;
  if (lamb) {
    double y = 0.5 * rng_uniform(g), theta = rb_asin(sqrt(2 * y)), phi = 2 * RB_PI * rng_uniform(g);
    double perp = sqrt(n.x * n.x + n.y * n.y);
    double theta_n = (n.x == 0 && n.y == 0 && n.z == 0 ? 0 : atan2(perp, n.z)) * 180. / RB_PI;
    double phi_n = (n.x == 0 && n.y == 0 ? 0 : atan2(n.y, n.x)) * 180. / RB_PI;
    double ph = (phi_n + 90) * RB_PI / 180., th = (theta_n + 180) * RB_PI / 180.;
    double sp = sin(ph), cp = cos(ph), st = sin(th), ct = cos(th);
    V3 v = v3(sin(theta) * cos(phi), sin(theta) * sin(phi), cos(theta));
    d2 = v3(cp * v.x - ct * sp * v.y + st * sp * v.z, sp * v.x + ct * cp * v.y - st * cp * v.z, st * v.y + ct * v.z);
  } else d2 = v3(d1.x - 2 * n.x * cos1, d1.y - 2 * n.y * cos1, d1.z - 2 * n.z * cos1);
  if (!absorbed) set_direction(r, d2);
  double t = r.t + h.step / (RB_C_CM / n1);
  // nav->Step() backwards by 1e-6 (+1e-6 of ROOT's Step) and relocate; the recorded vertex is the stepped-back point
  V3 back = along(pos, d1, -2e-6);
  loc = relocate_back<K>(sc, h, back);
  if (tp.quirks & RBG_QUIRK_STEPBACK) pos = back;
  add_point(r, pos, t, h.next_node, h.hs);
  r.on_boundary = 0;
}

// AOpticsManager::DoFresnel.  Returns true when the ray is to be reflected instead (multilayer reflection, total internal
// reflection, Fresnel reflection: the reference calls DoReflection from here, :91, :100, :138) — the caller does that.
template <class K> RB_HD inline bool do_fresnel(const DScene& sc, const DTraceParams& tp, RayReg& r, V3& pos, const Hit& h, double n1, double n2,
                                                 double k2, Philox& g, V3 n) {
  V3 d1 = r.d;
  double cos1 = dot(d1, n), sin1 = sqrt(1 - cos1 * cos1), sin2 = n1 * sin1 / n2, cos2 = sqrt(1 - sin2 * sin2);
  bool absorbed = false, decided = false;
  const rbg_border* c = h.border >= 0 ? &sc.borders[h.border] : nullptr;
  if constexpr ((K::phys & RB_PH_MULTILAYER) != 0) {
    if (c && c->multilayer >= 0) {
      double R, T;
      tmm_mixed(sc, c->multilayer, rb_acos(cos1), r.lambda, R, T);
      double rnd = rng_uniform(g);
      if (rnd < R) return true;
      decided = true;
      if (!(rnd < R + T)) absorbed = true;
    }
  }
  if (!decided) {
    if (sin2 > 1.) return true;
    if (!tp.disable_fresnel) {
      double Rs, Rp;
      if (k2 <= 0.) {
        double e1s = n1 * cos1, e2s = n2 * cos2, e1p = n1 / cos1, e2p = n2 / cos2;
        Rs = sqr((e1s - e2s) / (e1s + e2s));
        Rp = sqr((e1p - e2p) / (e1p + e2p));
      } else {
        double x1S = n1 * cos1, x1P = n1 / cos1;
        double u = sqr(n2) - sqr(k2) - sqr(n1 * sin1), v = 2 * n2 * k2, tmp = sqrt(sqr(u) + sqr(v));
        double cosxi2 = sqrt(1 + u / tmp) / sqrt(2.), sinxi2 = sqrt(1 - u / tmp) / sqrt(2.);
        double x2S = sqrt(tmp) * cosxi2, y2S = sqrt(tmp) * sinxi2;
        tmp = sqr(x2S) + sqr(y2S);
        double x2P = (2 * n2 * k2 * y2S + (sqr(n2) - sqr(k2)) * x2S) / tmp, y2P = (2 * n2 * k2 * x2S - (sqr(n2) - sqr(k2)) * y2S) / tmp;
        Rs = (sqr(x1S - x2S) + sqr(y2S)) / (sqr(x1S + x2S) + sqr(y2S));
        Rp = (sqr(x1P - x2P) + sqr(y2P)) / (sqr(x1P + x2P) + sqr(y2P));
      }
      if (rng_uniform(g) < (Rs + Rp) / 2.) return true;
    }
  }
  V3 d2 = d1;
  if (sin1 != 0) {
    double f = sin2 / sin1;
    d2 = v3((d1.x - cos1 * n.x) * f + n.x * cos2, (d1.y - cos1 * n.y) * f + n.y * cos2, (d1.z - cos1 * n.z) * f + n.z * cos2);
  }
  add_point(r, pos, r.t + h.step / (RB_C_CM / n1), h.next_node, h.hs);
  if (absorbed) r.status = RBG_ABSORB;
  else set_direction(r, d2);
  return false;
}

// One iteration of the while(ray->IsRunning()) loop, src/AOpticsManager.cxx:359-518
// (the interaction / termination half; `nav` is the navigator copy whose point sits on the boundary, `so` the step record)
template <class K>
RB_HD inline void trace_shade(const DScene& sc, const DTraceParams& tp, RayReg& r, const RayReg& nav, const StepOut& so, Philox& g, const HistSink* hs = nullptr) {
  V3 x1 = r.p;
  double t1 = r.t;
  int cur = r.cur;
  int cur_vol = cur < 0 ? -1 : sc.nodes[cur].volume;
  int typeCurrent = cur < 0 ? RBG_NULL : sc.nodes[cur].type;
  r.on_boundary = nav.on_boundary;
  Hit h;
  h.cur_vol = cur_vol;
  h.next_node = so.next;
  h.next_vol = so.next < 0 ? -1 : sc.nodes[so.next].volume;
  h.crossed = so.crossed;
  h.sel = so.sel;
  h.step = so.step;
  h.so = &so;
  h.hs = hs;
  h.border = find_border(sc, cur_vol, h.next_vol);
  int typeNext = so.next < 0 ? RBG_NULL : sc.nodes[so.next].type;
  V3 pos = nav.p;
  int loc = so.next;
  double lambda = r.lambda;
  constexpr bool kLens = (K::phys & RB_PH_LENS) != 0;  // without ALens volumes every lens branch is dead code
  if (kLens && typeCurrent == RBG_LENS) {
    double k = index_k(sc, sc.volumes[cur_vol].index, lambda);
    if (k > 0) {
      double abs = lambda / (4 * RB_PI * k);
      if (abs > 0 && abs < 1e300) {
        double abs_step = -abs * log(rng_uniform(g));
        if (abs_step < so.step) {
          double n1 = index_n(sc, sc.volumes[cur_vol].index, lambda);
          add_point(r, along(x1, r.d, abs_step), t1 + abs_step / (RB_C_CM / n1), so.next, hs);
          r.status = RBG_ABSORB;
          r.cur = loc;
          return;
        }
      }
    }
  }
  bool curVac = typeCurrent == RBG_NULL || typeCurrent == RBG_OPT || typeCurrent == RBG_OTHER;
  bool curLens = kLens && typeCurrent == RBG_LENS;
  const bool nextVac = typeNext == RBG_NULL || typeNext == RBG_OPT || typeNext == RBG_OTHER;
  bool reflect = (curVac || curLens) && typeNext == RBG_MIRROR;
  const bool fresnel = kLens && ((curVac && typeNext == RBG_LENS) || (curLens && (typeNext == RBG_LENS || nextVac)));
  // QE(theta) of a focal surface needs the facet normal as well (src/AOpticsManager.cxx:495-505); none of the branches leading
  // there draws a random number before it, so GetFacetNormal can be evaluated here, once, for whichever branch wants it
  bool qe_angle = false;
  if constexpr ((K::phys & RB_PH_QE) != 0) {
    if (typeNext == RBG_FOCUS && !(typeCurrent == RBG_FOCUS || typeCurrent == RBG_OBS || typeCurrent == RBG_MIRROR)) {
      const int fo = sc.volumes[h.next_vol].focal;
      qe_angle = fo >= 0 && sc.focals[fo].qe_angle >= 0;
    }
  }
  V3 n = v3(0, 0, 1);
  if (reflect || fresnel || qe_angle) {
    RayReg at = r;
    at.p = pos;
    n = facet_normal<K>(sc, at, h, g);
  }
  double n1 = 1.;
  if constexpr (kLens) n1 = curLens ? index_n(sc, sc.volumes[cur_vol].index, lambda) : 1.;
  if (fresnel) {
    if constexpr (kLens) {
      double n2 = 1., k2 = 0.;
      if (typeNext == RBG_LENS) {
        const int ix = sc.volumes[h.next_vol].index;
        n2 = index_n(sc, ix, lambda);
        k2 = index_k(sc, ix, lambda);
      }
      reflect = do_fresnel<K>(sc, tp, r, pos, h, n1, n2, k2, g, n);
    }
  } else if (!reflect) {
    if ((curVac || curLens) && (typeNext == RBG_OBS || typeNext == RBG_FOCUS)) add_point(r, pos, t1 + so.step / (RB_C_CM / n1), so.next, hs);
    else if (curVac && (typeNext == RBG_OTHER || typeNext == RBG_OPT)) add_point(r, pos, t1 + so.step / RB_C_CM, so.next, hs);
  }
  if (reflect) do_reflection<K>(sc, tp, r, pos, loc, h, n1, g, n);
  // termination (evaluated after the interaction, src/AOpticsManager.cxx:485-513)
  if (typeNext == RBG_NULL) {
    add_point(r, pos, t1 + so.step / RB_C_CM, so.next, hs);
    r.status = RBG_EXIT;
  } else if (typeCurrent == RBG_FOCUS || typeCurrent == RBG_OBS || typeCurrent == RBG_MIRROR || typeNext == RBG_OBS) {
    r.status = RBG_STOP;
  } else if (typeNext == RBG_FOCUS) {
    double qe = 1.;
    if constexpr ((K::phys & RB_PH_QE) != 0) {
      const rbg_volume& fv = sc.volumes[h.next_vol];
      if (fv.focal >= 0) {
        const rbg_focal f = sc.focals[fv.focal];
        if (f.qe_lambda >= 0) qe = graph_eval(sc, f.qe_lambda, lambda);
        if (f.qe_angle >= 0) qe *= graph_eval(sc, f.qe_angle, rb_acos(dot(r.d, n)));
      }
    }
    if (qe == 1 || rng_uniform(g) < qe) r.status = RBG_FOCUSED;
    else r.status = RBG_STOP;
  }
  r.cur = loc;
  if (r.status == RBG_RUN && r.npoints >= tp.limit) r.status = RBG_SUSPEND;
}

template <class K> RB_HD inline void trace_step(const DScene& sc, const DTraceParams& tp, RayReg& r, Philox& g, const HistSink* hs = nullptr) {
  RayReg nav = r;  // navigator copy: nav.p advances to the boundary, r.p stays at the segment start
  NavStep st;
  next_boundary<K>(sc, nav, (tp.quirks & RBG_QUIRK_BOUNDARY_PUSH) != 0, st);
  trace_shade<K>(sc, tp, r, nav, st.o, g, hs);
}

// locate the start point (InitTrack -> FindNode)
template <class K> RB_HD inline int locate_start(const DScene& sc, V3 p) { return search_node<K>(sc, 0, p, -1, true); }

#endif  // RB_DEVICE_CUH
