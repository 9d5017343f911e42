// oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY: nothing under robast_b200/ (the product)
// may include, link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker or as the reported CPU baseline.
//
// What it is: a plain, one-ray-at-a-time fp64 restatement of the reference algorithm for
//   AOpticsManager::TraceNonSequential            /root/reference/src/AOpticsManager.cxx:52-587
// with a hierarchical TGeoNavigator-like state (path stack + per-level global matrices), the
// reference's own shapes (src/AGeoAsphericDisk.cxx, src/AGeoWinstonCone2D.cxx,
// src/AGeoWinstonConePoly.cxx), its optics (src/AMultilayer.cxx, src/A*Formula.cxx, src/AMirror.cxx,
// src/ALens.cxx, src/AFocalSurface.cxx, src/AOpticalComponent.cxx) and the ROOT behaviour it relies on.
//
// PARITY PINNING.  ROOT (libGeom: TGeoNavigator, TGeoBBox/Tube/Sphere/Paraboloid/Pgon/Pcon/
// CompositeShape, TGraph, TH2) is an un-vendored, un-pinned third-party dependency that is absent
// from /root/reference and from this environment.  Its published algorithms are
// restated here from the ROOT 6 sources as recalled.  Pinned against the
// reference's own golden vectors: TMM (tutorials/unittest_robast.py:629-641), Sellmeier / AGF N-BK7
// (:524-561), TGraph interpolation (:415-426), Snell (:428-468), limit (:390-413), Fresnel n=3 (:157).
// Per-ray positions THROUGH TGeo shapes are "parity unpinned" by the reference's tests; they are pinned
// here by closed-form optics (tests/test_oracle_closed_form.py).
//
// Random numbers: the reference uses the global gRandom (TRandom3), shared and unlocked across
// threads; stochastic parity with it can only be statistical.  The oracle uses Philox4x32-10 keyed by
// (seed, global ray id, draw index) — the same stream layout as the CUDA path — so that stochastic
// branches can also be compared per ray against the GPU.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <thread>
#include <vector>

#include "../include/robast_b200.h"

namespace {

const double kBig = 1e30;     // TGeoShape::Big()
const double kTol = 1e-10;    // TGeoShape::Tolerance()
const double kPi = 3.14159265358979323846;
const double kInf = std::numeric_limits<double>::infinity();
const double kEpsilon = 1e-6;  // src/AOpticsManager.cxx:20
const double kC = 2.99792458e8 * 100.;  // TMath::C()*m()  [cm/s]

inline double sq(double v) { return v * v; }
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double ATan2(double y, double x) {  // TMath::ATan2
  if (x != 0) return atan2(y, x);
  if (y == 0) return 0;
  return y > 0 ? kPi / 2 : -kPi / 2;
}
inline double ACosT(double x) { return x < -1. ? kPi : (x > 1. ? 0 : acos(x)); }
inline double ASinT(double x) { return x < -1. ? -kPi / 2 : (x > 1. ? kPi / 2 : asin(x)); }

// ------------------------------------------------------------------ Philox4x32-10
struct Rng {
  uint32_t key[2];
  uint32_t id[2];
  uint32_t ndraw;
  void block(uint32_t out[4]) {
    uint32_t c[4] = {id[0], id[1], ndraw++, 0u}, k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
      uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
      uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
      c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
  }
  static double u53(uint32_t a, uint32_t b) { return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) / 9007199254740992.0; }
  double uniform() {  // (0,1)
    uint32_t o[4];
    block(o);
    return u53(o[0], o[1]);
  }
  double gaus(double mean, double sigma) {
    uint32_t o[4];
    block(o);
    double u1 = u53(o[0], o[1]), u2 = u53(o[2], o[3]);
    return mean + sigma * sqrt(-2. * log(u1)) * cos(2 * kPi * u2);
  }
};

// ------------------------------------------------------------------ matrices
struct Mat {
  double r[9], t[3];
};
const Mat kIdentity = {{1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 0, 0}};
inline void l2m(const Mat& m, const double* l, double* o) {
  for (int i = 0; i < 3; i++) o[i] = m.t[i] + l[0] * m.r[3 * i] + l[1] * m.r[3 * i + 1] + l[2] * m.r[3 * i + 2];
}
inline void l2mv(const Mat& m, const double* l, double* o) {
  for (int i = 0; i < 3; i++) o[i] = l[0] * m.r[3 * i] + l[1] * m.r[3 * i + 1] + l[2] * m.r[3 * i + 2];
}
inline void m2l(const Mat& m, const double* p, double* o) {
  double a = p[0] - m.t[0], b = p[1] - m.t[1], c = p[2] - m.t[2];
  for (int i = 0; i < 3; i++) o[i] = a * m.r[i] + b * m.r[i + 3] + c * m.r[i + 6];
}
inline void m2lv(const Mat& m, const double* p, double* o) {
  for (int i = 0; i < 3; i++) o[i] = p[0] * m.r[i] + p[1] * m.r[i + 3] + p[2] * m.r[i + 6];
}
inline Mat mul(const Mat& a, const Mat& b) {  // a*b: apply b first
  Mat o;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) o.r[3 * i + j] = a.r[3 * i] * b.r[j] + a.r[3 * i + 1] * b.r[3 + j] + a.r[3 * i + 2] * b.r[6 + j];
    o.t[i] = a.t[i] + a.r[3 * i] * b.t[0] + a.r[3 * i + 1] * b.t[1] + a.r[3 * i + 2] * b.t[2];
  }
  return o;
}

// Daughter look-up acceleration, the role TGeoVoxelFinder plays for TGeoNavigator (ROOT voxelises every volume with daughters at
// CloseGeometry; the brute-force walk over all daughters is what TGeo does only for tiny volumes).  Per volume: the bounding box
// of every daughter in the mother's frame and a binary tree over them.  It is a pure culling structure: a daughter is skipped
// only if the ray cannot reach its (padded) box before the current step / the point lies outside the box, i.e. exactly when
// its DistFromOutside / Contains could not have changed the outcome, and the survivors are evaluated in AddNode order with
// the same running step as the full walk — the results are bit-identical with and without it (tests/test_oracle_golden.py).
struct VoxNode {
  double lo[3], hi[3];
  int left, right;   // children, or -1
  int first, count;  // leaf: slice of VolVox::order
};
struct VolVox {
  std::vector<double> box;  // 6 per daughter: lo[3], hi[3] in the mother's frame (padded)
  std::vector<int> order;   // daughter indices (0 .. nnodes-1) as the leaves hold them
  std::vector<VoxNode> tree;
};
static int g_use_voxels = 1;

struct Scene {
  const rbg_scene_desc* d;
  std::vector<int> subtree;  // physical nodes in the subtree of each volume (incl. itself)
  std::vector<VolVox> vox;   // per volume; empty tree = walk all daughters
  Mat mat(int id) const {
    if (id < 0) return kIdentity;
    Mat m;
    memcpy(m.r, d->matrices[id].rot, sizeof(m.r));
    memcpy(m.t, d->matrices[id].tr, sizeof(m.t));
    return m;
  }
  int count(int vol) {
    if (subtree[vol] >= 0) return subtree[vol];
    int c = 1;
    const rbg_volume& v = d->volumes[vol];
    for (int k = 0; k < v.nnodes; k++) c += count(d->nodes[v.first_node + k].volume);
    return subtree[vol] = c;
  }
  explicit Scene(const rbg_scene_desc* desc) : d(desc), subtree(desc->nvolumes, -1) {
    for (int i = 0; i < desc->nvolumes; i++) count(i);
  }
};

// ================================================================== 1-D / 2-D tables
// TGraph::Eval (no spline), ROOT 6: linear interpolation, linear extrapolation
double graph_eval(const rbg_scene_desc* d, int g, double x) {
  const rbg_graph& gr = d->graphs[g];
  const double *X = d->gx + gr.first, *Y = d->gy + gr.first;
  int n = gr.n;
  if (n == 0) return 0;
  if (n == 1) return Y[0];
  int low = -1, up = -1, low2 = -1, up2 = -1;
  for (int i = 0; i < n; ++i) {
    if (X[i] < x) {
      if (low == -1 || X[i] > X[low]) { low2 = low; low = i; }
      else if (low2 == -1 || X[i] > X[low2]) low2 = i;
    } else if (X[i] > x) {
      if (up == -1 || X[i] < X[up]) { up2 = up; up = i; }
      else if (up2 == -1 || X[i] < X[up2]) up2 = i;
    } else return Y[i];
  }
  if (up == -1) { up = low; low = low2; }
  if (low == -1) { low = up; up = up2; }
  if (X[low] == X[up]) return Y[low];
  return Y[up] + (x - X[up]) * (Y[low] - Y[up]) / (X[low] - X[up]);
}

// TH2::Interpolate: bilinear between the 4 surrounding bin centres
double th2_interp(const rbg_scene_desc* d, int h, double x, double y) {
  const rbg_th2& H = d->th2[h];
  const double* v = d->th2v + H.first;
  double wx = (H.xmax - H.xmin) / H.nx, wy = (H.ymax - H.ymin) / H.ny;
  auto findbin = [](double x, double lo, double hi, int n) { return x < lo ? 0 : (!(x < hi) ? n + 1 : 1 + int(n * (x - lo) / (hi - lo))); };
  int bx = findbin(x, H.xmin, H.xmax, H.nx), by = findbin(y, H.ymin, H.ymax, H.ny);
  if (bx < 1 || bx > H.nx || by < 1 || by > H.ny) return 0;  // ROOT prints an error and returns 0
  double dx = (H.xmin + bx * wx) - x, dy = (H.ymin + by * wy) - y;
  int ix1 = dx <= wx / 2 ? bx : bx - 1, iy1 = dy <= wy / 2 ? by : by - 1;
  double x1 = H.xmin + (ix1 - 0.5) * wx, x2 = H.xmin + (ix1 + 0.5) * wx, y1 = H.ymin + (iy1 - 0.5) * wy, y2 = H.ymin + (iy1 + 0.5) * wy;
  int bx1 = std::max(ix1, 1), bx2 = std::min(ix1 + 1, H.nx), by1 = std::max(iy1, 1), by2 = std::min(iy1 + 1, H.ny);
  auto C = [&](int i, int j) { return v[(i - 1) + H.nx * (j - 1)]; };
  double q11 = C(bx1, by1), q12 = C(bx1, by2), q21 = C(bx2, by1), q22 = C(bx2, by2), dd = 1.0 * (x2 - x1) * (y2 - y1);
  return 1.0 * q11 / dd * (x2 - x) * (y2 - y) + 1.0 * q21 / dd * (x - x1) * (y2 - y) + 1.0 * q12 / dd * (x2 - x) * (y - y1) +
         1.0 * q22 / dd * (x - x1) * (y - y1);
}

// TGraph2D::Interpolate (src/AMirror.cxx:47): ROOT interpolates linearly inside the Delaunay triangle that holds
// (x, y) and returns 0 outside the convex hull.  The triangle list comes with the scene description (the host layer
// triangulates, include/robast/RootCompat.h); the plane through the three vertices is evaluated here in the
// point-normal form z = z0 - (nx (x-x0) + ny (y-y0)) / nz, after a same-side test against each edge.
double graph2d_interp(const rbg_scene_desc* d, int g, double x, double y) {
  const rbg_graph2d& G = d->graph2d[g];
  for (int k = 0; k < G.ntri; k++) {
    const int32_t* t = d->tri + 3 * (G.first_tri + k);
    double X[3], Y[3], Z[3];
    for (int i = 0; i < 3; i++) { X[i] = d->g2x[t[i]]; Y[i] = d->g2y[t[i]]; Z[i] = d->g2z[t[i]]; }
    double area2 = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0]);
    if (area2 == 0) continue;
    bool inside = true;
    for (int i = 0; i < 3 && inside; i++) {  // signed area of (edge i, point) relative to the triangle's own orientation
      int a = (i + 1) % 3, b = (i + 2) % 3;
      double w = ((X[a] - x) * (Y[b] - y) - (X[b] - x) * (Y[a] - y)) / area2;
      if (w < -1e-9) inside = false;
    }
    if (!inside) continue;
    double ux = X[1] - X[0], uy = Y[1] - Y[0], uz = Z[1] - Z[0], vx = X[2] - X[0], vy = Y[2] - Y[0], vz = Z[2] - Z[0];
    double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
    return Z[0] - (nx * (x - X[0]) + ny * (y - Y[0])) / nz;
  }
  return 0.;
}

// ================================================================== refractive indices
// include/ARefractiveIndex.h:36-65, src/ASellmeierFormula.cxx:46-54, src/ASchottFormula.cxx:43-55,
// src/ACauchyFormula.cxx:40-46, include/AMixedRefractiveIndex.h:36-45
double index_n(const rbg_scene_desc* d, int id, double lambda);
double index_k(const rbg_scene_desc* d, int id, double lambda) {
  if (id < 0) return 0.;
  const rbg_index& x = d->indices[id];
  if (x.kind == RBG_INDEX_MIXED) return index_k(d, x.mix_a, lambda) * x.frac_a + index_k(d, x.mix_b, lambda) * x.frac_b;
  return x.kgraph >= 0 ? graph_eval(d, x.kgraph, lambda) : 0.;
}
double index_n(const rbg_scene_desc* d, int id, double lambda) {
  if (id < 0) return 1.;
  const rbg_index& x = d->indices[id];
  const double* p = x.par;
  double l = lambda / 1e-4;  // cm -> um  (AOpticsManager::um())
  switch (x.kind) {
    case RBG_INDEX_SELLMEIER: {
      double l2 = l * l;
      return sqrt(1 + p[0] * l2 / (l2 - p[3]) + p[1] * l2 / (l2 - p[4]) + p[2] * l2 / (l2 - p[5]));
    }
    case RBG_INDEX_SCHOTT:
      return sqrt(p[0] + p[1] * pow(l, 2.) + p[2] * pow(l, -2.) + p[3] * pow(l, -4.) + p[4] * pow(l, -6.) + p[5] * pow(l, -8.));
    case RBG_INDEX_CAUCHY:
      return p[0] + p[1] * pow(l, -2) + p[2] * pow(l, -4);
    case RBG_INDEX_MIXED:
      return index_n(d, x.mix_a, lambda) * x.frac_a + index_n(d, x.mix_b, lambda) * x.frac_b;
    default:
      return x.ngraph >= 0 ? graph_eval(d, x.ngraph, lambda) : 1.;
  }
}
double index_abslen(const rbg_scene_desc* d, int id, double lambda) {
  double k = index_k(d, id, lambda);
  return k <= 0. ? kInf : lambda / (4 * kPi * k);
}

// ================================================================== AMultilayer::CoherentTMM
// src/AMultilayer.cxx:26-56 (interface_rt), :120-176 (IsForwardAngle), :178-209 (ListSnell), :240-481
typedef std::complex<double> cplx;
bool is_forward_angle(cplx n, cplx theta) {
  cplx ncostheta = n * std::cos(theta);
  const double EPS = std::numeric_limits<double>::epsilon();
  if (std::abs(ncostheta.imag()) > 100 * EPS) return ncostheta.imag() > 0;
  return ncostheta.real() > 0;
}
// CoherentTMM on explicit lists (the reference reverses copies of its lists for `reverse`, :268-277, and builds a
// temporary AMultilayer per coherent stack inside IncoherentTMM, :612-624)
void coherent_tmm_lists(const std::vector<cplx>& n_list, const std::vector<double>& thick, int pol /*0=S,1=P*/, cplx th_0, double lam, double& R, double& T) {
  int N = (int)n_list.size();
  std::vector<cplx> th_list(N), kz(N), cos_th(N), delta(N), t_list(N), r_list(N);
  cplx n0_sinth0 = n_list[0] * std::sin(th_0);
  for (int i = 0; i < N; i++) th_list[i] = std::asin(n0_sinth0 / n_list[i]);
  if (!is_forward_angle(n_list[0], th_list[0])) th_list[0] = kPi - th_list[0];
  if (!is_forward_angle(n_list[N - 1], th_list[N - 1])) th_list[N - 1] = kPi - th_list[N - 1];
  for (int i = 0; i < N; i++) {
    cos_th[i] = std::cos(th_list[i]);
    kz[i] = 2 * kPi * n_list[i] * cos_th[i] / lam;
    delta[i] = kz[i] * thick[i];
  }
  for (int i = 1; i < N - 1; i++)
    if (delta[i].imag() > 35) delta[i] = delta[i].real() + cplx(0, 35);
  for (int i = 0; i < N - 1; i++) {
    cplx n_i = n_list[i], n_f = n_list[i + 1], th_i = th_list[i], th_f = th_list[i + 1];
    cplx ii = n_i * std::cos(th_i);
    if (pol == 0) {
      cplx ff = n_f * std::cos(th_f);
      r_list[i] = (ii - ff) / (ii + ff);
      t_list[i] = 2. * ii / (ii + ff);
    } else {
      cplx fi = n_f * std::cos(th_i), i_f = n_i * std::cos(th_f);
      r_list[i] = (fi - i_f) / (fi + i_f);
      t_list[i] = 2. * ii / (fi + i_f);
    }
  }
  // Mtilde = prod_{i=1}^{N-2} (1/t_i) diag(e^{-i delta}, e^{i delta}) [[1,r_i],[r_i,1]]
  cplx m00(1, 0), m01(0, 0), m10(0, 0), m11(1, 0);
  const cplx j(0, 1);
  for (int i = 1; i < N - 1; i++) {
    cplx em = std::exp(-j * delta[i]), ep = std::exp(j * delta[i]);
    // reference evaluates (1./t * D) * B with D = diag(e^{-i delta}, e^{i delta}), B = [[1,r],[r,1]]
    cplx s = 1. / t_list[i];
    cplx d00 = s * em, d11 = s * ep;
    cplx a00 = d00, a01 = d00 * r_list[i], a10 = d11 * r_list[i], a11 = d11;
    cplx n00 = m00 * a00 + m01 * a10, n01 = m00 * a01 + m01 * a11, n10 = m10 * a00 + m11 * a10, n11 = m10 * a01 + m11 * a11;
    m00 = n00; m01 = n01; m10 = n10; m11 = n11;
  }
  {
    cplx b00 = cplx(1, 0) / t_list[0], b01 = r_list[0] / t_list[0], b10 = r_list[0] / t_list[0], b11 = cplx(1, 0) / t_list[0];
    cplx n00 = b00 * m00 + b01 * m10, n01 = b00 * m01 + b01 * m11, n10 = b10 * m00 + b11 * m10, n11 = b10 * m01 + b11 * m11;
    m00 = n00; m01 = n01; m10 = n10; m11 = n11;
  }
  cplx r = m10 / m00, t = 1. / m00;
  R = std::abs(r) * std::abs(r);
  cplx n_i = n_list[0], n_f = n_list[N - 1], th_i = th_0, th_f = th_list[N - 1];
  if (pol == 0) T = std::abs(t * t) * (((n_f * std::cos(th_f)).real()) / (n_i * std::cos(th_i)).real());
  else T = std::abs(t * t) * (((n_f * std::conj(std::cos(th_f))).real()) / (n_i * std::conj(std::cos(th_i))).real());
}
void multilayer_lists(const rbg_scene_desc* d, int ml, double lam, std::vector<cplx>& n_list, std::vector<double>& thick, std::vector<bool>& coherent) {
  const rbg_multilayer& M = d->multilayers[ml];
  n_list.resize(M.n); thick.resize(M.n); coherent.resize(M.n);
  for (int i = 0; i < M.n; i++) {
    const rbg_layer& L = d->layers[M.first + i];
    n_list[i] = cplx(index_n(d, L.index, lam), index_k(d, L.index, lam));
    thick[i] = L.thickness;
    coherent[i] = !(i == 0 || i == M.n - 1 || L.incoherent);
  }
}
void coherent_tmm_cplx(const rbg_scene_desc* d, int ml, int pol, cplx th_0, double lam, double& R, double& T, bool reverse) {
  std::vector<cplx> n_list;
  std::vector<double> thick;
  std::vector<bool> coh;
  multilayer_lists(d, ml, lam, n_list, thick, coh);
  if (reverse) {
    std::reverse(n_list.begin(), n_list.end());
    std::reverse(thick.begin(), thick.end());
  }
  coherent_tmm_lists(n_list, thick, pol, th_0, lam, R, T);
}
void coherent_tmm(const rbg_scene_desc* d, int ml, int pol /*0=S,1=P*/, double th0r, double lam, double& R, double& T) {
  coherent_tmm_cplx(d, ml, pol, cplx(th0r, 0.), lam, R, T, false);
}
// tmm.interface_R / interface_T (src/AMultilayer.cxx:26-56 interface_r/t, then |r|^2 and the power factor of T)
void interface_RT(int pol, cplx n_i, cplx n_f, cplx th_i, cplx th_f, double& R, double& T) {
  cplx r, t;
  if (pol == 0) {
    r = (n_i * std::cos(th_i) - n_f * std::cos(th_f)) / (n_i * std::cos(th_i) + n_f * std::cos(th_f));
    t = 2. * n_i * std::cos(th_i) / (n_i * std::cos(th_i) + n_f * std::cos(th_f));
    T = std::abs(t * t) * (((n_f * std::cos(th_f)).real()) / (n_i * std::cos(th_i)).real());
  } else {
    r = (n_f * std::cos(th_i) - n_i * std::cos(th_f)) / (n_f * std::cos(th_i) + n_i * std::cos(th_f));
    t = 2. * n_i * std::cos(th_i) / (n_f * std::cos(th_i) + n_i * std::cos(th_f));
    T = std::abs(t * t) * (((n_f * std::conj(std::cos(th_f))).real()) / (n_i * std::conj(std::cos(th_i))).real());
  }
  R = std::abs(r) * std::abs(r);
}
// AMultilayer::IncoherentTMM, src/AMultilayer.cxx:484-731 (tmm.inc_group_layers + tmm.inc_tmm)
void incoherent_tmm(const rbg_scene_desc* d, int ml, int pol, cplx th_0, double lam, double& R, double& T) {
  std::vector<cplx> n_list;
  std::vector<double> thick;
  std::vector<bool> coh;
  multilayer_lists(d, ml, lam, n_list, thick, coh);
  const int N = (int)n_list.size();
  // inc_group_layers: stacks of consecutive coherent layers, each bracketed by the incoherent layers next to it
  std::vector<std::vector<int>> all_from_stack;
  std::vector<int> all_from_inc, stack_from_inc;
  bool in_stack = false;
  for (int i = 0; i < N; i++) {
    if (coh[i]) {
      if (!in_stack) { in_stack = true; all_from_stack.push_back({i - 1, i}); }
      else all_from_stack.back().push_back(i);
    } else {
      all_from_inc.push_back(i);
      if (!in_stack) stack_from_inc.push_back(-1);
      else {
        in_stack = false;
        stack_from_inc.push_back((int)all_from_stack.size() - 1);
        all_from_stack.back().push_back(i);
      }
    }
  }
  // ListSnell
  std::vector<cplx> th_list(N);
  for (int i = 0; i < N; i++) th_list[i] = std::asin(n_list[0] * std::sin(th_0) / n_list[i]);
  if (!is_forward_angle(n_list[0], th_list[0])) th_list[0] = kPi - th_list[0];
  if (!is_forward_angle(n_list[N - 1], th_list[N - 1])) th_list[N - 1] = kPi - th_list[N - 1];
  // coherent stacks, forwards and backwards
  std::vector<std::pair<double, double>> fwd, bwd;
  for (auto& st : all_from_stack) {
    std::vector<cplx> sn;
    std::vector<double> sd;
    for (size_t k = 0; k < st.size(); k++) {
      sn.push_back(n_list[st[k]]);
      sd.push_back(k == 0 || k + 1 == st.size() ? kInf : thick[st[k]]);
    }
    double r, t;
    coherent_tmm_lists(sn, sd, pol, th_list[st.front()], lam, r, t);
    fwd.push_back({r, t});
    std::reverse(sn.begin(), sn.end());
    std::reverse(sd.begin(), sd.end());
    coherent_tmm_lists(sn, sd, pol, th_list[st.back()], lam, r, t);
    bwd.push_back({r, t});
  }
  const int NI = (int)all_from_inc.size();
  std::vector<double> P(NI, 0.);
  for (int k = 1; k < NI - 1; k++) {
    int i = all_from_inc[k];
    P[k] = exp(-4 * kPi * thick[i] * (n_list[i] * std::cos(th_list[i])).imag() / lam);
    if (P[k] < 1e-30) P[k] = 1e-30;
  }
  std::vector<std::vector<double>> Tl(NI, std::vector<double>(NI, 0.)), Rl(NI, std::vector<double>(NI, 0.));
  for (int k = 0; k < NI - 1; k++) {
    int a = all_from_inc[k], next_stack = stack_from_inc[k + 1];
    if (next_stack < 0) {
      interface_RT(pol, n_list[a], n_list[a + 1], th_list[a], th_list[a + 1], Rl[k][k + 1], Tl[k][k + 1]);
      interface_RT(pol, n_list[a + 1], n_list[a], th_list[a + 1], th_list[a], Rl[k + 1][k], Tl[k + 1][k]);
    } else {
      Rl[k][k + 1] = fwd[next_stack].first; Tl[k][k + 1] = fwd[next_stack].second;
      Rl[k + 1][k] = bwd[next_stack].first; Tl[k + 1][k] = bwd[next_stack].second;
    }
  }
  double L[2][2] = {{1 / Tl[0][1], -Rl[1][0] / Tl[0][1]}, {Rl[0][1] / Tl[0][1], (Tl[1][0] * Tl[0][1] - Rl[1][0] * Rl[0][1]) / Tl[0][1]}};
  for (int k = 1; k < NI - 1; k++) {
    double L1[2][2] = {{1 / P[k], 0}, {0, P[k]}};
    double L2[2][2] = {{1, -Rl[k + 1][k]}, {Rl[k][k + 1], Tl[k + 1][k] * Tl[k][k + 1] - Rl[k + 1][k] * Rl[k][k + 1]}};
    double Lk[2][2], Ln[2][2];
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 2; c++) Lk[r][c] = (L1[r][0] * L2[0][c] + L1[r][1] * L2[1][c]) * (1 / Tl[k][k + 1]);
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 2; c++) Ln[r][c] = L[r][0] * Lk[0][c] + L[r][1] * Lk[1][c];
    memcpy(L, Ln, sizeof(L));
  }
  T = 1 / L[0][0];
  R = L[1][0] / L[0][0];
}
// include/AMultilayer.h:114-132
void coherent_tmm_mixed(const rbg_scene_desc* d, int ml, double th, double lam, double& R, double& T) {
  const rbg_multilayer& M = d->multilayers[ml];
  if (M.table_r >= 0 && M.table_t >= 0) {
    R = th2_interp(d, M.table_r, lam, th);
    T = th2_interp(d, M.table_t, lam, th);
    return;
  }
  double rp, tp, rs, ts;
  coherent_tmm(d, ml, 1, th, lam, rp, tp);
  coherent_tmm(d, ml, 0, th, lam, rs, ts);
  R = (rp + rs) / 2.;
  T = (tp + ts) / 2.;
}

// ================================================================== shapes
// Generic helper for shapes handled analytically (sphere, pgon, pcon): given all candidate
// surface-crossing parameters along the ray, classify the intervals between them by Contains() at
// their midpoints and return the first transition.
template <class F> double first_transition(std::vector<double>& c, const double* p, const double* d, bool from_inside, F inside_at) {
  std::sort(c.begin(), c.end());
  double prev = 0;
  for (double t : c) {
    if (!(t > 1e-11) || t > 1e29) continue;
    if (t - prev < 1e-12) { prev = t; continue; }
    bool in = inside_at(0.5 * (prev + t));
    if (from_inside ? !in : in) return prev;
    prev = t;
  }
  if (from_inside) return prev;
  return kBig;
}
inline void add_quadratic(std::vector<double>& c, double A, double B, double C) {  // A t^2 + B t + C = 0
  if (fabs(A) < 1e-300 || fabs(A) < 1e-14 * fabs(B)) {
    if (B != 0) c.push_back(-C / B);
    return;
  }
  double disc = B * B - 4 * A * C;
  if (disc < 0) return;
  double s = sqrt(disc), q = -0.5 * (B + (B >= 0 ? s : -s));
  c.push_back(q / A);
  if (q != 0) c.push_back(C / q);
}

// ---- TGeoBBox
bool bbox_contains(const double* P, const double* p) {
  return !(fabs(p[0] - P[3]) > P[0] || fabs(p[1] - P[4]) > P[1] || fabs(p[2] - P[5]) > P[2]);
}
double bbox_dist_in(const double* P, const double* p, const double* d) {
  double np[3] = {p[0] - P[3], p[1] - P[4], p[2] - P[5]}, smin = kBig;
  for (int i = 0; i < 3; i++)
    if (d[i] != 0) {
      double s = d[i] > 0 ? (P[i] - np[i]) / d[i] : -(P[i] + np[i]) / d[i];
      if (s < 0) return 0.0;
      if (s < smin) smin = s;
    }
  return smin;
}
double bbox_dist_out(const double* P, const double* p, const double* d, double step) {
  double np[3] = {p[0] - P[3], p[1] - P[4], p[2] - P[5]}, saf[3];
  bool in = true;
  for (int i = 0; i < 3; i++) {
    saf[i] = fabs(np[i]) - P[i];
    if (saf[i] >= step) return kBig;
    if (in && saf[i] > 0) in = false;
  }
  if (in) {
    int j = 0;
    double ss = saf[0];
    if (saf[1] > ss) { ss = saf[1]; j = 1; }
    if (saf[2] > ss) j = 2;
    if (np[j] * d[j] > 0) return kBig;
    return 0.0;
  }
  for (int i = 0; i < 3; i++) {
    if (saf[i] < 0) continue;
    if (np[i] * d[i] >= 0) continue;
    double snxt = saf[i] / fabs(d[i]);
    bool ok = true;
    for (int j = 0; j < 3; j++) {
      if (j == i) continue;
      if (fabs(np[j] + snxt * d[j]) > P[j]) { ok = false; break; }
    }
    if (ok) return snxt;
  }
  return kBig;
}
void bbox_normal(const double* P, const double* p, const double* d, double* n) {
  double saf[3] = {fabs(P[0] - fabs(p[0] - P[3])), fabs(P[1] - fabs(p[1] - P[4])), fabs(P[2] - fabs(p[2] - P[5]))};
  int i = saf[1] < saf[0] ? 1 : 0;
  if (saf[2] < saf[i]) i = 2;
  n[0] = n[1] = n[2] = 0;
  n[i] = d[i] > 0 ? 1 : -1;
}

// ---- TGeoTube
void dist_to_tube(double rsq, double nsq, double rdotn, double radius, double& b, double& delta) {
  double t1 = 1. / nsq, t3 = rsq - radius * radius;
  b = t1 * rdotn;
  double c = t1 * t3;
  delta = b * b - c;
  if (delta > 0) delta = sqrt(delta);
  else delta = -1;
}
bool tube_contains(const double* P, const double* p) {
  if (fabs(p[2]) > P[2]) return false;
  double r2 = p[0] * p[0] + p[1] * p[1];
  return !(r2 < P[0] * P[0] || r2 > P[1] * P[1]);
}
double tube_dist_in(double rmin, double rmax, double dz, const double* p, const double* d) {
  double sz = kBig;
  if (d[2]) {
    sz = ((d[2] >= 0 ? dz : -dz) - p[2]) / d[2];
    if (sz <= 0) return 0.0;
  }
  double nsq = d[0] * d[0] + d[1] * d[1];
  if (fabs(nsq) < kTol) return sz;
  double rsq = p[0] * p[0] + p[1] * p[1], rdotn = p[0] * d[0] + p[1] * d[1], b, dl;
  if (rmin > 0) {
    if (rsq <= rmin * rmin + kTol) {
      if (rdotn < 0) return 0.0;
    } else if (rdotn < 0) {
      dist_to_tube(rsq, nsq, rdotn, rmin, b, dl);
      if (dl > 0) {
        double sr = -b - dl;
        if (sr > 0) return std::min(sz, sr);
      }
    }
  }
  if (rsq >= rmax * rmax - kTol) {
    if (rdotn >= 0) return 0.0;
  }
  dist_to_tube(rsq, nsq, rdotn, rmax, b, dl);
  if (dl > 0) {
    double sr = -b + dl;
    if (sr > 0) return std::min(sz, sr);
  }
  return 0.;
}
double tube_dist_out(double rmin, double rmax, double dz, const double* p, const double* d) {
  double rmaxsq = rmax * rmax, rminsq = rmin * rmin;
  double zi = dz - fabs(p[2]);
  bool inz = !(zi < 0);
  if (!inz) {
    if (p[2] * d[2] >= 0) return kBig;
    double s = -zi / fabs(d[2]);
    double xi = p[0] + s * d[0], yi = p[1] + s * d[1], r2 = xi * xi + yi * yi;
    if (rminsq <= r2 && r2 <= rmaxsq) return s;
  }
  double rsq = p[0] * p[0] + p[1] * p[1], nsq = d[0] * d[0] + d[1] * d[1], rdotn = p[0] * d[0] + p[1] * d[1], b, dl;
  bool inrmax = rsq <= rmaxsq + kTol, inrmin = rsq >= rminsq - kTol;
  bool in = inz && inrmin && inrmax;
  if (in) {
    bool checkout = false;
    double r = sqrt(rsq);
    if (zi < rmax - r) {
      if (fabs(rmin) < kTol || zi < r - rmin) {
        if (p[2] * d[2] < 0) return 0.0;
        return kBig;
      }
    }
    if ((rmaxsq - rsq) < (rsq - rminsq)) checkout = true;
    if (checkout) {
      if (rdotn >= 0) return kBig;
      return 0.0;
    }
    if (fabs(rmin) < kTol) return 0.0;
    if (rdotn >= 0) return 0.0;
    if (fabs(nsq) < kTol) return kBig;
    dist_to_tube(rsq, nsq, rdotn, rmin, b, dl);
    if (dl > 0) {
      double s = -b + dl;
      if (s > 0) {
        zi = p[2] + s * d[2];
        if (fabs(zi) <= dz) return s;
      }
    }
    return kBig;
  }
  if (fabs(nsq) < kTol) return kBig;
  if (!inrmax) {
    dist_to_tube(rsq, nsq, rdotn, rmax, b, dl);
    if (dl > 0) {
      double s = -b - dl;
      if (s > 0) {
        zi = p[2] + s * d[2];
        if (fabs(zi) <= dz) return s;
      }
    }
  }
  if (rmin > 0) {
    dist_to_tube(rsq, nsq, rdotn, rmin, b, dl);
    if (dl > 0) {
      double s = -b + dl;
      if (s > 0) {
        zi = p[2] + s * d[2];
        if (fabs(zi) <= dz) return s;
      }
    }
  }
  return kBig;
}
void tube_normal(const double* P, const double* p, const double* d, double* n) {
  double rsq = p[0] * p[0] + p[1] * p[1], r = sqrt(rsq);
  double saf[3] = {fabs(P[2] - fabs(p[2])), P[0] > 1e-10 ? fabs(r - P[0]) : kBig, fabs(P[1] - r)};
  int i = saf[1] < saf[0] ? 1 : 0;
  if (saf[2] < saf[i]) i = 2;
  if (i == 0) {
    n[0] = n[1] = 0;
    n[2] = d[2] >= 0 ? 1 : -1;
    return;
  }
  n[2] = 0;
  double phi = ATan2(p[1], p[0]);
  n[0] = cos(phi);
  n[1] = sin(phi);
  if (n[0] * d[0] + n[1] * d[1] < 0) { n[0] = -n[0]; n[1] = -n[1]; }
}

// ---- TGeoParaboloid  (z = a r^2 + b)
struct Para {
  double rlo, rhi, dz, a, b;
  explicit Para(const double* P) : rlo(P[0]), rhi(P[1]), dz(P[2]) {
    double dd = 1. / (rhi * rhi - rlo * rlo);
    a = 2. * dz * dd;
    b = -dz * (rlo * rlo + rhi * rhi) * dd;
  }
};
bool para_contains(const Para& q, const double* p) {
  if (fabs(p[2]) > q.dz) return false;
  double aa = q.a * (p[2] - q.b);
  if (aa < 0) return false;
  double rsq = p[0] * p[0] + p[1] * p[1];
  return !(aa < q.a * q.a * rsq);
}
double dist_to_paraboloid(const Para& q, const double* p, const double* d, bool in) {
  double rsq = p[0] * p[0] + p[1] * p[1];
  double a = q.a * (d[0] * d[0] + d[1] * d[1]), b = 2. * q.a * (p[0] * d[0] + p[1] * d[1]) - d[2], c = q.a * rsq + q.b - p[2];
  double dist = kBig;
  if (fabs(a) < kTol) {
    if (fabs(b) < kTol) return dist;
    dist = -c / b;
    if (dist < 0) return kBig;
    return dist;
  }
  double ainv = 1. / a, sum = -b * ainv, prod = c * ainv, delta = sum * sum - 4. * prod;
  if (delta < 0) return dist;
  delta = sqrt(delta);
  double sone = ainv >= 0 ? 1. : -1.;
  int i = -1;
  while (i < 2) {
    dist = 0.5 * (sum + i * sone * delta);
    i += 2;
    if (dist < 0) continue;
    if (dist < 1.E-8) {
      double talf = -2. * q.a * sqrt(rsq), phi = ATan2(p[1], p[0]);
      double ndotd = talf * (cos(phi) * d[0] + sin(phi) * d[1]) + d[2];
      if (!in) ndotd *= -1;
      if (ndotd < 0) return dist;
    } else return dist;
  }
  return kBig;
}
double para_dist_in(const Para& q, const double* p, const double* d) {
  double dz = kBig;
  if (d[2] < 0) dz = -(p[2] + q.dz) / d[2];
  else if (d[2] > 0) dz = (q.dz - p[2]) / d[2];
  return std::min(dz, dist_to_paraboloid(q, p, d, true));
}
double para_dist_out(const Para& q, const double* p, const double* d) {
  if (p[2] <= -q.dz) {
    if (d[2] <= 0) return kBig;
    double snxt = -(q.dz + p[2]) / d[2], xn = p[0] + snxt * d[0], yn = p[1] + snxt * d[1];
    if (xn * xn + yn * yn <= q.rlo * q.rlo) return snxt;
  } else if (p[2] >= q.dz) {
    if (d[2] >= 0) return kBig;
    double snxt = (q.dz - p[2]) / d[2], xn = p[0] + snxt * d[0], yn = p[1] + snxt * d[1];
    if (xn * xn + yn * yn <= q.rhi * q.rhi) return snxt;
  }
  double snxt = dist_to_paraboloid(q, p, d, false);
  if (snxt > 1E20) return snxt;
  double zn = p[2] + snxt * d[2];
  if (fabs(zn) <= q.dz) return snxt;
  return kBig;
}
void para_normal(const Para& q, const double* p, const double* d, double* n) {
  n[0] = n[1] = 0.0;
  if ((fabs(p[2]) - q.dz) > -1E-5) { n[2] = d[2] >= 0 ? 1. : -1.; return; }
  double safz = q.dz - fabs(p[2]), r = sqrt(p[0] * p[0] + p[1] * p[1]);
  double safr = fabs(r - sqrt((p[2] - q.b) / q.a));
  if (safz < safr) { n[2] = d[2] >= 0 ? 1. : -1.; return; }
  double talf = -2. * q.a * r, calf = 1. / sqrt(1. + talf * talf), salf = talf * calf, phi = ATan2(p[1], p[0]);
  n[0] = salf * cos(phi);
  n[1] = salf * sin(phi);
  n[2] = calf;
  if (dot3(n, d) < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
}

// ---- TGeoSphere (rmin,rmax,theta1,theta2,phi1,phi2 in deg)
bool sphere_contains(const double* P, const double* p) {
  double r2 = dot3(p, p);
  if (P[0] > 0 && r2 < P[0] * P[0]) return false;
  if (r2 > P[1] * P[1]) return false;
  if (r2 < 1E-20) return true;
  bool phiseg = fabs(P[5] - P[4] - 360.) > 1e-9;
  if (phiseg) {
    double phi = ATan2(p[1], p[0]) * 180. / kPi;
    while (phi < P[4]) phi += 360.;
    if (phi - P[4] > P[5] - P[4]) return false;
  }
  if (P[2] > 0 || P[3] < 180) {
    double theta = ACosT(p[2] / sqrt(r2)) * 180. / kPi;
    if (theta < P[2] || theta > P[3]) return false;
  }
  return true;
}
void sphere_candidates(const double* P, const double* p, const double* d, std::vector<double>& c) {
  double a = dot3(d, d), b = 2 * dot3(p, d), pp = dot3(p, p);
  if (P[0] > 0) add_quadratic(c, a, b, pp - P[0] * P[0]);
  add_quadratic(c, a, b, pp - P[1] * P[1]);
  for (int k = 2; k <= 3; k++) {
    double th = P[k];
    if ((k == 2 && th <= 0) || (k == 3 && th >= 180)) continue;
    double co = cos(th * kPi / 180.), si = sin(th * kPi / 180.), c2 = co * co, s2 = si * si;
    add_quadratic(c, (d[0] * d[0] + d[1] * d[1]) * c2 - d[2] * d[2] * s2, 2 * ((p[0] * d[0] + p[1] * d[1]) * c2 - p[2] * d[2] * s2),
                  (p[0] * p[0] + p[1] * p[1]) * c2 - p[2] * p[2] * s2);
  }
  if (fabs(P[5] - P[4] - 360.) > 1e-9)
    for (int k = 4; k <= 5; k++) {
      double co = cos(P[k] * kPi / 180.), si = sin(P[k] * kPi / 180.), den = d[1] * co - d[0] * si;
      if (den != 0) c.push_back(-(p[1] * co - p[0] * si) / den);
    }
}
double sphere_dist(const double* P, const double* p, const double* d, bool from_inside) {
  std::vector<double> c;
  sphere_candidates(P, p, d, c);
  return first_transition(c, p, d, from_inside, [&](double t) {
    double q[3] = {p[0] + t * d[0], p[1] + t * d[1], p[2] + t * d[2]};
    return sphere_contains(P, q);
  });
}
void sphere_normal(const double* P, const double* p, const double* d, double* n) {
  double r2 = dot3(p, p), r = sqrt(r2), rxy = sqrt(p[0] * p[0] + p[1] * p[1]);
  double saf[6] = {P[0] > 0 ? fabs(r - P[0]) : kBig, fabs(P[1] - r), kBig, kBig, kBig, kBig};
  double th = ACosT(r > 0 ? p[2] / r : 1.);
  if (P[2] > 0) saf[2] = r * fabs(sin(th - P[2] * kPi / 180.));
  if (P[3] < 180) saf[3] = r * fabs(sin(P[3] * kPi / 180. - th));
  bool phiseg = fabs(P[5] - P[4] - 360.) > 1e-9;
  if (phiseg)
    for (int k = 4; k <= 5; k++) {
      double co = cos(P[k] * kPi / 180.), si = sin(P[k] * kPi / 180.);
      saf[k] = fabs(p[1] * co - p[0] * si);
    }
  int i = 0;
  for (int k = 1; k < 6; k++)
    if (saf[k] < saf[i]) i = k;
  if (i < 2) {
    if (r > 0) { n[0] = p[0] / r; n[1] = p[1] / r; n[2] = p[2] / r; }
    else { n[0] = n[1] = 0; n[2] = 1; }
  } else if (i < 4) {  // theta cone: normal = d(theta)/dx direction
    double cth = cos(P[i] * kPi / 180.), sth = sin(P[i] * kPi / 180.);
    double cph = rxy > 0 ? p[0] / rxy : 1, sph = rxy > 0 ? p[1] / rxy : 0;
    n[0] = cth * cph; n[1] = cth * sph; n[2] = -sth;
  } else {
    double co = cos(P[i] * kPi / 180.), si = sin(P[i] * kPi / 180.);
    n[0] = -si; n[1] = co; n[2] = 0;
  }
  if (dot3(n, d) < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
}

// ---- TGeoPgon / TGeoPcon  (phi1,dphi,[nedges,]nz, sections (z,rmin,rmax))
struct Poly {
  bool pgon;
  double phi1, dphi;
  int nedges, nz;
  const double* sec;
  Poly(const double* P, bool is_pgon) : pgon(is_pgon), phi1(P[0]), dphi(P[1]) {
    if (pgon) { nedges = (int)P[2]; nz = (int)P[3]; sec = P + 4; }
    else { nedges = 0; nz = (int)P[2]; sec = P + 3; }
  }
  double z(int i) const { return sec[3 * i]; }
  double rmin(int i) const { return sec[3 * i + 1]; }
  double rmax(int i) const { return sec[3 * i + 2]; }
};
bool poly_contains(const Poly& q, const double* p) {
  if (p[2] < q.z(0)) return false;
  if (p[2] > q.z(q.nz - 1)) return false;
  double r;
  if (q.pgon) {
    double divphi = q.dphi / q.nedges;
    double phi = ATan2(p[1], p[0]) * 180. / kPi;
    while (phi < q.phi1) phi += 360.0;
    double ddp = phi - q.phi1;
    if (ddp > q.dphi) return false;
    int ipsec = std::min(int(ddp / divphi), q.nedges - 1);
    double ph0 = (q.phi1 + divphi * (ipsec + 0.5)) * kPi / 180.;
    r = p[0] * cos(ph0) + p[1] * sin(ph0);
  } else {
    r = sqrt(p[0] * p[0] + p[1] * p[1]);
    if (fabs(q.dphi - 360.) > 1e-9) {
      double phi = ATan2(p[1], p[0]) * 180. / kPi;
      while (phi < q.phi1) phi += 360.0;
      if (phi - q.phi1 > q.dphi) return false;
    }
  }
  int iz = 0;  // TMath::BinarySearch: largest i with z[i] <= p[2]
  for (int i = 0; i < q.nz; i++)
    if (q.z(i) <= p[2]) iz = i;
  if (iz == q.nz - 1) return !(r < q.rmin(iz) || r > q.rmax(iz));
  double dz = q.z(iz + 1) - q.z(iz);
  if (dz < 1E-8) {
    double rmin = std::min(q.rmin(iz), q.rmin(iz + 1)), rmax = std::max(q.rmax(iz), q.rmax(iz + 1));
    return !(r < rmin || r > rmax);
  }
  double dzrat = (p[2] - q.z(iz)) / dz;
  double rmin = q.rmin(iz) + dzrat * (q.rmin(iz + 1) - q.rmin(iz));
  if (r < rmin) return false;
  double rmax = q.rmax(iz) + dzrat * (q.rmax(iz + 1) - q.rmax(iz));
  return !(r > rmax);
}
void poly_candidates(const Poly& q, const double* p, const double* d, std::vector<double>& c) {
  for (int i = 0; i < q.nz; i++)
    if (d[2] != 0) c.push_back((q.z(i) - p[2]) / d[2]);
  for (int k = 0; k + 1 < q.nz; k++) {
    double z0 = q.z(k), z1 = q.z(k + 1), dz = z1 - z0;
    if (dz < 1E-8) continue;
    for (int w = 0; w < 2; w++) {
      double r0 = w ? q.rmax(k) : q.rmin(k), r1 = w ? q.rmax(k + 1) : q.rmin(k + 1);
      if (!w && r0 <= 0 && r1 <= 0) continue;
      double s = (r1 - r0) / dz;
      size_t first = c.size();
      if (q.pgon) {
        double divphi = q.dphi / q.nedges;
        for (int e = 0; e < q.nedges; e++) {
          double ph = (q.phi1 + divphi * (e + 0.5)) * kPi / 180., ux = cos(ph), uy = sin(ph);
          double den = d[0] * ux + d[1] * uy - s * d[2];
          if (den != 0) c.push_back((r0 + (p[2] - z0) * s - (p[0] * ux + p[1] * uy)) / den);
        }
      } else {
        double a0 = r0 + (p[2] - z0) * s, b0 = s * d[2];
        add_quadratic(c, d[0] * d[0] + d[1] * d[1] - b0 * b0, 2 * (p[0] * d[0] + p[1] * d[1] - a0 * b0), p[0] * p[0] + p[1] * p[1] - a0 * a0);
      }
      // keep only crossings inside this z slab
      size_t o = first;
      for (size_t i = first; i < c.size(); i++) {
        double zz = p[2] + c[i] * d[2];
        if (zz >= z0 - 1e-9 && zz <= z1 + 1e-9) c[o++] = c[i];
      }
      c.resize(o);
    }
  }
  if (fabs(q.dphi - 360.) > 1e-9)
    for (int k = 0; k < 2; k++) {
      double ph = (q.phi1 + k * q.dphi) * kPi / 180., co = cos(ph), si = sin(ph), den = d[1] * co - d[0] * si;
      if (den != 0) c.push_back(-(p[1] * co - p[0] * si) / den);
    }
}
double poly_dist(const Poly& q, const double* p, const double* d, bool from_inside) {
  std::vector<double> c;
  poly_candidates(q, p, d, c);
  return first_transition(c, p, d, from_inside, [&](double t) {
    double x[3] = {p[0] + t * d[0], p[1] + t * d[1], p[2] + t * d[2]};
    return poly_contains(q, x);
  });
}
void poly_normal(const Poly& q, const double* p, const double* d, double* n) {
  // closest of: end caps, radius-changing planes, outer/inner lateral faces of the slab containing z
  double best = kBig;
  n[0] = n[1] = 0; n[2] = 1;
  double ux = 0, uy = 0, r;
  if (q.pgon) {
    double divphi = q.dphi / q.nedges, phi = ATan2(p[1], p[0]) * 180. / kPi;
    while (phi < q.phi1) phi += 360.0;
    int ipsec = std::max(0, std::min(int((phi - q.phi1) / divphi), q.nedges - 1));
    double ph0 = (q.phi1 + divphi * (ipsec + 0.5)) * kPi / 180.;
    ux = cos(ph0); uy = sin(ph0);
    r = p[0] * ux + p[1] * uy;
  } else {
    r = sqrt(p[0] * p[0] + p[1] * p[1]);
    ux = r > 0 ? p[0] / r : 1; uy = r > 0 ? p[1] / r : 0;
  }
  for (int i = 0; i < q.nz; i++) {
    bool cap = i == 0 || i == q.nz - 1;
    bool step = (i + 1 < q.nz && q.z(i + 1) - q.z(i) < 1e-8) || (i > 0 && q.z(i) - q.z(i - 1) < 1e-8);
    if (!cap && !step) continue;
    double s = fabs(p[2] - q.z(i));
    if (s < best) { best = s; n[0] = n[1] = 0; n[2] = 1; }
  }
  for (int k = 0; k + 1 < q.nz; k++) {
    double z0 = q.z(k), z1 = q.z(k + 1), dz = z1 - z0;
    if (dz < 1e-8 || p[2] < z0 - 1e-6 || p[2] > z1 + 1e-6) continue;
    for (int w = 0; w < 2; w++) {
      double r0 = w ? q.rmax(k) : q.rmin(k), r1 = w ? q.rmax(k + 1) : q.rmin(k + 1);
      if (!w && r0 <= 0 && r1 <= 0) continue;
      double s = (r1 - r0) / dz, rr = r0 + (p[2] - z0) * s, nn = sqrt(1 + s * s), dist = fabs(r - rr) / nn;
      if (dist < best) { best = dist; n[0] = ux / nn; n[1] = uy / nn; n[2] = -s / nn; }
    }
  }
  if (fabs(q.dphi - 360.) > 1e-9)  // azimuthal segment: the two phi planes (TGeoShape::IsCloseToPhi / NormalPhi)
    for (int k = 0; k < 2; k++) {
      double ph = (q.phi1 + k * q.dphi) * kPi / 180., co = cos(ph), si = sin(ph);
      if (p[0] * co + p[1] * si < 0) continue;
      double dist = fabs(p[1] * co - p[0] * si);
      if (dist < best) { best = dist; n[0] = -si; n[1] = co; n[2] = 0; }
    }
  if (dot3(n, d) < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
}

// ---- AGeoAsphericDisk   src/AGeoAsphericDisk.cxx
struct Asph {
  double z1, z2, c1, c2, k1, k2, rmin, rmax, oz, dz;
  int n1, n2;
  const double *K1, *K2;
  explicit Asph(const double* P)
      : z1(P[0]), z2(P[1]), c1(P[2]), c2(P[3]), k1(P[4]), k2(P[5]), rmin(P[6]), rmax(P[7]), oz(P[10]), dz(P[11]), n1((int)P[8]), n2((int)P[9]),
        K1(P + 12), K2(P + 12 + (int)P[8]) {}
  // :126-153  (returns false where the reference throws)
  bool F(int s, double r, double& out) const {
    double c = s == 1 ? c1 : c2, kap = s == 1 ? k1 : k2, z0 = s == 1 ? z1 : z2;
    const double* K = s == 1 ? K1 : K2;
    int n = s == 1 ? n1 : n2;
    double p = r * r * c * c * kap;
    if (1 - p < 0) return false;
    double ret = z0 + r * r * c / (1 + sqrt(1 - p));
    for (int i = 0; i < n; i++) ret += K[i] * pow(r, 2 * (i + 1));
    out = ret;
    return true;
  }
  // :96-123
  bool dF(int s, double r, double& out) const {
    double c = s == 1 ? c1 : c2, kap = s == 1 ? k1 : k2;
    const double* K = s == 1 ? K1 : K2;
    int n = s == 1 ? n1 : n2;
    double p = r * r * c * c * kap;
    if (1 - p <= 0) return false;
    double ret = r * c / sqrt(1 - p);
    for (int i = 0; i < n; i++) ret += 2 * (i + 1) * K[i] * pow(r, 2 * (i + 1) - 1);
    out = ret;
    return true;
  }
};
bool asph_contains(const Asph& a, const double* p) {  // :330-347
  double r = sqrt(p[0] * p[0] + p[1] * p[1]);
  if (r > a.rmax || r < a.rmin) return false;
  double f1, f2;
  if (!a.F(1, r, f1) || !a.F(2, r, f2)) return false;
  return !(p[2] < f1 || f2 < p[2]);
}
double asph_dist_to_asphere(const Asph& a, int n, const double* point, const double* dir) {  // :415-526
  double H2 = point[0] * point[0] + point[1] * point[1];
  double d = n == 1 ? a.z1 : a.z2, curve = n == 1 ? a.c1 : a.c2, kappa = n == 1 ? a.k1 : a.k2;
  const double* K = n == 1 ? a.K1 : a.K2;
  int npol = n == 1 ? a.n1 : a.n2;
  double p = -((point[2] - d) * dir[2] + point[0] * dir[0] + point[1] * dir[1]);
  double M = p * dir[2] + point[2] - d;
  double M2 = (point[2] - d) * (point[2] - d) + H2 - p * p;
  double check = 1 - (M2 * curve - 2 * M) * curve / dir[2] / dir[2];
  if (check < 0) return kBig;
  double q = p + (M2 * curve - 2 * M) / (dir[2] * (1 + sqrt(check)));
  double np[3] = {point[0] + q * dir[0], point[1] + q * dir[1], point[2] + q * dir[2] - d};
  for (int i = 0;; i++) {
    if (i > 100) return kBig;
    H2 = np[0] * np[0] + np[1] * np[1];
    check = 1 - kappa * H2 * curve * curve;
    if (check < 0) return kBig;
    double l = sqrt(check);
    double x = 0;
    if (curve != 0) x += (1 - l) / curve / kappa;
    for (int j = 0; j < npol; j++) x += K[j] * pow(H2, j + 1);
    double v = 0;
    for (int j = 0; j < npol; j++) v += 2 * (j + 1) * K[j] * pow(H2, j);
    v = curve * kappa + l * v;
    double m = -np[0] * v, nn = -np[1] * v;
    double norm = sqrt(l * l + m * m + nn * nn);
    l /= norm; m /= norm; nn /= norm;
    check = dir[2] * l + dir[0] * m + dir[1] * nn;
    if (check == 0) return kBig;
    double e = l * (x - np[2]) / check;
    for (int j = 0; j < 3; j++) np[j] += e * dir[j];
    if (fabs(e) < 1e-10) break;
  }
  np[2] += d;
  check = dir[0] * (np[0] - point[0]) + dir[1] * (np[1] - point[1]) + dir[2] * (np[2] - point[2]);
  if (check < 0) return kBig;
  double dist_to_zaxis = pow(np[0] * np[0] + np[1] * np[1], 0.5);
  if (dist_to_zaxis < a.rmin || dist_to_zaxis > a.rmax) return kBig;
  return sqrt(pow(np[0] - point[0], 2) + pow(np[1] - point[1], 2) + pow(np[2] - point[2], 2));
}
double asph_dist_to_cyl(const Asph& a, double R, const double* point, const double* dir) {  // :529-684
  double rsq = point[0] * point[0] + point[1] * point[1], nsq = dir[0] * dir[0] + dir[1] * dir[1];
  if (sqrt(nsq) < kTol) return kBig;
  double rdotn = point[0] * dir[0] + point[1] * dir[1], b, delta;
  dist_to_tube(rsq, nsq, rdotn, R, b, delta);
  if (delta < 0) return kBig;
  double t1 = -b + delta, t2 = -b - delta;
  if (t1 < 0 && t2 < 0) return kBig;
  double zmin, zmax;
  if (!a.F(1, R, zmin) || !a.F(2, R, zmax)) throw std::runtime_error("AGeoAsphericDisk: CalcF out of domain in DistToInner/Outer");
  if (t2 > 0) {
    if (t1 > 0) {
      double z1 = t1 * dir[2] + point[2], z2 = t2 * dir[2] + point[2];
      if (z1 < zmin || zmax < z1) t1 = kBig;
      if (z2 < zmin || zmax < z2) t2 = kBig;
      return t1 < t2 ? t1 : t2;
    }
  } else if (t2 == 0) {
    if (t1 > 0) {
      if (zmin <= point[2] && point[2] <= zmax) return 0;
      double z1 = t1 * dir[2] + point[2];
      if (zmin <= z1 && z1 <= zmax) return t1;
    } else if (t1 == 0) {
      if (zmin <= point[2] && point[2] <= zmax) return 0;
    }
  } else {
    if (t1 > 0) {
      double z1 = t1 * dir[2] + point[2];
      if (zmin <= z1 && z1 <= zmax) return t1;
    } else if (t1 == 0) {
      if (zmin <= point[2] && point[2] <= zmax) return 0;
    }
  }
  return kBig;
}
double asph_dist4(const Asph& a, const double* p, const double* d) {  // :376-382 / :405-411 (LocMin = first minimum)
  double v[4] = {asph_dist_to_asphere(a, 1, p, d), asph_dist_to_asphere(a, 2, p, d), a.rmin > 0 ? asph_dist_to_cyl(a, a.rmin, p, d) : kBig,
                 asph_dist_to_cyl(a, a.rmax, p, d)};
  double m = v[0];
  for (int i = 1; i < 4; i++)
    if (v[i] < m) m = v[i];
  return m;
}
double asph_dist_out(const Asph& a, const double* p, const double* d, double step) {  // :386-412
  double p_[3] = {p[0], p[1], p[2] - a.oz};
  double sdist = tube_dist_out(a.rmin, a.rmax, a.dz, p_, d);
  if (sdist >= step) return kBig;
  return asph_dist4(a, p, d);
}
void asph_normal(const Asph& a, const double* p, const double* d, double* n) {  // :246-327
  double r = sqrt(p[0] * p[0] + p[1] * p[1]), phi = atan2(p[1], p[0]);
  double saf[4];
  saf[0] = a.rmin > 0 ? fabs(r - a.rmin) : kBig;
  saf[1] = fabs(r - a.rmax);
  double f1, f2, df1 = kBig, df2 = kBig;
  if (!a.F(1, r, f1)) saf[2] = kBig;
  else if (a.dF(1, r, df1)) saf[2] = fabs(f1 - p[2]) / sqrt(1 + df1 * df1);
  else saf[2] = kBig;
  if (!a.F(2, r, f2)) saf[3] = kBig;
  else if (a.dF(2, r, df2)) saf[3] = fabs(f2 - p[2]) / sqrt(1 + df2 * df2);
  else saf[3] = kBig;
  int i = 0;
  for (int k = 1; k < 4; k++)
    if (saf[k] < saf[i]) i = k;
  double nx = 0, nz = 0;
  if (i == 0 || i == 1) nx = 1;
  else {
    double df = i == 2 ? df1 : df2;
    if (df == 0) nz = 1;
    else { nx = df / sqrt(1 + df * df); nz = -1 / sqrt(1 + df * df); }
  }
  n[0] = nx * cos(phi);
  n[1] = nx * sin(phi);
  n[2] = nz;
  if (dot3(n, d) < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
}

// ---- AGeoWinstonCone2D / Poly   src/AGeoWinstonCone2D.cxx, src/AGeoWinstonConePoly.cxx
struct Winston {
  double r1, r2, dy, theta, dz, f;
  int npoly;
  Winston(const double* P, bool poly) : r1(P[0]), r2(P[1]), dy(poly ? 0 : P[2]), npoly(poly ? (int)P[2] : 0) {
    theta = asin(r2 / r1);               // 2D:551-567
    dz = (r1 + r2) / tan(theta) / 2.;
    f = r2 * (1 + sin(theta));
  }
  bool R(double z, double& out) const {  // 2D:82-98
    if (fabs(z) > dz + 1e-10) return false;
    double sint = sin(theta), cost = cos(theta), t = z + dz;
    double a0 = t * t * sint * sint - 4. * f * (t * cost + f), a1 = 2. * t * sint * cost + 4. * f * sint, a2 = cost * cost;
    out = (-a1 + sqrt(a1 * a1 - 4. * a0 * a2)) / (2 * a2) - r2;
    return true;
  }
  bool dRdZ(double z, double& out) const {  // 2D:59-79
    if (fabs(z) > dz + 1e-10) return false;
    double sint = sin(theta), cost = cos(theta), t = z + dz;
    double a0 = t * t * sint * sint - 4. * f * (t * cost + f), a1 = 2. * t * sint * cost + 4. * f * sint, a2 = cost * cost;
    double da0dt = 2 * t * sint * sint - 4 * f * cost, da1dt = 2 * sint * cost;
    out = (-da1dt + (a1 * da1dt - 2 * da0dt * a2) / sqrt(a1 * a1 - 4 * a0 * a2)) / (2 * a2);
    return true;
  }
};
double winston_dist_to_parabola(const Winston& w, const double* point, const double* dir, double phi, double open) {  // 2D:325-428
  double x = cos(phi) * point[0] + sin(phi) * point[1], y = -sin(phi) * point[0] + cos(phi) * point[1], z = point[2];
  double px = cos(phi) * dir[0] + sin(phi) * dir[1], py = -sin(phi) * dir[0] + cos(phi) * dir[1], pz = dir[2];
  if (px == 0 && pz == 0) return kBig;
  double cost = cos(w.theta), sint = sin(w.theta);
  double X = cost * (x + w.r2) + (z + w.dz) * sint, Z = -sint * (x + w.r2) + (z + w.dz) * cost + w.f;
  double alpha = ATan2(pz, px), ALPHA = alpha - w.theta, tanA = tan(ALPHA);
  double dist[2];
  double tmp = tanA * tanA - (X * tanA - Z) / w.f;
  if (tmp < 0) return kBig;
  double Xp, Xm;
  if (w.dz * 2 / fabs(tanA) < kTol) { Xp = X; Xm = X; }
  else { Xp = 2 * w.f * (tanA + sqrt(tmp)); Xm = 2 * w.f * (tanA - sqrt(tmp)); }
  double Xc[2] = {Xp, Xm};
  for (int k = 0; k < 2; k++) {
    double Zc = Xc[k] * Xc[k] / 4. / w.f;
    double xc = cost * Xc[k] - sint * (Zc - w.f) - w.r2, zc = sint * Xc[k] + cost * (Zc - w.f) - w.dz, yc;
    if (fabs(px) <= fabs(pz) && fabs(py) <= fabs(pz)) yc = y + (zc - z) * py / pz;
    else if (fabs(py) <= fabs(px) && fabs(pz) <= fabs(px)) yc = y + (xc - x) * py / px;
    else yc = y + (fabs(px) < 1e-5 ? (zc - z) * py / pz : (xc - x) * py / px);
    double dx = xc - x, dy = yc - y, dz = zc - z;
    if (xc < w.r2 || w.r1 < xc || zc < -w.dz || w.dz < zc || dx * px + dz * pz < 0) dist[k] = kBig;
    else if (fabs(ATan2(yc, xc)) <= open / 2.) dist[k] = sqrt(dx * dx + dy * dy + dz * dz);
    else dist[k] = kBig;
  }
  return std::min(dist[0], dist[1]);
}
bool winston_inside_polygon(const Winston& w, double x, double y, double r) {  // Poly:293-309
  double theta = ATan2(y, x);
  while (theta > kPi / w.npoly) theta -= 2 * kPi / w.npoly;
  while (theta < -kPi / w.npoly) theta += 2 * kPi / w.npoly;
  return !(sqrt(x * x + y * y) * cos(theta) > r);
}
bool winston_contains(const Winston& w, const double* p) {
  if (w.npoly) {  // Poly:125-137
    if (fabs(p[2]) > w.dz) return false;
    double r;
    if (!w.R(p[2], r)) throw std::runtime_error("AGeoWinstonConePoly::Contains: CalcR threw");
    return winston_inside_polygon(w, p[0], p[1], r);
  }
  if (fabs(p[1]) > w.dy || fabs(p[2]) > w.dz) return false;  // 2D:175-190
  double r;
  if (!w.R(p[2], r)) throw std::runtime_error("AGeoWinstonCone2D::Contains: CalcR threw");
  return !(fabs(p[0]) > r);
}
double winston_dist_in(const Winston& w, const double* p, const double* d) {
  double dzd = kBig;
  if (d[2] < 0) dzd = -(p[2] + w.dz) / d[2];
  else if (d[2] > 0) dzd = (w.dz - p[2]) / d[2];
  double best = dzd;
  if (w.npoly) {  // Poly:150-178
    for (int i = 0; i < w.npoly; i++) {
      double v = winston_dist_to_parabola(w, p, d, i * 2 * kPi / w.npoly, kPi);
      if (v < best) best = v;
    }
    return best;
  }
  double dyd = kBig;  // 2D:203-238
  if (d[1] < 0) dyd = -(p[1] + w.dy) / d[1];
  else if (d[1] > 0) dyd = (w.dy - p[1]) / d[1];
  if (dyd < best) best = dyd;
  double v = winston_dist_to_parabola(w, p, d, 0., kPi);
  if (v < best) best = v;
  v = winston_dist_to_parabola(w, p, d, kPi, kPi);
  if (v < best) best = v;
  return best;
}
double winston_dist_out(const Winston& w, const double* p, const double* d) {
  if (w.npoly) {  // Poly:181-226
    if (p[2] <= -w.dz) {
      if (d[2] <= 0) return kBig;
      double s = -(w.dz + p[2]) / d[2];
      if (winston_inside_polygon(w, p[0] + s * d[0], p[1] + s * d[1], w.r2)) return s;
    } else if (p[2] >= w.dz) {
      if (d[2] >= 0) return kBig;
      double s = (w.dz - p[2]) / d[2];
      if (winston_inside_polygon(w, p[0] + s * d[0], p[1] + s * d[1], w.r1)) return s;
    }
    double best = kBig;
    for (int i = 0; i < w.npoly; i++) {
      double v = winston_dist_to_parabola(w, p, d, i * 2 * kPi / w.npoly, 2 * kPi / w.npoly);
      if (v < best) best = v;
    }
    return best;
  }
  // 2D:241-322
  if (p[2] <= -w.dz) {
    if (d[2] <= 0) return kBig;
    double s = -(w.dz + p[2]) / d[2];
    if (fabs(p[0] + s * d[0]) <= w.r2 && fabs(p[1] + s * d[1]) <= w.dy) return s;
  } else if (p[2] >= w.dz) {
    if (d[2] >= 0) return kBig;
    double s = (w.dz - p[2]) / d[2];
    if (fabs(p[0] + s * d[0]) <= w.r1 && fabs(p[1] + s * d[1]) <= w.dy) return s;
  }
  if (p[1] <= -w.dy) {
    if (d[1] <= 0) return kBig;
    double s = -(w.dy + p[1]) / d[1], xn = p[0] + s * d[0], zn = p[2] + s * d[2], r;
    if (fabs(zn) <= w.dz && w.R(zn, r) && fabs(xn) <= r) return s;
  } else if (p[1] >= w.dy) {
    if (d[1] >= 0) return kBig;
    double s = (w.dy - p[1]) / d[1], xn = p[0] + s * d[0], zn = p[2] + s * d[2], r;
    if (fabs(zn) <= w.dz && w.R(zn, r) && fabs(xn) <= r) return s;
  }
  double dd[2];
  double s = winston_dist_to_parabola(w, p, d, 0., kPi);
  dd[0] = fabs(p[1] + s * d[1]) <= w.dy ? s : kBig;
  s = winston_dist_to_parabola(w, p, d, kPi, kPi);
  dd[1] = fabs(p[1] + s * d[1]) <= w.dy ? s : kBig;
  return std::min(dd[0], dd[1]);
}
void winston_normal(const Winston& w, const double* p, const double* d, double* n) {
  double x = p[0], y = p[1], z = p[2], r, dr;
  if (w.npoly) {  // Poly:67-122
    double saf[2];
    saf[0] = fabs(fabs(w.dz) - fabs(z));
    double phi = ATan2(y, x);
    while (phi > kPi / w.npoly) phi -= 2 * kPi / w.npoly;
    while (phi < -kPi / w.npoly) phi += 2 * kPi / w.npoly;
    saf[1] = w.R(z, r) ? fabs(r - sqrt(x * x + y * y) * cos(phi)) : kBig;
    if (!(saf[1] < saf[0])) { n[0] = 0; n[1] = 0; n[2] = 1; }
    else {
      phi = ATan2(y, x);
      if (phi < -kPi / w.npoly) phi += 2 * kPi;
      int k = (int)floor((phi + kPi / w.npoly) / (2 * kPi / w.npoly));
      if (!w.dRdZ(z, dr)) throw std::runtime_error("AGeoWinstonConePoly::ComputeNormal: CalcdRdZ threw");
      n[0] = cos(k * 2 * kPi / w.npoly);
      n[1] = sin(k * 2 * kPi / w.npoly);
      n[2] = -dr;
    }
  } else {  // 2D:120-172
    double saf[3] = {fabs(fabs(w.dy) - fabs(y)), fabs(fabs(w.dz) - fabs(z)), w.R(z, r) ? fabs(r - fabs(x)) : kBig};
    int i = saf[1] < saf[0] ? 1 : 0;
    if (saf[2] < saf[i]) i = 2;
    if (i == 0) { n[0] = 0; n[1] = 1; n[2] = 0; }
    else if (i == 1) { n[0] = 0; n[1] = 0; n[2] = 1; }
    else {
      if (!w.dRdZ(z, dr)) throw std::runtime_error("AGeoWinstonCone2D::ComputeNormal: CalcdRdZ threw");
      n[0] = 1; n[1] = 0; n[2] = x > 0 ? -dr : dr;
    }
  }
  double mag = sqrt(dot3(n, n));
  n[0] /= mag; n[1] /= mag; n[2] /= mag;
  if (dot3(n, d) < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
}

// ---- TGeoArb8 (ROOT, external; call sites tutorials/AshraOptics.C:264-284,403-441, src/AGeoUtil.cxx:47-82).
// Parity unpinned (no reference test traces through an Arb8).  Restated face by face: the section of the solid at height z is the
// quadrilateral of the vertices interpolated between the two z faces (TGeoArb8::Contains / InsidePolygon, vertices clockwise);
// a lateral face is the ruled surface swept by one edge of that quadrilateral.  A ray meets the surface of edge i where the point
// lies on the carrier line of the edge at the point's own height — a quadratic in the ray parameter whose coefficients are
// obtained here by sampling that condition at three parameters — and the hit counts when it lies between the edge's end points.
struct Arb8 {
  double dz, v[8][2];
  explicit Arb8(const double* P) : dz(P[0]) {
    for (int i = 0; i < 8; i++) { v[i][0] = P[1 + 2 * i]; v[i][1] = P[2 + 2 * i]; }
    double s1 = 0, s2 = 0;  // TGeoArb8::ComputeTwist re-orders counter-clockwise input
    for (int i = 0; i < 4; i++) {
      int j = (i + 1) % 4;
      s1 += v[i][0] * v[j][1] - v[j][0] * v[i][1];
      s2 += v[i + 4][0] * v[j + 4][1] - v[j + 4][0] * v[i + 4][1];
    }
    if (s1 > 1e-10 || s2 > 1e-10) {
      std::swap(v[1][0], v[3][0]); std::swap(v[1][1], v[3][1]);
      std::swap(v[5][0], v[7][0]); std::swap(v[5][1], v[7][1]);
    }
  }
  void section(double z, double q[4][2]) const {
    double cf = 0.5 * (dz - z) / dz;  // ROOT: poly = top + cf (bottom - top)
    for (int i = 0; i < 4; i++)
      for (int k = 0; k < 2; k++) q[i][k] = v[i + 4][k] + cf * (v[i][k] - v[i + 4][k]);
  }
  double side(int i, const double* x) const {  // > 0 on the inner side of edge i at the height of x
    double q[4][2];
    section(x[2], q);
    int j = (i + 1) % 4;
    return (x[0] - q[i][0]) * (q[j][1] - q[i][1]) - (x[1] - q[i][1]) * (q[j][0] - q[i][0]);
  }
};
bool arb8_contains(const Arb8& a, const double* p) {
  if (fabs(p[2]) > a.dz) return false;
  for (int i = 0; i < 4; i++)
    if (a.side(i, p) < 0) return false;
  return true;
}
// boundary crossings of the ray that lie on the solid's surface, ascending
void arb8_hits(const Arb8& a, const double* p, const double* d, std::vector<double>& hits) {
  hits.clear();
  auto at = [&](double t, double* x) { for (int k = 0; k < 3; k++) x[k] = p[k] + t * d[k]; };
  for (int sgn = -1; sgn <= 1; sgn += 2) {  // z faces
    if (d[2] == 0) continue;
    double t = (sgn * a.dz - p[2]) / d[2], x[3];
    if (!(t > 1e-11 && t < 1e29)) continue;
    at(t, x);
    x[2] = sgn * a.dz;
    bool in = true;
    for (int i = 0; i < 4; i++)
      if (a.side(i, x) < -1e-9) in = false;
    if (in) hits.push_back(t);
  }
  for (int i = 0; i < 4; i++) {
    double x0[3], x1[3], x2[3];
    at(0, x0); at(1, x1); at(2, x2);
    double f0 = a.side(i, x0), f1 = a.side(i, x1), f2 = a.side(i, x2);
    double c0 = f0, c2 = 0.5 * (f2 - 2 * f1 + f0), c1 = f1 - f0 - c2, r[2];
    int nr = 0;
    double scale = fabs(c0) + fabs(c1) + fabs(c2);
    if (scale == 0) continue;  // degenerate edge (coinciding vertices): no face
    if (fabs(c2) < 1e-13 * scale) {
      if (c1 != 0) r[nr++] = -c0 / c1;
    } else {
      double disc = c1 * c1 - 4 * c2 * c0;
      if (disc < 0) continue;
      double sq_ = sqrt(disc), q = -0.5 * (c1 + (c1 >= 0 ? sq_ : -sq_));
      r[nr++] = q / c2;
      if (q != 0) r[nr++] = c0 / q;
    }
    for (int k = 0; k < nr; k++) {
      double t = r[k], x[3], q[4][2];
      if (!(t > 1e-11 && t < 1e29)) continue;
      at(t, x);
      if (fabs(x[2]) > a.dz + 1e-9) continue;
      a.section(x[2] > a.dz ? a.dz : (x[2] < -a.dz ? -a.dz : x[2]), q);
      int j = (i + 1) % 4;
      double ex = q[j][0] - q[i][0], ey = q[j][1] - q[i][1], len2 = ex * ex + ey * ey;
      if (len2 < 1e-20) continue;
      double fr = ((x[0] - q[i][0]) * ex + (x[1] - q[i][1]) * ey) / len2;
      if (fr < -1e-9 || fr > 1 + 1e-9) continue;
      hits.push_back(t);
    }
  }
  std::sort(hits.begin(), hits.end());
}
double arb8_dist_in(const Arb8& a, const double* p, const double* d) {
  std::vector<double> h;
  arb8_hits(a, p, d, h);
  for (size_t k = 0; k < h.size(); k++) {  // leave through the first surface point behind which the ray is outside
    double t = 0.5 * (h[k] + (k + 1 < h.size() ? h[k + 1] : h[k] + 1.)), x[3] = {p[0] + t * d[0], p[1] + t * d[1], p[2] + t * d[2]};
    if (k + 1 < h.size() && h[k + 1] - h[k] < 1e-12) continue;
    if (!arb8_contains(a, x)) return h[k];
  }
  return h.empty() ? 0. : h.back();
}
double arb8_dist_out(const Arb8& a, const double* p, const double* d) {
  std::vector<double> h;
  arb8_hits(a, p, d, h);
  if (!h.empty()) {  // already inside before the first surface point
    double t = 0.5 * h[0], x[3] = {p[0] + t * d[0], p[1] + t * d[1], p[2] + t * d[2]};
    if (h[0] >= 1e-12 && arb8_contains(a, x)) return 0.;
  }
  for (size_t k = 0; k < h.size(); k++) {
    if (k + 1 < h.size() && h[k + 1] - h[k] < 1e-12) continue;
    double t = 0.5 * (h[k] + (k + 1 < h.size() ? h[k + 1] : h[k] + 1.)), x[3] = {p[0] + t * d[0], p[1] + t * d[1], p[2] + t * d[2]};
    if (arb8_contains(a, x)) return h[k];
  }
  return kBig;
}
// TGeoArb8::ComputeNormal: z face within 10 tolerances; else the face of the closest edge of the section, normal = edge x ruling
void arb8_normal(const Arb8& a, const double* p, const double* d, double* n) {
  if (a.dz - fabs(p[2]) < 10 * kTol) { n[0] = n[1] = 0; n[2] = d[2] >= 0 ? 1 : -1; return; }
  double z = p[2] > a.dz ? a.dz : (p[2] < -a.dz ? -a.dz : p[2]), q[4][2];
  a.section(z, q);
  double best = kBig, frac = 0;
  int iseg = 0;
  for (int i = 0; i < 4; i++) {
    int j = (i + 1) % 4;
    double ex = q[j][0] - q[i][0], ey = q[j][1] - q[i][1], len2 = ex * ex + ey * ey, ux = p[0] - q[i][0], uy = p[1] - q[i][1], f = 0, d2;
    bool degenerate = len2 < 1e-20;
    if (degenerate) d2 = ux * ux + uy * uy;
    else {
      f = (ux * ex + uy * ey) / len2;
      f = std::min(1., std::max(0., f));
      d2 = sq(ux - f * ex) + sq(uy - f * ey);
    }
    if (d2 < best - (degenerate ? 0. : 1e-24)) { best = d2; iseg = i; frac = f; }
  }
  int j = (iseg + 1) % 4;
  double ex = q[j][0] - q[iseg][0], ey = q[j][1] - q[iseg][1];
  if (ex * ex + ey * ey < 1e-20) {  // apex: take the edge direction from the opposite z face
    double q2[4][2];
    a.section(z < 0 ? a.dz : -a.dz, q2);
    ex = q2[j][0] - q2[iseg][0]; ey = q2[j][1] - q2[iseg][1];
  }
  double rx = (1 - frac) * (a.v[iseg + 4][0] - a.v[iseg][0]) + frac * (a.v[j + 4][0] - a.v[j][0]);
  double ry = (1 - frac) * (a.v[iseg + 4][1] - a.v[iseg][1]) + frac * (a.v[j + 4][1] - a.v[j][1]);
  double rz = 2 * a.dz;
  n[0] = ey * rz; n[1] = -ex * rz; n[2] = ex * ry - ey * rx;
  double mag = sqrt(dot3(n, n));
  if (!(mag > 0)) { n[0] = n[1] = 0; n[2] = d[2] >= 0 ? 1 : -1; return; }
  for (int k = 0; k < 3; k++) n[k] /= mag;
  if (dot3(n, d) < 0) for (int k = 0; k < 3; k++) n[k] = -n[k];
}

// ---- TGeoXtru (ROOT, external; call sites tutorials/AshraOptics.C:791-1021, src/AGeoUtil.cxx:84-125).  Parity unpinned.
// The outline (any simple polygon) is placed per section at (x0,y0) with a scale, interpolated linearly between sections
// (TGeoXtru::SetCurrentZ); a lateral face is the planar trapezoid between the copies of one edge on two consecutive sections.
struct Xtru {
  int nv, nz;
  const double *V, *sec;
  explicit Xtru(const double* P) : nv((int)P[0]), nz((int)P[1]), V(P + 2), sec(P + 2 + 2 * (int)P[0]) {}
  double z(int k) const { return sec[4 * k]; }
  void frame(int k, double f, double& x0, double& y0, double& sc) const {  // between sections k and k+1
    x0 = sec[4 * k + 1] + f * (sec[4 * k + 5] - sec[4 * k + 1]);
    y0 = sec[4 * k + 2] + f * (sec[4 * k + 6] - sec[4 * k + 2]);
    sc = sec[4 * k + 3] + f * (sec[4 * k + 7] - sec[4 * k + 3]);
  }
  bool in_outline(double x, double y) const {  // winding number
    int wn = 0;
    for (int k = 0; k < nv; k++) {
      int j = (k + 1) % nv;
      double x1 = V[2 * k], y1 = V[2 * k + 1], x2 = V[2 * j], y2 = V[2 * j + 1];
      double left = (x2 - x1) * (y - y1) - (x - x1) * (y2 - y1);
      if (y1 <= y) { if (y2 > y && left > 0) wn++; }
      else if (y2 <= y && left < 0) wn--;
    }
    return wn != 0;
  }
};
bool xtru_contains(const Xtru& X, const double* p) {
  if (p[2] < X.z(0) || p[2] > X.z(X.nz - 1)) return false;
  int k = 0;
  for (int i = 0; i + 1 < X.nz; i++)
    if (X.z(i) <= p[2]) k = i;
  double dzs = X.z(k + 1) - X.z(k), f = dzs > 1e-8 ? (p[2] - X.z(k)) / dzs : 0., x0, y0, sc;
  X.frame(k, f, x0, y0, sc);
  if (!(sc > 0)) return false;
  return X.in_outline((p[0] - x0) / sc, (p[1] - y0) / sc);
}
void xtru_hits(const Xtru& X, const double* p, const double* d, std::vector<double>& hits) {
  hits.clear();
  if (d[2] != 0)
    for (int i = 0; i < X.nz; i++) {  // section planes (end faces, outline jumps; a plane inside the solid is classified away later)
      double t = (X.z(i) - p[2]) / d[2];
      if (t > 1e-11 && t < 1e29) hits.push_back(t);
    }
  for (int s = 0; s + 1 < X.nz; s++) {
    double z0 = X.z(s), z1 = X.z(s + 1);
    if (z1 - z0 < 1e-8) continue;
    double xa, ya, sa, xb, yb, sb;
    X.frame(s, 0., xa, ya, sa);
    X.frame(s, 1., xb, yb, sb);
    for (int k = 0; k < X.nv; k++) {
      int j = (k + 1) % X.nv;
      // plane through A0, B0 (lower copies of the edge's end points) and A1 (upper copy of the first)
      double A0[3] = {xa + sa * X.V[2 * k], ya + sa * X.V[2 * k + 1], z0}, B0[3] = {xa + sa * X.V[2 * j], ya + sa * X.V[2 * j + 1], z0};
      double A1[3] = {xb + sb * X.V[2 * k], yb + sb * X.V[2 * k + 1], z1};
      double e1[3] = {B0[0] - A0[0], B0[1] - A0[1], 0.}, e2[3] = {A1[0] - A0[0], A1[1] - A0[1], z1 - z0};
      double nn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
      double den = dot3(nn, d);
      if (den == 0) continue;
      double w[3] = {A0[0] - p[0], A0[1] - p[1], A0[2] - p[2]}, t = dot3(nn, w) / den;
      if (!(t > 1e-11 && t < 1e29)) continue;
      double zz = p[2] + t * d[2];
      if (zz < z0 - 1e-9 || zz > z1 + 1e-9) continue;
      hits.push_back(t);
    }
  }
  std::sort(hits.begin(), hits.end());
}
double xtru_dist(const Xtru& X, const double* p, const double* d, bool from_inside) {
  std::vector<double> h;
  xtru_hits(X, p, d, h);
  double prev = 0;
  for (size_t k = 0; k < h.size(); k++) {
    if (h[k] - prev < 1e-12) { prev = h[k]; continue; }
    double t = 0.5 * (prev + h[k]), x[3] = {p[0] + t * d[0], p[1] + t * d[1], p[2] + t * d[2]};
    bool in = xtru_contains(X, x);
    if (from_inside ? !in : in) return prev;
    prev = h[k];
  }
  return from_inside ? prev : kBig;
}
void xtru_normal(const Xtru& X, const double* p, const double* d, double* n) {
  double best = kBig;
  n[0] = n[1] = 0; n[2] = 1;
  for (int i = 0; i < X.nz; i++) {
    bool cap = i == 0 || i == X.nz - 1;
    bool jump = (i + 1 < X.nz && X.z(i + 1) - X.z(i) < 1e-8) || (i > 0 && X.z(i) - X.z(i - 1) < 1e-8);
    if (!cap && !jump) continue;
    double s = fabs(p[2] - X.z(i));
    if (s < best) { best = s; n[0] = n[1] = 0; n[2] = 1; }
  }
  for (int s = 0; s + 1 < X.nz; s++) {
    double z0 = X.z(s), z1 = X.z(s + 1), dzs = z1 - z0;
    if (dzs < 1e-8 || p[2] < z0 - 1e-6 || p[2] > z1 + 1e-6) continue;
    double x0, y0, sc, xa, ya, sa, xb, yb, sb;
    X.frame(s, (p[2] - z0) / dzs, x0, y0, sc);
    X.frame(s, 0., xa, ya, sa);
    X.frame(s, 1., xb, yb, sb);
    for (int k = 0; k < X.nv; k++) {
      int j = (k + 1) % X.nv;
      double ax = x0 + sc * X.V[2 * k], ay = y0 + sc * X.V[2 * k + 1], ex = sc * (X.V[2 * j] - X.V[2 * k]), ey = sc * (X.V[2 * j + 1] - X.V[2 * k + 1]);
      double len2 = ex * ex + ey * ey;
      if (len2 < 1e-20) continue;
      double ux = p[0] - ax, uy = p[1] - ay, fr = std::min(1., std::max(0., (ux * ex + uy * ey) / len2));
      double foot = sqrt(sq(ux - fr * ex) + sq(uy - fr * ey));
      // direction in which vertex k travels per unit z
      double rx = ((xb + sb * X.V[2 * k]) - (xa + sa * X.V[2 * k])) / dzs, ry = ((yb + sb * X.V[2 * k + 1]) - (ya + sa * X.V[2 * k + 1])) / dzs;
      double fn[3] = {ey, -ex, ex * ry - ey * rx}, mag = sqrt(dot3(fn, fn)), dist = foot * sqrt(len2) / mag;
      if (dist < best) { best = dist; for (int q = 0; q < 3; q++) n[q] = fn[q] / mag; }
    }
  }
  if (dot3(n, d) < 0) for (int k = 0; k < 3; k++) n[k] = -n[k];
}

// ---- dispatch + TGeoBoolNode algorithms (TGeoUnion / TGeoIntersection / TGeoSubtraction, ROOT 6)
bool contains(const Scene& S, int sh, const double* p);
double dist_in(const Scene& S, int sh, const double* p, const double* d, int* sel);
double dist_out(const Scene& S, int sh, const double* p, const double* d, double step, int* sel);

bool contains(const Scene& S, int sh, const double* p) {
  const rbg_shape& s = S.d->shapes[sh];
  const double* P = S.d->dpar + s.ipar;
  switch (s.type) {
    case RBG_SHAPE_BBOX: return bbox_contains(P, p);
    case RBG_SHAPE_TUBE: return tube_contains(P, p);
    case RBG_SHAPE_SPHERE: return sphere_contains(P, p);
    case RBG_SHAPE_PARABOLOID: return para_contains(Para(P), p);
    case RBG_SHAPE_PGON: return poly_contains(Poly(P, true), p);
    case RBG_SHAPE_PCON: return poly_contains(Poly(P, false), p);
    case RBG_SHAPE_ASPHERE: return asph_contains(Asph(P), p);
    case RBG_SHAPE_WINSTON2D: return winston_contains(Winston(P, false), p);
    case RBG_SHAPE_WINSTONPOLY: return winston_contains(Winston(P, true), p);
    case RBG_SHAPE_ARB8: return arb8_contains(Arb8(P), p);
    case RBG_SHAPE_XTRU: return xtru_contains(Xtru(P), p);
    default: break;
  }
  double l[3], r[3];
  m2l(S.mat(s.lmat), p, l);
  m2l(S.mat(s.rmat), p, r);
  if (s.type == RBG_SHAPE_UNION) return contains(S, s.left, l) || contains(S, s.right, r);
  if (s.type == RBG_SHAPE_INTERSECTION) return contains(S, s.left, l) && contains(S, s.right, r);
  return contains(S, s.left, l) && !contains(S, s.right, r);
}

double dist_in(const Scene& S, int sh, const double* p, const double* d, int* sel) {
  const rbg_shape& s = S.d->shapes[sh];
  const double* P = S.d->dpar + s.ipar;
  *sel = 0;
  switch (s.type) {
    case RBG_SHAPE_BBOX: return bbox_dist_in(P, p, d);
    case RBG_SHAPE_TUBE: return tube_dist_in(P[0], P[1], P[2], p, d);
    case RBG_SHAPE_SPHERE: return sphere_dist(P, p, d, true);
    case RBG_SHAPE_PARABOLOID: return para_dist_in(Para(P), p, d);
    case RBG_SHAPE_PGON: return poly_dist(Poly(P, true), p, d, true);
    case RBG_SHAPE_PCON: return poly_dist(Poly(P, false), p, d, true);
    case RBG_SHAPE_ASPHERE: return asph_dist4(Asph(P), p, d);
    case RBG_SHAPE_WINSTON2D: return winston_dist_in(Winston(P, false), p, d);
    case RBG_SHAPE_WINSTONPOLY: return winston_dist_in(Winston(P, true), p, d);
    case RBG_SHAPE_ARB8: return arb8_dist_in(Arb8(P), p, d);
    case RBG_SHAPE_XTRU: return xtru_dist(Xtru(P), p, d, true);
    default: break;
  }
  Mat ML = S.mat(s.lmat), MR = S.mat(s.rmat);
  double ll[3], lr[3], ldir[3], rdir[3];
  m2lv(ML, d, ldir);
  m2lv(MR, d, rdir);
  m2l(ML, p, ll);
  m2l(MR, p, lr);
  int s1 = 0, s2 = 0;
  if (s.type == RBG_SHAPE_INTERSECTION) {
    double d1 = dist_in(S, s.left, ll, ldir, &s1), d2 = dist_in(S, s.right, lr, rdir, &s2);
    if (d1 < d2) { *sel = 1 | (s1 << 2); return d1; }
    *sel = 2 | (s2 << 2);
    return d2;
  }
  if (s.type == RBG_SHAPE_SUBTRACTION) {
    double d1 = dist_in(S, s.left, ll, ldir, &s1), d2 = dist_out(S, s.right, lr, rdir, kBig, &s2);
    if (d1 < d2) { *sel = 1 | (s1 << 2); return d1; }
    *sel = 2 | (s2 << 2);
    return d2;
  }
  // TGeoUnion::DistFromInside
  double master[3] = {p[0], p[1], p[2]}, pushed[3], local[3], d1 = 0., d2 = 0., snxt = 0., eps = 0.;
  bool inside1 = contains(S, s.left, ll), inside2 = contains(S, s.right, lr);
  if (inside1) d1 = dist_in(S, s.left, ll, ldir, &s1);
  if (inside2) d2 = dist_in(S, s.right, lr, rdir, &s2);
  if (!(inside1 || inside2)) {
    d1 = dist_out(S, s.left, ll, ldir, kBig, &s1);
    if (d1 < 2. * kTol) {
      eps = d1 + kTol;
      for (int i = 0; i < 3; i++) ll[i] += eps * ldir[i];
      inside1 = true;
      d1 = dist_in(S, s.left, ll, ldir, &s1) + eps;
    } else {
      d2 = dist_out(S, s.right, lr, rdir, kBig, &s2);
      if (d2 < 2. * kTol) {
        eps = d2 + kTol;
        for (int i = 0; i < 3; i++) lr[i] += eps * rdir[i];
        inside2 = true;
        d2 = dist_in(S, s.right, lr, rdir, &s2) + eps;
      }
    }
  }
  while (inside1 || inside2) {
    bool take1 = inside1 && (!inside2 || d1 < d2);
    if (take1) {
      snxt += d1;
      *sel = 1 | (s1 << 2);
      inside1 = false;
      for (int i = 0; i < 3; i++) { master[i] += d1 * d[i]; pushed[i] = master[i] + (1. + d1) * kTol * d[i]; }
      m2l(MR, pushed, local);
      inside2 = contains(S, s.right, local);
      if (!inside2) return snxt;
      d2 = dist_in(S, s.right, local, rdir, &s2);
      if (d2 < kTol) return snxt;
      d2 += (1. + d1) * kTol;
    } else {
      snxt += d2;
      *sel = 2 | (s2 << 2);
      inside2 = false;
      for (int i = 0; i < 3; i++) { master[i] += d2 * d[i]; pushed[i] = master[i] + (1. + d2) * kTol * d[i]; }
      m2l(ML, pushed, local);
      inside1 = contains(S, s.left, local);
      if (!inside1) return snxt;
      d1 = dist_in(S, s.left, local, ldir, &s1);
      if (d1 < kTol) return snxt;
      d1 += (1. + d2) * kTol;
    }
  }
  return snxt;
}

double dist_out(const Scene& S, int sh, const double* p, const double* d, double step, int* sel) {
  const rbg_shape& s = S.d->shapes[sh];
  const double* P = S.d->dpar + s.ipar;
  *sel = 0;
  switch (s.type) {
    case RBG_SHAPE_BBOX: return bbox_dist_out(P, p, d, step);
    case RBG_SHAPE_TUBE: return tube_dist_out(P[0], P[1], P[2], p, d);
    case RBG_SHAPE_SPHERE: return sphere_dist(P, p, d, false);
    case RBG_SHAPE_PARABOLOID: return para_dist_out(Para(P), p, d);
    case RBG_SHAPE_PGON: return poly_dist(Poly(P, true), p, d, false);
    case RBG_SHAPE_PCON: return poly_dist(Poly(P, false), p, d, false);
    case RBG_SHAPE_ASPHERE: return asph_dist_out(Asph(P), p, d, step);
    case RBG_SHAPE_WINSTON2D: return winston_dist_out(Winston(P, false), p, d);
    case RBG_SHAPE_WINSTONPOLY: return winston_dist_out(Winston(P, true), p, d);
    case RBG_SHAPE_ARB8: return arb8_dist_out(Arb8(P), p, d);
    case RBG_SHAPE_XTRU: return xtru_dist(Xtru(P), p, d, false);
    default: break;
  }
  Mat ML = S.mat(s.lmat), MR = S.mat(s.rmat);
  double ll[3], lr[3], ldir[3], rdir[3];
  m2lv(ML, d, ldir);
  m2lv(MR, d, rdir);
  m2l(ML, p, ll);
  m2l(MR, p, lr);
  int s1 = 0, s2 = 0;
  if (s.type == RBG_SHAPE_UNION) {
    double d1 = dist_out(S, s.left, ll, ldir, step, &s1), d2 = dist_out(S, s.right, lr, rdir, step, &s2);
    if (d1 < d2) { *sel = 1 | (s1 << 2); return d1; }
    *sel = 2 | (s2 << 2);
    return d2;
  }
  double master[3] = {p[0], p[1], p[2]};
  if (s.type == RBG_SHAPE_INTERSECTION) {
    bool inleft = contains(S, s.left, ll), inright = contains(S, s.right, lr);
    double snext = 0.0, d1, d2;
    if (inleft && inright) {
      d1 = dist_in(S, s.left, ll, ldir, &s1);
      d2 = dist_in(S, s.right, lr, rdir, &s2);
      if (d1 < 1.E-3) inleft = false;
      if (d2 < 1.E-3) inright = false;
      if (inleft && inright) return snext;
    }
    for (int guard = 0; guard < 1000; guard++) {
      d1 = d2 = 0;
      if (!inleft) {
        d1 = std::max(dist_out(S, s.left, ll, ldir, kBig, &s1), kTol);
        if (d1 > 1E20) return kBig;
      }
      if (!inright) {
        d2 = std::max(dist_out(S, s.right, lr, rdir, kBig, &s2), kTol);
        if (d2 > 1E20) return kBig;
      }
      if (d1 > d2) {
        snext += d1;
        *sel = 1 | (s1 << 2);
        inleft = true;
        for (int i = 0; i < 3; i++) master[i] += d1 * d[i];
        m2l(ML, master, ll);
        m2l(MR, master, lr);
        for (int i = 0; i < 3; i++) lr[i] += kTol * rdir[i];
        inright = contains(S, s.right, lr);
        if (inright) return snext;
      } else {
        snext += d2;
        *sel = 2 | (s2 << 2);
        inright = true;
        for (int i = 0; i < 3; i++) master[i] += d2 * d[i];
        m2l(MR, master, lr);
        m2l(ML, master, ll);
        for (int i = 0; i < 3; i++) ll[i] += kTol * ldir[i];
        inleft = contains(S, s.left, ll);
        if (inleft) return snext;
      }
    }
    return kBig;
  }
  // TGeoSubtraction::DistFromOutside
  bool inside = contains(S, s.right, lr);
  double snxt = 0., epsil = 0., d1, d2;
  for (int guard = 0; guard < 1000; guard++) {
    if (inside) {
      d1 = dist_in(S, s.right, lr, rdir, &s2);
      *sel = 2 | (s2 << 2);
      snxt += d1 + epsil;
      for (int i = 0; i < 3; i++) master[i] += (d1 + 1E-8) * d[i];
      epsil = 1.E-8;
      m2l(ML, master, ll);
      if (contains(S, s.left, ll)) return snxt;
    }
    m2l(ML, master, ll);
    d2 = dist_out(S, s.left, ll, ldir, kBig, &s1);
    if (d2 > 1E20) return kBig;
    m2l(MR, master, lr);
    int s2b = 0;
    d1 = dist_out(S, s.right, lr, rdir, kBig, &s2b);
    if (d2 < d1 - kTol) {
      snxt += d2 + epsil;
      *sel = 1 | (s1 << 2);
      return snxt;
    }
    snxt += d1 + epsil;
    for (int i = 0; i < 3; i++) master[i] += (d1 + 1E-8) * d[i];
    epsil = 1.E-8;
    m2l(MR, master, lr);
    inside = true;
  }
  return kBig;
}

// normal of the primitive that produced the selected boundary (TGeoBoolNode::fSelected semantics)
void normal(const Scene& S, int sh, const double* p, const double* d, int sel, double* n) {
  const rbg_shape& s = S.d->shapes[sh];
  const double* P = S.d->dpar + s.ipar;
  switch (s.type) {
    case RBG_SHAPE_BBOX: bbox_normal(P, p, d, n); return;
    case RBG_SHAPE_TUBE: tube_normal(P, p, d, n); return;
    case RBG_SHAPE_SPHERE: sphere_normal(P, p, d, n); return;
    case RBG_SHAPE_PARABOLOID: para_normal(Para(P), p, d, n); return;
    case RBG_SHAPE_PGON: poly_normal(Poly(P, true), p, d, n); return;
    case RBG_SHAPE_PCON: poly_normal(Poly(P, false), p, d, n); return;
    case RBG_SHAPE_ASPHERE: asph_normal(Asph(P), p, d, n); return;
    case RBG_SHAPE_WINSTON2D: winston_normal(Winston(P, false), p, d, n); return;
    case RBG_SHAPE_WINSTONPOLY: winston_normal(Winston(P, true), p, d, n); return;
    case RBG_SHAPE_ARB8: arb8_normal(Arb8(P), p, d, n); return;
    case RBG_SHAPE_XTRU: xtru_normal(Xtru(P), p, d, n); return;
    default: break;
  }
  int side = sel & 3;
  if (side == 0) {  // no distance call selected an operand: decide geometrically like TGeoBoolNode::ComputeNormal
    double l[3], r[3];
    m2l(S.mat(s.lmat), p, l);
    m2l(S.mat(s.rmat), p, r);
    bool inl = contains(S, s.left, l), inr = contains(S, s.right, r);
    if (s.type == RBG_SHAPE_SUBTRACTION) side = inr ? 2 : 1;
    else if (s.type == RBG_SHAPE_UNION) side = inl ? (inr ? 1 : 1) : 2;
    else side = inl ? 2 : 1;
  }
  Mat M = S.mat(side == 1 ? s.lmat : s.rmat);
  double lp[3], ld[3], ln[3];
  m2l(M, p, lp);
  m2lv(M, d, ld);
  normal(S, side == 1 ? s.left : s.right, lp, ld, sel >> 2, ln);
  l2mv(M, ln, n);
}

// ================================================================== navigator (TGeoNavigator restated)
// ================================================================== daughter boxes (see VolVox)
// axis-aligned bounding box of a shape in its own frame
static void shape_box(const Scene& S, int sh, double* lo, double* hi) {
  const rbg_shape& s = S.d->shapes[sh];
  const double* P = S.d->dpar + s.ipar;
  auto sym = [&](double x, double y, double z) { lo[0] = -x; lo[1] = -y; lo[2] = -z; hi[0] = x; hi[1] = y; hi[2] = z; };
  switch (s.type) {
    case RBG_SHAPE_BBOX:
      for (int i = 0; i < 3; i++) { lo[i] = P[3 + i] - P[i]; hi[i] = P[3 + i] + P[i]; }
      return;
    case RBG_SHAPE_TUBE: sym(P[1], P[1], P[2]); return;
    case RBG_SHAPE_SPHERE: sym(P[1], P[1], P[1]); return;
    case RBG_SHAPE_PARABOLOID: { double r = std::max(P[0], P[1]); sym(r, r, P[2]); return; }
    case RBG_SHAPE_PGON:
    case RBG_SHAPE_PCON: {
      const bool pg = s.type == RBG_SHAPE_PGON;
      int nz = (int)P[pg ? 3 : 2];
      const double* sec = P + (pg ? 4 : 3);
      double rmax = 0, z0 = kBig, z1 = -kBig;
      for (int i = 0; i < nz; i++) { rmax = std::max(rmax, sec[3 * i + 2]); z0 = std::min(z0, sec[3 * i]); z1 = std::max(z1, sec[3 * i]); }
      if (pg) rmax /= cos(0.5 * P[1] / P[2] * kPi / 180.);  // rmax of a polygon section is the apothem
      sym(rmax, rmax, 0);
      lo[2] = z0; hi[2] = z1;
      return;
    }
    case RBG_SHAPE_ASPHERE: sym(P[7], P[7], 0); lo[2] = P[10] - P[11]; hi[2] = P[10] + P[11]; return;
    case RBG_SHAPE_WINSTON2D:
    case RBG_SHAPE_WINSTONPOLY: {
      double th = asin(P[1] / P[0]), dz = (P[0] + P[1]) / (2 * tan(th));
      if (s.type == RBG_SHAPE_WINSTON2D) sym(P[0], P[2], dz);
      else { double r = P[0] / cos(kPi / P[2]); sym(r, r, dz); }
      return;
    }
    case RBG_SHAPE_ARB8: {
      double x = 0, y = 0;
      for (int i = 0; i < 8; i++) { x = std::max(x, fabs(P[1 + 2 * i])); y = std::max(y, fabs(P[2 + 2 * i])); }
      sym(x, y, P[0]);
      return;
    }
    case RBG_SHAPE_XTRU: {
      int nv = (int)P[0], nz = (int)P[1];
      const double* V = P + 2;
      const double* Z = P + 2 + 2 * nv;
      for (int i = 0; i < 3; i++) { lo[i] = kBig; hi[i] = -kBig; }
      for (int k = 0; k < nz; k++) {
        lo[2] = std::min(lo[2], Z[4 * k]); hi[2] = std::max(hi[2], Z[4 * k]);
        for (int i = 0; i < nv; i++) {
          double x = Z[4 * k + 1] + Z[4 * k + 3] * V[2 * i], y = Z[4 * k + 2] + Z[4 * k + 3] * V[2 * i + 1];
          lo[0] = std::min(lo[0], x); hi[0] = std::max(hi[0], x); lo[1] = std::min(lo[1], y); hi[1] = std::max(hi[1], y);
        }
      }
      return;
    }
    default: break;
  }
  // booleans: the left operand bounds a subtraction; an intersection lies in both operands' boxes; a union needs both
  auto operand = [&](int sub, int m, double* l, double* h) {
    double a[3], b[3];
    shape_box(S, sub, a, b);
    Mat M = S.mat(m);
    for (int i = 0; i < 3; i++) { l[i] = kBig; h[i] = -kBig; }
    for (int c = 0; c < 8; c++) {
      double q[3] = {c & 1 ? b[0] : a[0], c & 2 ? b[1] : a[1], c & 4 ? b[2] : a[2]}, w[3];
      l2m(M, q, w);
      for (int i = 0; i < 3; i++) { l[i] = std::min(l[i], w[i]); h[i] = std::max(h[i], w[i]); }
    }
  };
  operand(s.left, s.lmat, lo, hi);
  if (s.type == RBG_SHAPE_UNION) {
    double l2[3], h2[3];
    operand(s.right, s.rmat, l2, h2);
    for (int i = 0; i < 3; i++) { lo[i] = std::min(lo[i], l2[i]); hi[i] = std::max(hi[i], h2[i]); }
  } else if (s.type == RBG_SHAPE_INTERSECTION) {
    double l2[3], h2[3];
    operand(s.right, s.rmat, l2, h2);
    for (int i = 0; i < 3; i++) { lo[i] = std::max(lo[i], l2[i]); hi[i] = std::min(hi[i], h2[i]); }
  }
}
static int vox_split(VolVox& V, int first, int count) {
  VoxNode nd;
  for (int i = 0; i < 3; i++) { nd.lo[i] = kBig; nd.hi[i] = -kBig; }
  for (int k = first; k < first + count; k++) {
    const double* b = &V.box[6 * V.order[k]];
    for (int i = 0; i < 3; i++) { nd.lo[i] = std::min(nd.lo[i], b[i]); nd.hi[i] = std::max(nd.hi[i], b[3 + i]); }
  }
  nd.left = nd.right = -1;
  nd.first = first;
  nd.count = count;
  int id = (int)V.tree.size();
  V.tree.push_back(nd);
  if (count > 4) {
    int ax = 0;
    for (int i = 1; i < 3; i++)
      if (nd.hi[i] - nd.lo[i] > nd.hi[ax] - nd.lo[ax]) ax = i;
    std::sort(V.order.begin() + first, V.order.begin() + first + count,
              [&](int a, int b) { return V.box[6 * a + ax] + V.box[6 * a + 3 + ax] < V.box[6 * b + ax] + V.box[6 * b + 3 + ax]; });
    int l = vox_split(V, first, count / 2), r = vox_split(V, first + count / 2, count - count / 2);
    V.tree[id].left = l;
    V.tree[id].right = r;
  }
  return id;
}
static void build_voxels(Scene& S) {
  S.vox.assign(S.d->nvolumes, VolVox());
  for (int vi = 0; vi < S.d->nvolumes; vi++) {
    const rbg_volume& v = S.d->volumes[vi];
    if (v.nnodes < 4) continue;  // nothing to gain below a handful of daughters
    VolVox& V = S.vox[vi];
    V.box.resize(6 * (size_t)v.nnodes);
    for (int k = 0; k < v.nnodes; k++) {
      const rbg_node& nd = S.d->nodes[v.first_node + k];
      double a[3], b[3];
      shape_box(S, S.d->volumes[nd.volume].shape, a, b);
      Mat M = S.mat(nd.matrix);
      double* o = &V.box[6 * k];
      for (int i = 0; i < 3; i++) { o[i] = kBig; o[3 + i] = -kBig; }
      for (int c = 0; c < 8; c++) {
        double q[3] = {c & 1 ? b[0] : a[0], c & 2 ? b[1] : a[1], c & 4 ? b[2] : a[2]}, w[3];
        l2m(M, q, w);
        for (int i = 0; i < 3; i++) { o[i] = std::min(o[i], w[i]); o[3 + i] = std::max(o[3 + i], w[i]); }
      }
      for (int i = 0; i < 3; i++) {  // padding: far above every tolerance the shape algorithms use (1e-8 nudges, 1e-6 steps)
        double pad = 1e-4 + 1e-9 * (fabs(o[i]) + fabs(o[3 + i]));
        o[i] -= pad;
        o[3 + i] += pad;
      }
      V.order.push_back(k);
    }
    vox_split(V, 0, v.nnodes);
  }
}
// daughters (ascending index) whose box the ray p + t d meets for some 0 <= t <= tmax
static void vox_ray(const VolVox& V, const double* p, const double* d, double tmax, std::vector<int>& out) {
  out.clear();
  int stack[64], sp = 0;
  stack[sp++] = 0;
  auto hit = [&](const double* lo, const double* hi) {
    double t0 = 0, t1 = tmax;
    for (int i = 0; i < 3; i++) {
      if (d[i] != 0) {
        double a = (lo[i] - p[i]) / d[i], b = (hi[i] - p[i]) / d[i];
        if (a > b) std::swap(a, b);
        if (a > t0) t0 = a;
        if (b < t1) t1 = b;
      } else if (p[i] < lo[i] || p[i] > hi[i]) return false;
    }
    return t0 <= t1;
  };
  while (sp > 0) {
    const VoxNode& n = V.tree[stack[--sp]];
    if (!hit(n.lo, n.hi)) continue;
    if (n.left < 0) {
      for (int k = n.first; k < n.first + n.count; k++) {
        const double* b = &V.box[6 * V.order[k]];
        if (hit(b, b + 3)) out.push_back(V.order[k]);
      }
    } else { stack[sp++] = n.left; stack[sp++] = n.right; }
  }
  std::sort(out.begin(), out.end());
}
// daughters (ascending index) whose box holds the point
static void vox_point(const VolVox& V, const double* p, std::vector<int>& out) {
  out.clear();
  int stack[64], sp = 0;
  stack[sp++] = 0;
  auto in = [&](const double* lo, const double* hi) { return p[0] >= lo[0] && p[0] <= hi[0] && p[1] >= lo[1] && p[1] <= hi[1] && p[2] >= lo[2] && p[2] <= hi[2]; };
  while (sp > 0) {
    const VoxNode& n = V.tree[stack[--sp]];
    if (!in(n.lo, n.hi)) continue;
    if (n.left < 0) {
      for (int k = n.first; k < n.first + n.count; k++) {
        const double* b = &V.box[6 * V.order[k]];
        if (in(b, b + 3)) out.push_back(V.order[k]);
      }
    } else { stack[sp++] = n.left; stack[sp++] = n.right; }
  }
  std::sort(out.begin(), out.end());
}

// ================================================================== navigator
const int kMaxLevel = 16;
struct Nav {
  const Scene* S;
  int level;                 // -1: outside the top volume
  int node[kMaxLevel];       // index into desc.nodes for level>=1 (level 0 = top volume)
  int vol[kMaxLevel];
  Mat glob[kMaxLevel];       // local(level) -> master
  double P[3], D[3];
  double step;
  bool on_boundary;
  // boundary crossed by the last step (for FindNormal): shape, its global matrix and boolean selection
  int n_shape;
  Mat n_mat;
  int n_sel;

  void reset_top() {
    level = 0;
    vol[0] = S->d->top_volume;
    node[0] = -1;
    glob[0] = kIdentity;
  }
  void cd_down(int node_idx) {
    const rbg_node& nd = S->d->nodes[node_idx];
    glob[level + 1] = mul(glob[level], S->mat(nd.matrix));
    level++;
    node[level] = node_idx;
    vol[level] = nd.volume;
  }
  int shape_at(int lv) const { return S->d->volumes[vol[lv]].shape; }
  bool inside_level(int lv, const double* pt) const {
    double l[3];
    m2l(glob[lv], pt, l);
    return contains(*S, shape_at(lv), l);
  }
  bool many(int lv) const { return lv > 0 && S->d->nodes[node[lv]].overlap != 0; }
  // daughters of the current volume (not `skip`, index >= from) whose shape holds pt, in AddNode order
  int next_daughter_holding(const double* pt, int skip, int from) const {
    const rbg_volume& v = S->d->volumes[vol[level]];
    if (!S->vox.empty() && !S->vox[vol[level]].tree.empty()) {  // only daughters whose box holds the point (in the mother's frame)
      double lp[3];
      m2l(glob[level], pt, lp);
      std::vector<int> cand;
      vox_point(S->vox[vol[level]], lp, cand);
      for (int k : cand) {
        if (k < from) continue;
        int ni = v.first_node + k;
        if (ni == skip) continue;
        const rbg_node& nd = S->d->nodes[ni];
        Mat g = mul(glob[level], S->mat(nd.matrix));
        double l[3];
        m2l(g, pt, l);
        if (contains(*S, S->d->volumes[nd.volume].shape, l)) return k;
      }
      return -1;
    }
    for (int k = from; k < v.nnodes; k++) {
      int ni = v.first_node + k;
      if (ni == skip) continue;
      const rbg_node& nd = S->d->nodes[ni];
      Mat g = mul(glob[level], S->mat(nd.matrix));
      double l[3];
      m2l(g, pt, l);
      if (contains(*S, S->d->volumes[nd.volume].shape, l)) return k;
    }
    return -1;
  }
  // Downward half of TGeoNavigator::SearchNode.  Returns true when the branch entered an ordinary ("ONLY") node.
  // A daughter placed with AddNodeOverlap ("MANY") that holds the point forms a cluster with the later daughters holding it
  // (TGeoNavigator::GetTouchedCluster); TGeoNavigator::FindInCluster then takes the first member whose branch reaches an ONLY
  // node or that is the node FindNextBoundary announced (fNextNode, `prefer`), else the member whose branch ends deepest (the
  // first one on ties).
  bool descend(const double* pt, int skip, int prefer) {
    bool only = false;
    while (true) {
      const rbg_volume& v = S->d->volumes[vol[level]];
      int k = next_daughter_holding(pt, skip, 0);
      if (k < 0) return only;
      if (!S->d->nodes[v.first_node + k].overlap) {
        cd_down(v.first_node + k);
        only = true;
        skip = -1;
        continue;
      }
      const Nav top = *this;
      Nav best = *this;
      int best_level = -1;
      for (int m = k; m >= 0; m = top.next_daughter_holding(pt, skip, m + 1)) {
        Nav trial = top;
        trial.cd_down(v.first_node + m);
        bool o = !S->d->nodes[v.first_node + m].overlap;
        if (trial.descend(pt, -1, prefer)) o = true;
        if (o || v.first_node + m == prefer) { *this = trial; return true; }
        if (trial.level > best_level) { best = trial; best_level = trial.level; }
      }
      *this = best;
      return only;
    }
  }
  // TGeoNavigator::SearchNode(downwards, skipnode) — skip is a desc.nodes index or -1
  // returns false when the point is outside the top volume (level = -1)
  bool search_node(bool downwards, int skip, const double* pt, int prefer = -1) {
    if (!downwards) {
      while (true) {
        bool inside_current = (level > 0 && node[level] == skip) ? true : inside_level(level, pt);
        if (inside_current && many(level)) {  // GotoSafeLevel: up to the first ordinary node above the overlapping ones
          while (many(level)) level--;
          continue;
        }
        if (inside_current) break;
        skip = node[level];
        if (level == 0) { level = -1; return false; }
        level--;
      }
    }
    descend(pt, skip, prefer);
    return true;
  }
  // InitTrack -> FindNode
  void init_track(const double* p, const double* d) {
    memcpy(P, p, sizeof(P));
    memcpy(D, d, sizeof(D));
    reset_top();
    on_boundary = false;
    search_node(false, -1, P);
  }
  // physical (flattened, DFS pre-order) id of the current path; -1 outside
  int physical_id() const {
    if (level < 0) return -1;
    int id = 0;
    for (int lv = 1; lv <= level; lv++) {
      const rbg_volume& mv = S->d->volumes[vol[lv - 1]];
      id += 1;
      for (int k = mv.first_node; k < node[lv]; k++) id += S->subtree[S->d->nodes[k].volume];
    }
    return id;
  }
  // CrossBoundaryAndLocate: relocate at P + extra*D, then undo the push
  void cross_and_locate(bool downwards, int skip, int prefer = -1) {
    const double* tr = glob[level < 0 ? 0 : level].t;
    double trmax = 1. + fabs(tr[0]) + fabs(tr[1]) + fabs(tr[2]);
    double extra = 100. * (trmax + step) * kTol;
    double q[3] = {P[0] + extra * D[0], P[1] + extra * D[1], P[2] + extra * D[2]};
    search_node(downwards, skip, q, prefer);
  }
  // TGeoNavigator::FindNextBoundaryAndStep(Big).  Returns false if nothing is hit from outside.
  bool find_next_boundary_and_step(bool push_quirk) {
    double extra = (on_boundary && push_quirk) ? kTol : 0.0;
    on_boundary = false;
    for (int i = 0; i < 3; i++) P[i] += extra * D[i];
    step = kBig;
    n_sel = 0;
    if (level < 0) {
      int sel = 0;
      int topshape = S->d->volumes[S->d->top_volume].shape;
      double snext = dist_out(*S, topshape, P, D, kBig, &sel);
      if (snext > 1e29) {  // the top volume is not reachable: the ray stays outside
        n_shape = -1;
        return false;
      }
      if (snext <= 0) { snext = 0.0; step = snext; for (int i = 0; i < 3; i++) P[i] -= extra * D[i]; }
      else step = snext + extra;
      for (int i = 0; i < 3; i++) P[i] += snext * D[i];
      on_boundary = true;
      reset_top();
      n_shape = topshape; n_mat = kIdentity; n_sel = sel;
      cross_and_locate(true, -1);
      return true;
    }
    double lp[3], ld[3];
    m2l(glob[level], P, lp);
    m2lv(glob[level], D, ld);
    int sel_exit = 0;
    double snext = dist_in(*S, shape_at(level), lp, ld, &sel_exit);
    n_shape = shape_at(level); n_mat = glob[level]; n_sel = sel_exit;
    if (snext <= kTol) {
      snext = kTol;
      step = snext;
      on_boundary = true;
      for (int i = 0; i < 3; i++) P[i] += step * D[i];
      int skip = node[level];
      if (level == 0) { level = -1; return true; }
      level--;
      cross_and_locate(false, skip);
      return true;
    }
    bool exiting = false, entering = false;
    if (snext < step - kTol) { step = snext; exiting = true; }
    // FindNextDaughterBoundary: nearest daughter entry (first wins within tolerance)
    const rbg_volume& v = S->d->volumes[vol[level]];
    int idaughter = -1, dsel = 0;
    const bool voxels = !S->vox.empty() && !S->vox[vol[level]].tree.empty();
    std::vector<int> cand;
    if (voxels) vox_ray(S->vox[vol[level]], lp, ld, step > 1e29 ? kBig : step, cand);
    const int ncand = voxels ? (int)cand.size() : v.nnodes;
    for (int c = 0; c < ncand; c++) {
      const int k = voxels ? cand[c] : c;
      const rbg_node& nd = S->d->nodes[v.first_node + k];
      Mat lm = S->mat(nd.matrix);
      double dp[3], dd[3];
      m2l(lm, lp, dp);
      m2lv(lm, ld, dd);
      int sel = 0;
      double s = dist_out(*S, S->d->volumes[nd.volume].shape, dp, dd, step, &sel);
      if (s < step - kTol) { step = s; idaughter = v.first_node + k; dsel = sel; entering = true; exiting = false; }
    }
    // overlapping nodes on the branch (TGeoNavigator::FindNextBoundary, fNmany > 0): the mother of each such node and the node's
    // sisters bound the step as well
    int xkind = 0, xlevel = -1, xnode = -1, xsel = 0;
    for (int lv = level; lv >= 1; lv--) {
      if (!many(lv)) continue;
      double mp[3], md[3];
      m2l(glob[lv - 1], P, mp);
      m2lv(glob[lv - 1], D, md);
      int sel = 0;
      double s = dist_in(*S, shape_at(lv - 1), mp, md, &sel);
      if (s < step - kTol) { step = s; xkind = 1; xlevel = lv - 1; xsel = sel; entering = exiting = false; }
      const rbg_volume& mv = S->d->volumes[vol[lv - 1]];
      for (int k = 0; k < mv.nnodes; k++) {
        int ni = mv.first_node + k;
        if (ni == node[lv]) continue;
        const rbg_node& nd = S->d->nodes[ni];
        Mat lm = S->mat(nd.matrix);
        double dp[3], dd[3];
        m2l(lm, mp, dp);
        m2lv(lm, md, dd);
        int shp = S->d->volumes[nd.volume].shape;
        sel = 0;
        if (nd.overlap && contains(*S, shp, dp)) s = dist_in(*S, shp, dp, dd, &sel);
        else s = dist_out(*S, shp, dp, dd, step, &sel);
        if (s < step - kTol) { step = s; xkind = 2; xlevel = lv - 1; xnode = ni; xsel = sel; entering = exiting = false; }
      }
    }
    for (int i = 0; i < 3; i++) P[i] += step * D[i];
    step += extra;
    on_boundary = true;
    if (xkind == 1) {  // left the mother of an overlapping node
      n_shape = shape_at(xlevel); n_mat = glob[xlevel]; n_sel = xsel;
      int skip = node[xlevel];
      if (xlevel == 0) { level = -1; return true; }
      level = xlevel - 1;
      cross_and_locate(false, skip);
      return true;
    }
    if (xkind == 2) {  // met a sister of an overlapping node: relocate from their mother
      const rbg_node& nd = S->d->nodes[xnode];
      n_shape = S->d->volumes[nd.volume].shape; n_mat = mul(glob[xlevel], S->mat(nd.matrix)); n_sel = xsel;
      level = xlevel;
      cross_and_locate(false, -1, xnode);
      return true;
    }
    if (entering) {
      if (S->d->nodes[idaughter].overlap) {  // an ordinary sister holding the point has priority over an overlapping daughter
        const rbg_node& nd = S->d->nodes[idaughter];
        n_shape = S->d->volumes[nd.volume].shape; n_mat = mul(glob[level], S->mat(nd.matrix)); n_sel = dsel;
        cross_and_locate(false, -1, idaughter);
        return true;
      }
      cd_down(idaughter);
      n_shape = shape_at(level); n_mat = glob[level]; n_sel = dsel;
      cross_and_locate(true, -1);
      return true;
    }
    (void)exiting;
    int skip = node[level];
    if (level == 0) { level = -1; return true; }
    level--;
    cross_and_locate(false, skip);
    return true;
  }
  // FindNormal (FindNormalFast): normal of the crossed shape in master frame, normal.dir >= 0
  void find_normal(double* n) const {
    if (n_shape < 0) { n[0] = n[1] = 0; n[2] = 1; return; }
    double lp[3], ld[3], ln[3];
    m2l(n_mat, P, lp);
    m2lv(n_mat, D, ld);
    normal(*S, n_shape, lp, ld, n_sel, ln);
    l2mv(n_mat, ln, n);
  }
  // Step(is_geom=true, cross=true) after SetStep(s): move by s + 1e-6 along D, FindNode()
  void step_and_locate(double s) {
    for (int i = 0; i < 3; i++) P[i] += (s + 1e-6) * D[i];
    on_boundary = false;
    if (level < 0) reset_top();
    search_node(false, -1, P);
  }
};

// ================================================================== the tracer state machine
struct RayState {
  double x[4];   // last point
  double d[3];
  double lambda;
  int status, npoints, last_node;
};

struct Tracer {
  const Scene& S;
  const rbg_trace_opts& o;
  Nav nav;
  Rng rng;
  int limit;

  Tracer(const Scene& s, const rbg_trace_opts& opts) : S(s), o(opts) {
    nav.S = &S;
    limit = o.limit > 0 ? o.limit : 100;
  }
  int vol_type(int vol) const { return vol < 0 ? RBG_NULL : S.d->volumes[vol].type; }
  // AOpticalComponent::FindBorderSurfaceCondition   src/AOpticalComponent.cxx:51-65
  const rbg_border* find_border(int vol1, int vol2) const {
    if (vol1 < 0) return nullptr;
    const rbg_volume& v = S.d->volumes[vol1];
    for (int i = 0; i < v.nborders; i++)
      if (S.d->borders[v.first_border + i].vol2 == vol2) return &S.d->borders[v.first_border + i];
    return nullptr;
  }
  // src/AOpticsManager.cxx:250-301
  void get_facet_normal(int cur_vol, int next_vol, double* normal) {
    nav.find_normal(normal);
    const double* mom = nav.D;
    const rbg_border* c = find_border(cur_vol, next_vol);
    if (c && c->lambertian) return;
    if (c && c->sigma != 0) {
      double sigma_alpha = c->sigma, f_max = std::min(1., 4. * sigma_alpha), fn[3];
      do {
        double alpha;
        do {
          alpha = rng.gaus(0, sigma_alpha);
        } while (f_max * rng.uniform() > sin(alpha) || alpha >= kPi / 2);
        double phi = 2 * kPi * rng.uniform();
        double sa = sin(alpha), ca = cos(alpha), sp = sin(phi), cp = cos(phi);
        double px = sa * cp, py = sa * sp, pz = ca;
        // TVector3::RotateUz(normal)
        double u1 = normal[0], u2 = normal[1], u3 = normal[2], up = u1 * u1 + u2 * u2;
        if (up) {
          up = sqrt(up);
          fn[0] = (u1 * u3 * px - u2 * py + u1 * up * pz) / up;
          fn[1] = (u2 * u3 * px + u1 * py + u2 * up * pz) / up;
          fn[2] = (u3 * u3 * px - px + u3 * up * pz) / up;
        } else if (u3 < 0.) { fn[0] = -px; fn[1] = py; fn[2] = -pz; }
        else { fn[0] = px; fn[1] = py; fn[2] = pz; }
      } while (dot3(mom, fn) <= 0.0);
      normal[0] = fn[0]; normal[1] = fn[1]; normal[2] = fn[2];
    }
  }
  double mirror_reflectance(int vol, double lambda, double angle) const {  // src/AMirror.cxx:39-60
    const rbg_volume& v = S.d->volumes[vol];
    double ret = 1.0;
    if (v.mirror >= 0) {
      const rbg_mirror& m = S.d->mirrors[v.mirror];
      if (m.graph2d >= 0) ret = graph2d_interp(S.d, m.graph2d, lambda, angle);
      else if (m.th2 >= 0) ret = th2_interp(S.d, m.th2, lambda, angle);
      else if (m.graph1d >= 0) ret = graph_eval(S.d, m.graph1d, lambda);
      else ret = m.constant;
    }
    ret = ret > 1 ? 1 : ret;
    ret = ret < 0 ? 0 : ret;
    return ret;
  }
  // ARay keeps every point (TGeoTrack::AddPoint) and every node (ARay::AddNode, include/ARay.h:35) of the current ray
  std::vector<double> track;  // x,y,z,t per point; the caller seeds it with the start point
  std::vector<int> nodes;     // one entry per AddNode call
  void add_point(RayState& r, const double* p, double t) {
    r.x[0] = p[0]; r.x[1] = p[1]; r.x[2] = p[2]; r.x[3] = t;
    r.npoints++;
    track.insert(track.end(), {p[0], p[1], p[2], t});
  }
  void add_node(RayState& r, int node) {
    r.last_node = node;
    nodes.push_back(node);
  }
  // src/AOpticsManager.cxx:170-247
  void do_reflection(double n1, RayState& r, int cur_vol, int next_vol, int next_phys, const double* normal_in) {
    double step = nav.step;
    double n[3];
    if (normal_in) memcpy(n, normal_in, sizeof(n));
    else get_facet_normal(cur_vol, next_vol, n);
    double d1[3] = {r.d[0], r.d[1], r.d[2]};
    double cos1 = dot3(d1, n);
    const rbg_border* cond = find_border(cur_vol, next_vol);
    bool absorbed = false;
    if (vol_type(next_vol) == RBG_MIRROR) {
      double angle = ACosT(cos1), ref;
      if (cond && cond->multilayer >= 0) {
        double tr;
        coherent_tmm_mixed(S.d, cond->multilayer, angle, r.lambda, ref, tr);
      } else ref = mirror_reflectance(next_vol, r.lambda, angle);
      if (ref < rng.uniform()) { absorbed = true; r.status = RBG_ABSORB; }
    }
    double d2[3];
    if (cond && cond->lambertian) {
      double y = 0.5 * rng.uniform();
      double theta = ASinT(sqrt(2 * y));
      double phi = 2 * kPi * rng.uniform();
      // TVector3::Theta()/Phi() of n, then TGeoRotation("", phi_n+90, theta_n+180, 0).LocalToMaster
      double perp = sqrt(n[0] * n[0] + n[1] * n[1]);
      double theta_n = (n[0] == 0 && n[1] == 0 && n[2] == 0 ? 0 : atan2(perp, n[2])) * 180. / kPi;
      double phi_n = (n[0] == 0 && n[1] == 0 ? 0 : atan2(n[1], n[0])) * 180. / kPi;
      double ph = (phi_n + 90) * kPi / 180., th = (theta_n + 180) * kPi / 180.;
      double sinphi = sin(ph), cosphi = cos(ph), sinthe = sin(th), costhe = cos(th);
      double R[9] = {cosphi, -costhe * sinphi, sinthe * sinphi, sinphi, costhe * cosphi, -sinthe * cosphi, 0, sinthe, costhe};
      double v[3] = {sin(theta) * cos(phi), sin(theta) * sin(phi), cos(theta)};
      for (int i = 0; i < 3; i++) d2[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
    } else {
      for (int i = 0; i < 3; i++) d2[i] = d1[i] - 2 * n[i] * cos1;
    }
    if (!absorbed) {
      double mag = sqrt(dot3(d2, d2));
      if (mag > 0) { r.d[0] = d2[0] / mag; r.d[1] = d2[1] / mag; r.d[2] = d2[2] / mag; }
    }
    double speed = kC / n1;
    double t = r.x[3] + step / speed;
    if (o.quirks & RBG_QUIRK_STEPBACK) {
      nav.D[0] = -d1[0]; nav.D[1] = -d1[1]; nav.D[2] = -d1[2];
      nav.step_and_locate(kEpsilon);
    } else {
      // idealised variant: vertex exactly on the surface, relocate on the incoming side
      nav.D[0] = -d1[0]; nav.D[1] = -d1[1]; nav.D[2] = -d1[2];
      double keep[3] = {nav.P[0], nav.P[1], nav.P[2]};
      nav.step_and_locate(kEpsilon);
      memcpy(nav.P, keep, sizeof(keep));
    }
    nav.D[0] = r.d[0]; nav.D[1] = r.d[1]; nav.D[2] = r.d[2];
    if (absorbed) { nav.D[0] = d2[0]; nav.D[1] = d2[1]; nav.D[2] = d2[2]; }
    add_point(r, nav.P, t);
    add_node(r, next_phys);
  }
  // src/AOpticsManager.cxx:52-167
  void do_fresnel(double n1, double n2, double k2, RayState& r, int cur_vol, int next_vol, int next_phys) {
    double step = nav.step;
    double n[3];
    get_facet_normal(cur_vol, next_vol, n);
    double d1[3] = {r.d[0], r.d[1], r.d[2]};
    double cos1 = dot3(d1, n);
    double sin1 = sqrt(1 - cos1 * cos1);
    double sin2 = n1 * sin1 / n2;
    double cos2 = sqrt(1 - sin2 * sin2);
    const rbg_border* cond = find_border(cur_vol, next_vol);
    bool absorbed = false, skip_fresnel = false;
    if (cond && cond->multilayer >= 0) {
      double reflectance, transmittance;
      coherent_tmm_mixed(S.d, cond->multilayer, ACosT(cos1), r.lambda, reflectance, transmittance);
      double rnd = rng.uniform();
      if (rnd < reflectance) { do_reflection(n1, r, cur_vol, next_vol, next_phys, n); return; }
      else if (rnd < reflectance + transmittance) skip_fresnel = true;
      else { absorbed = true; skip_fresnel = true; }
    }
    if (!skip_fresnel) {
      if (sin2 > 1.) { do_reflection(n1, r, cur_vol, next_vol, next_phys, n); return; }
      if (!o.disable_fresnel) {
        double Rs, Rp;
        if (k2 <= 0.) {
          double eta1S = n1 * cos1, eta2S = n2 * cos2, eta1P = n1 / cos1, eta2P = n2 / cos2;
          Rs = sq((eta1S - eta2S) / (eta1S + eta2S));
          Rp = sq((eta1P - eta2P) / (eta1P + eta2P));
        } else {
          double eta1S = n1 * cos1, eta1P = n1 / cos1, x1S = eta1S, x1P = eta1P;
          double u = sq(n2) - sq(k2) - sq(n1 * sin1), v = 2 * n2 * k2;
          double tmp = sqrt(sq(u) + sq(v));
          double cosxi2 = sqrt(1 + u / tmp) / sqrt(2.), sinxi2 = sqrt(1 - u / tmp) / sqrt(2.);
          double x2S = sqrt(tmp) * cosxi2, y2S = sqrt(tmp) * sinxi2;
          tmp = sq(x2S) + sq(y2S);
          double x2P = (2 * n2 * k2 * y2S + (sq(n2) - sq(k2)) * x2S) / tmp;
          double y2P = (2 * n2 * k2 * x2S - (sq(n2) - sq(k2)) * y2S) / tmp;
          Rs = (sq(x1S - x2S) + sq(y2S)) / (sq(x1S + x2S) + sq(y2S));
          Rp = (sq(x1P - x2P) + sq(y2P)) / (sq(x1P + x2P) + sq(y2P));
        }
        double R = (Rs + Rp) / 2.;
        if (rng.uniform() < R) { do_reflection(n1, r, cur_vol, next_vol, next_phys, n); return; }
      }
    }
    double d2[3];
    for (int i = 0; i < 3; i++) d2[i] = sin1 != 0 ? (d1[i] - cos1 * n[i]) * sin2 / sin1 + n[i] * cos2 : d1[i];
    double speed = kC / n1;
    double t = r.x[3] + step / speed;
    add_point(r, nav.P, t);
    add_node(r, next_phys);
    if (absorbed) r.status = RBG_ABSORB;
    else {
      double mag = sqrt(dot3(d2, d2));
      if (mag > 0) { r.d[0] = d2[0] / mag; r.d[1] = d2[1] / mag; r.d[2] = d2[2] / mag; }
      nav.D[0] = r.d[0]; nav.D[1] = r.d[1]; nav.D[2] = r.d[2];
    }
  }
  double lens_n(int vol, double lam) const { return index_n(S.d, S.d->volumes[vol].index, lam); }
  double lens_k(int vol, double lam) const { return index_k(S.d, S.d->volumes[vol].index, lam); }

  // src/AOpticsManager.cxx:347-519, one ray
  void trace(RayState& r, uint64_t ray_id) {
    rng.key[0] = (uint32_t)o.seed; rng.key[1] = (uint32_t)(o.seed >> 32);
    rng.id[0] = (uint32_t)ray_id; rng.id[1] = (uint32_t)(ray_id >> 32);
    rng.ndraw = 0;
    double lambda = r.lambda;
    nav.init_track(r.x, r.d);
    while (r.status == RBG_RUN) {
      double x1[4] = {r.x[0], r.x[1], r.x[2], r.x[3]}, d1[3] = {r.d[0], r.d[1], r.d[2]};
      int cur_vol = nav.level < 0 ? -1 : nav.vol[nav.level];
      nav.find_next_boundary_and_step((o.quirks & RBG_QUIRK_BOUNDARY_PUSH) != 0);
      double step = nav.step;
      int next_vol = nav.level < 0 ? -1 : nav.vol[nav.level];
      int next_phys = nav.physical_id();
      int typeCurrent = vol_type(cur_vol), typeNext = vol_type(next_vol);
      if (typeCurrent == RBG_LENS) {  // :403-423
        double abs = index_abslen(S.d, S.d->volumes[cur_vol].index, lambda);
        if (abs > 0 && abs != kInf) {
          double abs_step = -abs * log(rng.uniform());
          if (abs_step < step) {
            double n1 = lens_n(cur_vol, lambda), speed = kC / n1;
            double x2[3] = {x1[0] + abs_step * d1[0], x1[1] + abs_step * d1[1], x1[2] + abs_step * d1[2]};
            add_point(r, x2, x1[3] + abs_step / speed);
            add_node(r, next_phys);
            r.status = RBG_ABSORB;
            continue;
          }
        }
      }
      bool curVac = typeCurrent == RBG_NULL || typeCurrent == RBG_OPT || typeCurrent == RBG_OTHER;
      if ((curVac || typeCurrent == RBG_LENS) && typeNext == RBG_MIRROR) {  // :425-432
        double n1 = typeCurrent == RBG_LENS ? lens_n(cur_vol, lambda) : 1.;
        do_reflection(n1, r, cur_vol, next_vol, next_phys, nullptr);
      } else if (curVac && typeNext == RBG_LENS) {  // :433-441
        do_fresnel(1, lens_n(next_vol, lambda), lens_k(next_vol, lambda), r, cur_vol, next_vol, next_phys);
      } else if ((curVac || typeCurrent == RBG_LENS) && (typeNext == RBG_OBS || typeNext == RBG_FOCUS)) {  // :442-457
        double speed = typeCurrent == RBG_LENS ? kC / lens_n(cur_vol, lambda) : kC;
        add_point(r, nav.P, x1[3] + step / speed);
        add_node(r, next_phys);
      } else if (curVac && (typeNext == RBG_OTHER || typeNext == RBG_OPT)) {  // :458-466
        add_point(r, nav.P, x1[3] + step / kC);
        add_node(r, next_phys);
      } else if (typeCurrent == RBG_LENS && typeNext == RBG_LENS) {  // :467-474
        do_fresnel(lens_n(cur_vol, lambda), lens_n(next_vol, lambda), lens_k(next_vol, lambda), r, cur_vol, next_vol, next_phys);
      } else if (typeCurrent == RBG_LENS && (typeNext == RBG_NULL || typeNext == RBG_OPT || typeNext == RBG_OTHER)) {  // :475-482
        do_fresnel(lens_n(cur_vol, lambda), 1, 0, r, cur_vol, next_vol, next_phys);
      }
      if (typeNext == RBG_NULL) {  // :485-491
        add_point(r, nav.P, x1[3] + step / kC);
        add_node(r, next_phys);
        r.status = RBG_EXIT;
      } else if (typeCurrent == RBG_FOCUS || typeCurrent == RBG_OBS || typeCurrent == RBG_MIRROR || typeNext == RBG_OBS) {
        r.status = RBG_STOP;
      } else if (typeNext == RBG_FOCUS) {  // :495-513
        const rbg_volume& fv = S.d->volumes[next_vol];
        double angle = 0., qe = 1.;
        bool has_angle = fv.focal >= 0 && S.d->focals[fv.focal].qe_angle >= 0;
        if (has_angle) {
          double n[3];
          get_facet_normal(cur_vol, next_vol, n);
          angle = ACosT(dot3(r.d, n));
        }
        if (fv.focal >= 0) {
          const rbg_focal& f = S.d->focals[fv.focal];
          if (f.qe_lambda >= 0) qe = graph_eval(S.d, f.qe_lambda, lambda);
          if (has_angle) qe *= graph_eval(S.d, f.qe_angle, angle);
        }
        if (qe == 1 || rng.uniform() < qe) r.status = RBG_FOCUSED;
        else r.status = RBG_STOP;
      }
      if (r.status == RBG_RUN && r.npoints >= limit) r.status = RBG_SUSPEND;  // :515-517
    }
  }
};

void check_desc(const rbg_scene_desc* d) {
  if (!d || d->abi_version != RBG_ABI_VERSION) throw std::runtime_error("bad scene desc");
}

}  // namespace

// ==================================================================== C interface (ctypes)
extern "C" {

// TraceNonSequential over a host SoA batch; nthreads contiguous chunks like src/AOpticsManager.cxx:529-568
int orc_trace_history(const rbg_scene_desc* desc, const rbg_trace_opts* opts, const rbg_rays* rays, const rbg_history* hist, int nthreads);
int orc_trace(const rbg_scene_desc* desc, const rbg_trace_opts* opts, const rbg_rays* rays, int nthreads) {
  return orc_trace_history(desc, opts, rays, nullptr, nthreads);
}
// same, also returning each ray's polyline and node history in the layout of rbg_history (point k of ray i at k*n + i)
int orc_trace_history(const rbg_scene_desc* desc, const rbg_trace_opts* opts, const rbg_rays* rays, const rbg_history* hist, int nthreads) {
  try {
    check_desc(desc);
    if (rays->on_device) return RBG_EINVAL;
    Scene S(desc);
    if (g_use_voxels) build_voxels(S);
    int64_t n = rays->n;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n > 0 ? (int)n : 1;
    std::vector<std::string> errs(nthreads);
    auto work = [&](int tid) {
      try {
        int64_t chunk = n / nthreads, b = chunk * tid, e = tid == nthreads - 1 ? n : chunk * (tid + 1);
        Tracer T(S, *opts);
        for (int64_t i = b; i < e; i++) {
          RayState r;
          r.x[0] = rays->x[i]; r.x[1] = rays->y[i]; r.x[2] = rays->z[i]; r.x[3] = rays->t[i];
          r.d[0] = rays->dx[i]; r.d[1] = rays->dy[i]; r.d[2] = rays->dz[i];
          {  // ARay's constructor normalises the direction (src/ARay.cxx:28-36, :210-223)
            double mag = sqrt(dot3(r.d, r.d));
            if (mag > 0) { r.d[0] /= mag; r.d[1] /= mag; r.d[2] /= mag; }
          }
          r.lambda = rays->lambda[i];
          r.status = RBG_RUN; r.npoints = 1; r.last_node = -1;
          T.track.assign(r.x, r.x + 4);
          T.nodes.clear();
          T.trace(r, opts->ray_id_offset + (uint64_t)i);
          if (hist && hist->max_points > 0) {
            if ((int)T.track.size() != 4 * r.npoints || (int)T.nodes.size() != r.npoints - 1) throw std::runtime_error("AddPoint/AddNode out of step");
            for (int k = 0; k < r.npoints && k < hist->max_points; k++) {
              int64_t o = (int64_t)k * n + i;
              hist->hx[o] = T.track[4 * k]; hist->hy[o] = T.track[4 * k + 1]; hist->hz[o] = T.track[4 * k + 2]; hist->ht[o] = T.track[4 * k + 3];
              hist->hnode[o] = k == 0 ? -1 : T.nodes[k - 1];
            }
          }
          rays->ox[i] = r.x[0]; rays->oy[i] = r.x[1]; rays->oz[i] = r.x[2]; rays->ot[i] = r.x[3];
          rays->odx[i] = r.d[0]; rays->ody[i] = r.d[1]; rays->odz[i] = r.d[2];
          rays->status[i] = r.status; rays->last_node[i] = r.last_node; rays->npoints[i] = r.npoints;
        }
      } catch (std::exception& ex) { errs[tid] = ex.what(); }
    };
    if (nthreads == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
      for (auto& t : th) t.join();
    }
    for (auto& e : errs)
      if (!e.empty()) { fprintf(stderr, "orc_trace: %s\n", e.c_str()); return RBG_EINTERNAL; }
    return RBG_OK;
  } catch (std::exception& ex) {
    fprintf(stderr, "orc_trace: %s\n", ex.what());
    return RBG_EINTERNAL;
  }
}

// daughter-box trees on (default) / off (walk all daughters of a volume like a TGeoVolume without voxels); returns the old setting
int orc_set_voxels(int on) {
  int old = g_use_voxels;
  g_use_voxels = on;
  return old;
}

// AMultilayer::CoherentTMM (mode 0; complex angle, optional reversed stack) and IncoherentTMM (mode 1); pol 0 = S, 1 = P
int orc_tmm_general(const rbg_scene_desc* desc, int ml, int mode, int pol, int reverse, double th_re, double th_im, double lambda, double* R, double* T) {
  try {
    if (mode == 0) coherent_tmm_cplx(desc, ml, pol, cplx(th_re, th_im), lambda, *R, *T, reverse != 0);
    else incoherent_tmm(desc, ml, pol, cplx(th_re, th_im), lambda, *R, *T);
    return RBG_OK;
  } catch (...) { return RBG_EINTERNAL; }
}
// pol: 0 = S, 1 = P, 2 = mixed (uses precalculated tables when present)
int orc_tmm(const rbg_scene_desc* desc, int ml, int pol, double theta, double lambda, double* R, double* T) {
  try {
    if (pol == 2) coherent_tmm_mixed(desc, ml, theta, lambda, *R, *T);
    else coherent_tmm(desc, ml, pol, theta, lambda, *R, *T);
    return RBG_OK;
  } catch (...) { return RBG_EINTERNAL; }
}
double orc_index_n(const rbg_scene_desc* desc, int id, double lambda) { return index_n(desc, id, lambda); }
double orc_index_k(const rbg_scene_desc* desc, int id, double lambda) { return index_k(desc, id, lambda); }
double orc_index_abslen(const rbg_scene_desc* desc, int id, double lambda) { return index_abslen(desc, id, lambda); }
double orc_graph_eval(const rbg_scene_desc* desc, int g, double x) { return graph_eval(desc, g, x); }
double orc_th2_interp(const rbg_scene_desc* desc, int h, double x, double y) { return th2_interp(desc, h, x, y); }
double orc_graph2d_interp(const rbg_scene_desc* desc, int g, double x, double y) { return graph2d_interp(desc, g, x, y); }

// AGeoUtil::ContainmentRadius, src/AGeoUtil.cxx:18-42 (SumInRadius) and :198-308, on a flat histogram: bins[i + nx*j]
// = content of bin (i+1, j+1), stats = {sum w, sum x, sum y, sum x^2, sum y^2} of the in-range fills (what TH2::GetMean /
// GetStdDev read), out = {r, x, y}.
static double sum_in_radius(const double* bins, int nx, double xmin, double wx, int ny, double ymin, double wy, double x, double y, double r) {
  double r2 = r * r, total = 0.;
  for (int ix = 1; ix <= nx; ++ix) {
    double cx = xmin + (ix - 1) * wx + 0.5 * wx;  // TAxis::GetBinCenter
    for (int iy = 1; iy <= ny; ++iy) {
      double cy = ymin + (iy - 1) * wy + 0.5 * wy;
      double c = bins[(ix - 1) + (size_t)nx * (iy - 1)];
      if (c <= 0) continue;
      double d2 = (cx - x) * (cx - x) + (cy - y) * (cy - y);
      if (d2 <= r2) total += c;
    }
  }
  return total;
}
int orc_containment_radius(const double* bins, int nx, double xmin, double xmax, int ny, double ymin, double ymax, const double* stats, double fraction,
                           double* out) {
  double wx = (xmax - xmin) / nx, wy = (ymax - ymin) / ny;
  auto S = [&](double x, double y, double r) { return sum_in_radius(bins, nx, xmin, wx, ny, ymin, wy, x, y, r); };
  double sw = stats[0];
  double x = sw != 0 ? stats[1] / sw : 0, y = sw != 0 ? stats[2] / sw : 0;
  double sdx = sw != 0 ? sqrt(fabs(stats[3] / sw - x * x)) : 0, sdy = sw != 0 ? sqrt(fabs(stats[4] / sw - y * y)) : 0;
  double r = sqrt(sdx * sdx + sdy * sdy) * 1.5;
  double dr = 0.1 * r;
  double integral = 0;
  for (size_t b = 0; b < (size_t)nx * ny; b++) integral += bins[b];
  double sum_goal = integral * fraction;
  int no_shift = 0, no_stable = 0;
  for (int i = 0; i < 100 && no_shift < 30; i++) {
    bool stable_r = false, stable_x = true, stable_y = true;
    double sum0 = S(x, y, r);
    double next_r = r;
    if (sum0 < sum_goal) {
      double sum1 = S(x, y, r + dr);
      if (sum1 == sum0) { dr *= 2.; continue; }
      next_r = r + dr * (sum_goal - sum0) / (sum1 - sum0);
    } else if (sum0 != sum_goal) {
      double sum1 = S(x, y, r - dr);
      if (sum1 == sum0) { dr *= 2.; continue; }
      next_r = r - dr * (sum0 - sum_goal) / (sum0 - sum1);
    }
    if (next_r < 0.) next_r = 0.5 * r;
    if (next_r < 0.5 * r) next_r = 0.5 * r;
    if (next_r > 2. * r) next_r = 2. * r;
    stable_r = fabs(next_r - r) < 0.0001 * r;
    r = next_r;
    {
      double sum1 = S(x, y, r);
      dr *= sum0 != sum_goal ? fabs((sum1 - sum_goal) / (sum0 - sum_goal)) : 0.5;
      if (dr > 0.5 * r) dr = 0.5 * r;
      if (dr < 0.0005 * r) dr = 0.0005 * r;
      no_shift++;
      for (double dx = 0.25 * r; dx > 0.1 * dr; dx *= 0.25) {
        double sum_x1 = S(x + dx, y, r), sum_x2 = S(x - dx, y, r);
        while (sum_x1 > sum1) { no_shift = 0; x += dx; sum_x2 = sum1; sum1 = sum_x1; sum_x1 = S(x + dx, y, r); stable_x = false; }
        while (sum_x2 > sum1) { no_shift = 0; x -= dx; sum_x1 = sum1; sum1 = sum_x2; sum_x2 = S(x - dx, y, r); stable_x = false; }
      }
    }
    for (double dy = 0.1 * r; dy > 0.1 * dr; dy *= 0.25) {
      double sum1 = S(x, y, r), sum_y1 = S(x, y + dy, r), sum_y2 = S(x, y - dy, r);
      while (sum_y1 > sum1) { no_shift = 0; y += dy; sum_y2 = sum1; sum1 = sum_y1; sum_y1 = S(x, y + dy, r); stable_y = false; }
      while (sum_y2 > sum1) { no_shift = 0; y -= dy; sum_y1 = sum1; sum1 = sum_y2; sum_y2 = S(x, y - dy, r); stable_y = false; }
    }
    if (stable_r && stable_x && stable_y) no_stable++;
    else no_stable = 0;
    if (no_stable >= 4) break;
  }
  out[0] = r; out[1] = x; out[2] = y;
  return RBG_OK;
}

// shape-level entry points (local frame) for shape parity tests
int orc_shape_contains(const rbg_scene_desc* desc, int shape, const double* p) {
  Scene S(desc);
  return contains(S, shape, p) ? 1 : 0;
}
double orc_shape_dist(const rbg_scene_desc* desc, int shape, const double* p, const double* d, int from_inside) {
  Scene S(desc);
  int sel;
  return from_inside ? dist_in(S, shape, p, d, &sel) : dist_out(S, shape, p, d, kBig, &sel);
}
int orc_shape_normal(const rbg_scene_desc* desc, int shape, const double* p, const double* d, double* n) {
  Scene S(desc);
  normal(S, shape, p, d, 0, n);
  return 0;
}

// ACorsikaIACTFile::GetRayArray restated (src/ACorsikaIACTFile.cxx:71-133): bunch i yields rays while j < photons[i]; rays
// [first, first+n) of the concatenation are written (tracer units: cm, s).  The random wavelength of an undetermined bunch uses
// the Philox stream of the global ray index instead of the shared gRandom.
int orc_shoot_bunches(const rbg_bunches* b, int64_t first, int64_t n, double* x, double* y, double* z, double* t, double* dx, double* dy, double* dz,
                      double* lambda) {
  const double cm = 1., ns = 1e-9, nm = 1e-7, m = 100.;
  int64_t ray = 0, out = 0;
  for (int64_t i = 0; i < b->nbunches && out < n; i++) {
    double airmass = -1. / b->cz[i];
    double tel_dist = (b->z - b->telescope_z * cm) * airmass;
    double speed = 2.99792458e8 * m / b->refractive_index;
    double px = b->x[i] * cm - tel_dist * b->cx[i], py = b->y[i] * cm - tel_dist * b->cy[i], pt = b->time[i] * ns - tel_dist / speed;
    for (int j = 0; j < b->photons[i] && out < n; j++, ray++) {
      if (ray < first) continue;
      double lam = b->lambda[i];
      if (lam == 0) {
        Rng r;
        r.key[0] = (uint32_t)b->seed; r.key[1] = (uint32_t)(b->seed >> 32);
        r.id[0] = (uint32_t)ray; r.id[1] = (uint32_t)((uint64_t)ray >> 32);
        r.ndraw = 0x40000000u;  // shooter draws live in their own counter range (as orc_shoot)
        lam = 1. / (1. / b->lambda_min_nm - r.uniform() * (1. / b->lambda_min_nm - 1. / b->lambda_max_nm));
      }
      x[out] = px; y[out] = py; z[out] = b->z; t[out] = pt;
      dx[out] = b->cx[i]; dy[out] = b->cy[i]; dz[out] = b->cz[i];
      lambda[out] = lam * nm;
      out++;
    }
  }
  return out == n ? 0 : -1;
}

// Philox uniform stream check: k-th uniform of ray `id`
double orc_uniform(uint64_t seed, uint64_t id, uint32_t k) {
  Rng r;
  r.key[0] = (uint32_t)seed; r.key[1] = (uint32_t)(seed >> 32);
  r.id[0] = (uint32_t)id; r.id[1] = (uint32_t)(id >> 32);
  r.ndraw = k;
  return r.uniform();
}

// ARayShooter restated with the Philox stream of the device shooters (src/ARayShooter.cxx:122-460)
int orc_shoot(const rbg_shoot_desc* s, int64_t first, int64_t n, double* x, double* y, double* z, double* t, double* dx, double* dy, double* dz,
              double* lambda) {
  Mat rot = kIdentity, tr = kIdentity;
  memcpy(rot.r, s->rot, sizeof(rot.r));
  memcpy(tr.t, s->tr, sizeof(tr.t));
  double nd[3];
  l2mv(rot, s->dir, nd);
  double mag = sqrt(dot3(nd, nd));
  if (mag > 0) { nd[0] /= mag; nd[1] /= mag; nd[2] /= mag; }
  // Circle: ring start offsets
  for (int64_t j = 0; j < n; j++) {
    uint64_t id = (uint64_t)(first + j);
    Rng r;
    r.key[0] = (uint32_t)s->seed; r.key[1] = (uint32_t)(s->seed >> 32);
    r.id[0] = (uint32_t)id; r.id[1] = (uint32_t)(id >> 32);
    r.ndraw = 0x40000000u;  // shooter draws live in their own counter range
    double p[3] = {0, 0, 0};
    if (s->kind >= 4 && s->kind <= 6) {  // src/ARayShooter.cxx:240-392: point sources at tr
      double v[3], w[3];
      if (s->kind == 4) {  // RandomCone(lambda, r = dx, d = dy, n, rot, tr)
        double rr = s->dx;
        do {
          p[0] = -rr + 2 * rr * r.uniform();
          p[1] = -rr + 2 * rr * r.uniform();
        } while (p[0] * p[0] + p[1] * p[1] > rr * rr);
        p[2] = s->dy;
        l2mv(rot, p, v);
      } else if (s->kind == 5) {  // RandomSphere: TRandom::Sphere(x, y, z, 1)
        double a = 0, b = 0, r2 = 1;
        while (r2 > 0.25) {
          a = r.uniform() - 0.5;
          b = r.uniform() - 0.5;
          r2 = a * a + b * b;
        }
        double scale = 8.0 * sqrt(0.25 - r2);
        v[0] = a * scale; v[1] = b * scale; v[2] = -1. + 8.0 * r2;
      } else {  // RandomSphericalCone(lambda, n, theta = dx deg, rot, tr)
        double c0 = cos(s->dx * kPi / 180.), ran = c0 + (1. - c0) * r.uniform(), th = acos(std::min(1., std::max(-1., ran))), phi = 2 * kPi * r.uniform();
        double l[3] = {sin(th) * cos(phi), sin(th) * sin(phi), cos(th)};
        l2mv(rot, l, v);
      }
      double vm = sqrt(dot3(v, v));
      if (vm > 0) { v[0] /= vm; v[1] /= vm; v[2] /= vm; }
      double zero[3] = {0, 0, 0};
      l2m(tr, zero, w);
      x[j] = w[0]; y[j] = w[1]; z[j] = w[2]; t[j] = 0;
      dx[j] = v[0]; dy[j] = v[1]; dz[j] = v[2];
      lambda[j] = s->lambda_min == s->lambda_max ? s->lambda_min : s->lambda_min + (s->lambda_max - s->lambda_min) * r.uniform();
      continue;
    }
    if (s->kind == 0) {
      int64_t i = (int64_t)id / s->ny, k = (int64_t)id % s->ny;
      double deltax = s->nx == 1 ? s->dx / 2 : s->dx / (s->nx - 1), deltay = s->ny == 1 ? s->dy / 2 : s->dy / (s->ny - 1);
      p[0] = i * deltax - s->dx / 2;
      p[1] = k * deltay - s->dy / 2;
    } else if (s->kind == 1) {
      p[0] = -s->dx / 2 + s->dx * r.uniform();
      p[1] = -s->dy / 2 + s->dy * r.uniform();
    } else if (s->kind == 2) {
      double rmax = s->dx;
      do {
        p[0] = -rmax + 2 * rmax * r.uniform();
        p[1] = -rmax + 2 * rmax * r.uniform();
      } while (sqrt(p[0] * p[0] + p[1] * p[1]) > rmax);
    } else if (s->kind == 3) {
      // ray 0 is the centre; ring i (1..nr) has nphi*i rays
      int64_t idx = (int64_t)id;
      if (idx > 0) {
        int64_t i = 0, acc = 1;
        while (idx >= acc + (int64_t)s->ny * (i + 1)) { acc += (int64_t)s->ny * (i + 1); i++; }
        int64_t k = idx - acc;
        double rr = s->dx * (i + 1) / s->nx, phi = 2 * kPi / s->ny / (i + 1) * k;
        p[0] = rr * cos(phi);
        p[1] = rr * sin(phi);
      }
    } else return RBG_EINVAL;
    double q[3], w[3];
    l2mv(rot, p, q);
    l2m(tr, q, w);
    x[j] = w[0]; y[j] = w[1]; z[j] = w[2]; t[j] = 0;
    dx[j] = nd[0]; dy[j] = nd[1]; dz[j] = nd[2];
    lambda[j] = s->lambda_min == s->lambda_max ? s->lambda_min : s->lambda_min + (s->lambda_max - s->lambda_min) * r.uniform();
  }
  return RBG_OK;
}

}  // extern "C"
