// rb_build.h — host-side scene preparation shared by the CUDA library and the host emulation used
// for debugging: validates the flat description, derives per-shape constants, flattens placed nodes
// into physical paths (DFS pre-order, id 0 = top) and builds one threaded BVH per mother node.
#ifndef RB_BUILD_H
#define RB_BUILD_H
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "rb_scene.h"

namespace {  // internal linkage: this header is compiled into more than one shared object

struct NotSupported : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct Invalid : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct Box {
  double lo[3], hi[3];
};
static Box box_empty() { return Box{{1e300, 1e300, 1e300}, {-1e300, -1e300, -1e300}}; }
static void box_add(Box& b, const double* p) {
  for (int i = 0; i < 3; i++) {
    b.lo[i] = std::min(b.lo[i], p[i]);
    b.hi[i] = std::max(b.hi[i], p[i]);
  }
}
static Box box_transform(const Box& b, const DMat& m) {  // local -> master
  Box o = box_empty();
  for (int c = 0; c < 8; c++) {
    double l[3] = {c & 1 ? b.hi[0] : b.lo[0], c & 2 ? b.hi[1] : b.lo[1], c & 4 ? b.hi[2] : b.lo[2]}, q[3];
    for (int i = 0; i < 3; i++) q[i] = m.t[i] + m.r[3 * i] * l[0] + m.r[3 * i + 1] * l[1] + m.r[3 * i + 2] * l[2];
    box_add(o, q);
  }
  return o;
}
static DMat mat_identity() {
  DMat m;
  memset(&m, 0, sizeof(m));
  m.r[0] = m.r[4] = m.r[8] = 1;
  return m;
}
static DMat mat_mul(const DMat& a, const DMat& b) {
  DMat o;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) o.r[3 * i + j] = a.r[3 * i] * b.r[j] + a.r[3 * i + 1] * b.r[3 + j] + a.r[3 * i + 2] * b.r[6 + j];
    o.t[i] = a.t[i] + a.r[3 * i] * b.t[0] + a.r[3 * i + 1] * b.t[1] + a.r[3 * i + 2] * b.t[2];
  }
  return o;
}

struct SceneBuilder {
  const rbg_scene_desc* D;
  std::vector<DShape> shapes;
  std::vector<double> dpar;
  std::vector<DMat> mats;
  std::vector<Box> shape_box;
  std::vector<int> shape_depth;
  std::vector<DNode> nodes;
  std::vector<DBvh> bvh;
  std::vector<DBox> boxes;
  std::vector<std::string> names;

  DMat mat(int id) const {
    if (id < 0) return mat_identity();
    DMat m;
    memcpy(m.r, D->matrices[id].rot, sizeof(m.r));
    memcpy(m.t, D->matrices[id].tr, sizeof(m.t));
    return m;
  }
  void build_shapes() {
    const double deg = M_PI / 180.;
    shapes.resize(D->nshapes);
    shape_box.resize(D->nshapes);
    shape_depth.assign(D->nshapes, 0);
    for (int i = 0; i < D->nmatrices; i++) mats.push_back(mat(i));
    for (int i = 0; i < D->nshapes; i++) {  // operands always precede their composite (export order)
      const rbg_shape& s = D->shapes[i];
      const double* P = D->dpar + s.ipar;
      DShape& o = shapes[i];
      o.type = s.type; o.left = s.left; o.right = s.right;
      auto tag = [&](int m) {  // RB_MAT_TRANS: the device skips the rotation of a pure translation
        if (m < 0) return m;
        const double* r = D->matrices[m].rot;
        const bool unit = r[0] == 1 && r[4] == 1 && r[8] == 1 && r[1] == 0 && r[2] == 0 && r[3] == 0 && r[5] == 0 && r[6] == 0 && r[7] == 0;
        return unit ? (m | RB_MAT_TRANS) : m;
      };
      o.lmat = tag(s.lmat); o.rmat = tag(s.rmat);
      o.ipar = (int)dpar.size();
      Box b = box_empty();
      auto setbox = [&](double x, double y, double zlo, double zhi) { b = Box{{-x, -y, zlo}, {x, y, zhi}}; };
      switch (s.type) {
        case RBG_SHAPE_BBOX:
          dpar.insert(dpar.end(), P, P + 6);
          b = Box{{P[3] - P[0], P[4] - P[1], P[5] - P[2]}, {P[3] + P[0], P[4] + P[1], P[5] + P[2]}};
          break;
        case RBG_SHAPE_TUBE:
          dpar.insert(dpar.end(), P, P + 3);
          setbox(P[1], P[1], -P[2], P[2]);
          break;
        case RBG_SHAPE_SPHERE: {
          dpar.insert(dpar.end(), P, P + 6);
          int flags = (P[2] > 0 ? 1 : 0) | (P[3] < 180 ? 2 : 0) | (fabs(P[5] - P[4] - 360.) > 1e-9 ? 4 : 0);
          double v[9] = {cos(P[2] * deg), sin(P[2] * deg), cos(P[3] * deg), sin(P[3] * deg), cos(P[4] * deg), sin(P[4] * deg), cos(P[5] * deg),
                         sin(P[5] * deg), (double)flags};
          dpar.insert(dpar.end(), v, v + 9);
          dpar.push_back(0);
          double zc[4] = {P[0] * v[0], P[0] * v[2], P[1] * v[0], P[1] * v[2]};
          double zlo = *std::min_element(zc, zc + 4), zhi = *std::max_element(zc, zc + 4);
          double rho = (P[2] <= 90 && 90 <= P[3]) ? P[1] : P[1] * std::max(v[1], v[3]);
          setbox(rho, rho, zlo, zhi);
          break;
        }
        case RBG_SHAPE_PARABOLOID: {
          dpar.insert(dpar.end(), P, P + 3);
          double dd = 1. / (P[1] * P[1] - P[0] * P[0]);
          dpar.push_back(2. * P[2] * dd);
          dpar.push_back(-P[2] * (P[0] * P[0] + P[1] * P[1]) * dd);
          double rm = std::max(P[0], P[1]);
          setbox(rm, rm, -P[2], P[2]);
          break;
        }
        case RBG_SHAPE_PGON:
        case RBG_SHAPE_PCON: {
          // device layout (both): phi1,dphi,nedges (0 = polycone),nz, nz x (z,rmin,rmax), nedges x (cos,sin) of the edge-centre
          // azimuths, then: general flag (hollow or azimuthal segment), cos/sin(phi1), cos/sin(phi1 + dphi)
          const bool pgon = s.type == RBG_SHAPE_PGON;
          int ne = pgon ? (int)P[2] : 0, nz = (int)P[pgon ? 3 : 2];
          const double* sec = P + (pgon ? 4 : 3);
          if ((pgon && ne < 3) || nz < 2) throw Invalid("TGeoPgon/TGeoPcon needs nedges >= 3 and nz >= 2");
          if (!(P[1] > 0) || P[1] > 360. + 1e-9) throw Invalid("TGeoPgon/TGeoPcon needs 0 < dphi <= 360");
          const bool seg = fabs(P[1] - 360.) > 1e-9;
          bool hollow = false;
          double rmx = 0;
          for (int k = 0; k < nz; k++) {
            if (sec[3 * k + 1] > 0) hollow = true;
            rmx = std::max(rmx, sec[3 * k + 2]);
            if (k > 0 && sec[3 * k] < sec[3 * (k - 1)]) throw Invalid("TGeoPgon/TGeoPcon sections must be ordered in z");
          }
          dpar.push_back(P[0]); dpar.push_back(P[1]); dpar.push_back((double)ne); dpar.push_back((double)nz);
          dpar.insert(dpar.end(), sec, sec + 3 * nz);
          for (int e = 0; e < ne; e++) {
            double ph = (P[0] + (e + 0.5) * P[1] / ne) * deg;
            dpar.push_back(cos(ph));
            dpar.push_back(sin(ph));
          }
          dpar.push_back(hollow || seg ? 1. : 0.);
          dpar.push_back(cos(P[0] * deg)); dpar.push_back(sin(P[0] * deg));
          dpar.push_back(cos((P[0] + P[1]) * deg)); dpar.push_back(sin((P[0] + P[1]) * deg));
          double Rb = pgon ? rmx / cos(0.5 * P[1] / ne * deg) : rmx;  // circumscribed radius of the polygon's corners
          setbox(Rb, Rb, sec[0], sec[3 * (nz - 1)]);
          break;
        }
        case RBG_SHAPE_ASPHERE: {
          int n1 = (int)P[8], n2 = (int)P[9];
          dpar.insert(dpar.end(), P, P + 12 + n1 + n2);
          setbox(P[7], P[7], P[10] - P[11], P[10] + P[11]);
          break;
        }
        case RBG_SHAPE_WINSTON2D:
        case RBG_SHAPE_WINSTONPOLY: {
          double r1 = P[0], r2 = P[1], theta = asin(r2 / r1), dz = (r1 + r2) / tan(theta) / 2., f = r2 * (1 + sin(theta));
          int npoly = s.type == RBG_SHAPE_WINSTONPOLY ? std::max(3, (int)P[2]) : 0;
          double v[10] = {r1, r2, s.type == RBG_SHAPE_WINSTONPOLY ? (double)npoly : P[2], theta, dz, f, cos(theta), sin(theta), tan(theta),
                          npoly ? tan(M_PI / npoly) : 0.};
          dpar.insert(dpar.end(), v, v + 10);
          for (int k = 0; k < npoly; k++) {  // face azimuths k 2pi/N
            dpar.push_back(cos(k * 2 * M_PI / npoly));
            dpar.push_back(sin(k * 2 * M_PI / npoly));
          }
          if (s.type == RBG_SHAPE_WINSTON2D) setbox(r1, P[2], -dz, dz);
          else {
            double R = r1 / cos(M_PI / std::max(3, (int)P[2]));
            setbox(R, R, -dz, dz);
          }
          break;
        }
        case RBG_SHAPE_ARB8: {
          // dz, 8 x (x,y).  TGeoArb8::ComputeTwist: both faces must turn the same way; counter-clockwise input is re-ordered
          // (vertices 1 <-> 3 and 5 <-> 7) so that the device sees ROOT's clockwise convention.
          if (s.npar < 17 || !(P[0] > 0)) throw Invalid("TGeoArb8 needs dz > 0 and 8 vertices");
          double v[16];
          memcpy(v, P + 1, sizeof(v));
          double sum1 = 0, sum2 = 0;
          for (int k = 0; k < 4; k++) {
            int j = (k + 1) % 4;
            sum1 += v[2 * k] * v[2 * j + 1] - v[2 * j] * v[2 * k + 1];
            sum2 += v[2 * k + 8] * v[2 * j + 9] - v[2 * j + 8] * v[2 * k + 9];
          }
          if (sum1 * sum2 < -1e-10) throw Invalid("TGeoArb8: lower and upper faces turn in opposite directions");
          if (sum1 > 1e-10 || sum2 > 1e-10) {
            std::swap(v[2], v[6]); std::swap(v[3], v[7]);
            std::swap(v[10], v[14]); std::swap(v[11], v[15]);
          }
          dpar.push_back(P[0]);
          dpar.insert(dpar.end(), v, v + 16);
          double xlo = v[0], xhi = v[0], ylo = v[1], yhi = v[1];
          for (int k = 1; k < 8; k++) {
            xlo = std::min(xlo, v[2 * k]); xhi = std::max(xhi, v[2 * k]);
            ylo = std::min(ylo, v[2 * k + 1]); yhi = std::max(yhi, v[2 * k + 1]);
          }
          b = Box{{xlo, ylo, -P[0]}, {xhi, yhi, P[0]}};
          break;
        }
        case RBG_SHAPE_XTRU: {
          // nvert, nz, nvert x (x,y), nz x (z,x0,y0,scale)
          int nv = s.npar >= 2 ? (int)P[0] : 0, nz = s.npar >= 2 ? (int)P[1] : 0;
          if (nv < 3 || nz < 2 || s.npar < 2 + 2 * nv + 4 * nz) throw Invalid("TGeoXtru needs >= 3 vertices and >= 2 sections");
          const double* V = P + 2;
          const double* sec = P + 2 + 2 * nv;
          double lo[3] = {1e300, 1e300, sec[0]}, hi[3] = {-1e300, -1e300, sec[4 * (nz - 1)]};
          for (int k = 0; k < nz; k++) {
            if (k > 0 && sec[4 * k] < sec[4 * (k - 1)]) throw Invalid("TGeoXtru sections must be ordered in z");
            if (!(sec[4 * k + 3] > 0)) throw Invalid("TGeoXtru section scale must be positive");
            for (int q = 0; q < nv; q++) {
              double x = sec[4 * k + 1] + sec[4 * k + 3] * V[2 * q], y = sec[4 * k + 2] + sec[4 * k + 3] * V[2 * q + 1];
              lo[0] = std::min(lo[0], x); hi[0] = std::max(hi[0], x);
              lo[1] = std::min(lo[1], y); hi[1] = std::max(hi[1], y);
            }
          }
          dpar.insert(dpar.end(), P, P + 2 + 2 * nv + 4 * nz);
          b = Box{{lo[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]}};
          break;
        }
        case RBG_SHAPE_UNION:
        case RBG_SHAPE_INTERSECTION:
        case RBG_SHAPE_SUBTRACTION: {
          if (s.left < 0 || s.left >= i || s.right < 0 || s.right >= i) throw Invalid("composite operand out of order");
          Box bl = box_transform(shape_box[s.left], mat(s.lmat)), br = box_transform(shape_box[s.right], mat(s.rmat));
          if (s.type == RBG_SHAPE_UNION)
            for (int k = 0; k < 3; k++) { b.lo[k] = std::min(bl.lo[k], br.lo[k]); b.hi[k] = std::max(bl.hi[k], br.hi[k]); }
          else if (s.type == RBG_SHAPE_INTERSECTION) {
            for (int k = 0; k < 3; k++) { b.lo[k] = std::max(bl.lo[k], br.lo[k]); b.hi[k] = std::min(bl.hi[k], br.hi[k]); }
            tighten_sphere_cut(s, b);
          } else b = bl;
          shape_depth[i] = 1 + std::max(shape_depth[s.left], shape_depth[s.right]);
          break;
        }
        default: throw Invalid("unknown shape type");
      }
      shape_box[i] = b;
    }
  }
  // Mirror facets are a thin spherical shell cut by a prism or a tube about the same axis direction ("mirSphere:transZ*mirCut",
  // tutorials/DaviesCotton.C:70, MST.C:135): the box of the cutter is as tall as the cutter (20 cm for a Davies-Cotton facet) while
  // the solid itself is only as tall as the sag of the cap over the cutter's footprint (0.75 cm).  The thin box keeps the facet
  // out of the candidate lists of the rays that pass over it or leave a neighbouring facet (profiles/r2_summary.md).
  // Solid = shell(rmin..rmax about c) with polar range, cut by a solid inside the cylinder of radius rc about the axis through a,
  // both operands placed by translations only.  z-range of the shell points whose distance from that axis is <= rc.
  void tighten_sphere_cut(const rbg_shape& s, Box& b) const {
    const rbg_shape *sp = &D->shapes[s.left], *cu = &D->shapes[s.right];
    int smat = s.lmat, cmat = s.rmat;
    if (sp->type != RBG_SHAPE_SPHERE) { std::swap(sp, cu); std::swap(smat, cmat); }
    if (sp->type != RBG_SHAPE_SPHERE) return;
    if (cu->type != RBG_SHAPE_PGON && cu->type != RBG_SHAPE_PCON && cu->type != RBG_SHAPE_TUBE) return;
    auto pure_translation = [&](int m) {
      if (m < 0) return true;
      const double* r = D->matrices[m].rot;
      return r[0] == 1 && r[4] == 1 && r[8] == 1 && r[1] == 0 && r[2] == 0 && r[3] == 0 && r[5] == 0 && r[6] == 0 && r[7] == 0;
    };
    if (!pure_translation(smat) || !pure_translation(cmat)) return;
    const double* P = D->dpar + sp->ipar;  // rmin, rmax, theta1, theta2, phi1, phi2
    const double* C = D->dpar + cu->ipar;
    double c[3] = {0, 0, 0}, a[3] = {0, 0, 0};
    if (smat >= 0) memcpy(c, D->matrices[smat].tr, sizeof(c));
    if (cmat >= 0) memcpy(a, D->matrices[cmat].tr, sizeof(a));
    double rc = 0;
    if (cu->type == RBG_SHAPE_TUBE) rc = C[1];
    else {
      const bool pg = cu->type == RBG_SHAPE_PGON;
      int nz = (int)C[pg ? 3 : 2];
      const double* sec = C + (pg ? 4 : 3);
      for (int k = 0; k < nz; k++) rc = std::max(rc, sec[3 * k + 2]);
      if (pg) rc /= cos(0.5 * C[1] / C[2] * M_PI / 180.);
    }
    const double rmin = P[0], rmax = P[1], h = std::hypot(c[0] - a[0], c[1] - a[1]);
    const double rho_lo = std::max(0., h - rc), rho_hi = h + rc;  // distance of the footprint from the sphere's axis
    if (rho_lo >= rmax) return;
    // heights (relative to the sphere centre) of the shell points above the footprint: +-sqrt(r^2 - rho^2)
    const double top_max = sqrt(std::max(0., rmax * rmax - rho_lo * rho_lo));                          // highest point of the upper half
    const double top_min = rho_hi < rmin ? sqrt(rmin * rmin - rho_hi * rho_hi) : 0.;                   // lowest point of the upper half
    const double deg = M_PI / 180.;
    const bool upper = P[2] < 90., lower = P[3] > 90.;  // polar range reaches above / below the equator
    double zlo = 1e300, zhi = -1e300;
    if (upper) { zlo = std::min(zlo, c[2] + (lower ? -top_max : top_min)); zhi = std::max(zhi, c[2] + top_max); }
    if (lower) { zlo = std::min(zlo, c[2] - top_max); zhi = std::max(zhi, c[2] + (upper ? top_max : -top_min)); }
    (void)deg;
    if (zlo > zhi) return;
    const double pad = 1e-6 * (1. + rmax);
    b.lo[2] = std::max(b.lo[2], zlo - pad);
    b.hi[2] = std::min(b.hi[2], zhi + pad);
  }
  // which flat evaluator can take the shape (RB_LEAF_*, rb_scene.h)
  int leaf_kind(int sh) const {
    auto prim = [&](int i) { return shapes[i].type < RBG_SHAPE_UNION || shapes[i].type > RBG_SHAPE_SUBTRACTION; };
    auto general_poly = [&](int i) {  // hollow or azimuthally cut TGeoPgon/TGeoPcon: flag behind the edge table (see build_shapes)
      if (shapes[i].type != RBG_SHAPE_PGON && shapes[i].type != RBG_SHAPE_PCON) return false;
      const double* P = dpar.data() + shapes[i].ipar;
      return P[4 + 3 * (int)P[3] + 2 * (int)P[2]] != 0.;
    };
    if (prim(sh)) return RB_LEAF_PRIM;
    const DShape& s = shapes[sh];
    if (prim(s.left) && prim(s.right)) {
      if (general_poly(s.left) || general_poly(s.right)) return RB_LEAF_GENERIC;
      return RB_LEAF_BOOL2 | (s.type << 4) | (shapes[s.left].type << 8) | (shapes[s.right].type << 12);
    }
    int i = sh, levels = 0;
    while (!prim(i)) {  // ((a + b) + c) + d
      if (shapes[i].type != RBG_SHAPE_UNION || !prim(shapes[i].right) || ++levels > 6) return RB_LEAF_GENERIC;
      i = shapes[i].left;
    }
    return RB_LEAF_UNIONS;
  }
  // DFS pre-order flattening of placed nodes into physical paths
  int flatten(int vol, const DMat& g, int mother, int overlap, const std::string& name) {
    if (vol < 0 || vol >= D->nvolumes) throw Invalid("bad volume id");
    int id = (int)nodes.size();
    if (id > 2000000) throw Invalid("too many physical nodes");
    DNode n;
    memset(&n, 0, sizeof(n));
    n.g = g;
    n.volume = vol;
    n.shape = D->volumes[vol].shape;
    n.type = D->volumes[vol].type;
    n.mother = mother;
    n.overlap = overlap;
    n.level = mother < 0 ? 0 : nodes[mother].level + 1;
    n.leaf = leaf_kind(n.shape);
    nodes.push_back(n);
    {
      Box wb = world_box(id);
      for (int k = 0; k < 3; k++) {  // same outward rounding + padding as the BVH leaves (bvh_rec)
        double pad = 2e-3 + 4e-7 * std::max(fabs(wb.lo[k]), fabs(wb.hi[k]));
        nodes[id].blo[k] = std::nextafterf((float)(wb.lo[k] - pad), -INFINITY);
        nodes[id].bhi[k] = std::nextafterf((float)(wb.hi[k] + pad), INFINITY);
      }
      const Box& lb = shape_box[n.shape];
      for (int k = 0; k < 3; k++) {
        const double c = 0.5 * (lb.lo[k] + lb.hi[k]), h = 0.5 * (lb.hi[k] - lb.lo[k]);
        nodes[id].lc[k] = (float)c;
        nodes[id].lh[k] = std::nextafterf((float)(h + fabs(c - (double)(float)c) + 2e-3 + 4e-7 * (fabs(c) + h)), INFINITY);
      }
    }
    names.push_back(name);
    const rbg_volume& v = D->volumes[vol];
    std::vector<int> kids;
    for (int k = 0; k < v.nnodes; k++) {
      const rbg_node& nd = D->nodes[v.first_node + k];
      std::string nm = std::string(D->names + D->volumes[nd.volume].name) + "_" + std::to_string(nd.copy_no);
      kids.push_back(flatten(nd.volume, mat_mul(g, mat(nd.matrix)), id, nd.overlap, nm));
    }
    build_bvh(id, kids);
    return id;
  }
  Box world_box(int node) const {
    Box b = box_transform(shape_box[nodes[node].shape], nodes[node].g);
    for (int k = 0; k < 3; k++) {
      double m = 1e-6 + 1e-9 * std::max(fabs(b.lo[k]), fabs(b.hi[k]));
      b.lo[k] -= m;
      b.hi[k] += m;
    }
    return b;
  }
  int bvh_rec(std::vector<std::pair<int, Box>>& items, int lo, int hi) {
    int me = (int)bvh.size();
    bvh.push_back(DBvh());
    Box b = box_empty();
    for (int i = lo; i < hi; i++) { box_add(b, items[i].second.lo); box_add(b, items[i].second.hi); }
    for (int k = 0; k < 3; k++) {  // outward rounding + padding that covers the fp32 error of the slab test
      double pad = 2e-3 + 4e-7 * std::max(fabs(b.lo[k]), fabs(b.hi[k]));
      bvh[me].lo[k] = std::nextafterf((float)(b.lo[k] - pad), -INFINITY);
      bvh[me].hi[k] = std::nextafterf((float)(b.hi[k] + pad), INFINITY);
    }
    if (hi - lo == 1) {
      bvh[me].child = items[lo].first;
    } else {
      bvh[me].child = -1;
      // surface-area heuristic, exact sweep over the three axes (N is at most a few thousand)
      auto area = [](const Box& x) {
        double dx = x.hi[0] - x.lo[0], dy = x.hi[1] - x.lo[1], dz = x.hi[2] - x.lo[2];
        return dx * dy + dy * dz + dz * dx;
      };
      int best_axis = -1, best_split = -1;
      double best_cost = 1e300;
      std::vector<double> right_area(hi - lo);
      for (int axis = 0; axis < 3; axis++) {
        std::sort(items.begin() + lo, items.begin() + hi, [axis](const std::pair<int, Box>& p, const std::pair<int, Box>& q) {
          return p.second.lo[axis] + p.second.hi[axis] < q.second.lo[axis] + q.second.hi[axis];
        });
        Box acc = box_empty();
        for (int i = hi - 1; i > lo; i--) {
          box_add(acc, items[i].second.lo);
          box_add(acc, items[i].second.hi);
          right_area[i - lo] = area(acc);
        }
        acc = box_empty();
        for (int i = lo; i < hi - 1; i++) {
          box_add(acc, items[i].second.lo);
          box_add(acc, items[i].second.hi);
          double cost = area(acc) * (i - lo + 1) + right_area[i + 1 - lo] * (hi - i - 1);
          if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = i + 1; }
        }
      }
      int axis = best_axis;
      std::sort(items.begin() + lo, items.begin() + hi, [axis](const std::pair<int, Box>& p, const std::pair<int, Box>& q) {
        return p.second.lo[axis] + p.second.hi[axis] < q.second.lo[axis] + q.second.hi[axis];
      });
      int mid = best_split;
      bvh_rec(items, lo, mid);
      bvh_rec(items, mid, hi);
    }
    bvh[me].skip = (int)bvh.size();  // fixed up to -1 at the end of this tree by the caller
    return me;
  }
  void build_bvh(int node, const std::vector<int>& kids) {
    nodes[node].bvh_first = (int)bvh.size();
    nodes[node].bvh_count = 0;
    if (kids.empty()) return;
    std::vector<std::pair<int, Box>> items;
    for (int k : kids) items.emplace_back(k, world_box(k));
    int first = (int)bvh.size();
    bvh_rec(items, 0, (int)items.size());
    int end = (int)bvh.size();
    for (int i = first; i < end; i++)
      if (bvh[i].skip >= end) bvh[i].skip = -1;
    nodes[node].bvh_count = end - first;
    nodes[node].box_first = (int)boxes.size();
    for (int i = first; i < end; i++)
      if (bvh[i].child >= 0) {
        DBox b;
        memset(&b, 0, sizeof(b));
        for (int k = 0; k < 3; k++) { b.lo[k] = bvh[i].lo[k]; b.hi[k] = bvh[i].hi[k]; }
        b.child = bvh[i].child;
        boxes.push_back(b);
      }
    nodes[node].box_count = (int)boxes.size() - nodes[node].box_first;
  }
};

// compile-time features a scene needs from the bounce kernel (see RB_PH_* / RB_SBIT in rb_device.cuh)
static void scene_features_needed(const rbg_scene_desc* D, unsigned& shapes, unsigned& phys) {
  shapes = phys = 0;
  for (int i = 0; i < D->nshapes; i++) shapes |= 1u << D->shapes[i].type;
  for (int i = 0; i < D->nvolumes; i++) {
    if (D->volumes[i].type == RBG_LENS) phys |= 1u;
    if (D->volumes[i].focal >= 0) phys |= 16u;
  }
  for (int i = 0; i < D->nborders; i++) {
    if (D->borders[i].multilayer >= 0) phys |= 2u;
    if (D->borders[i].sigma != 0) phys |= 4u;
    if (D->borders[i].lambertian) phys |= 8u;
  }
  for (int i = 0; i < D->nmirrors; i++)
    if (D->mirrors[i].graph1d >= 0 || D->mirrors[i].th2 >= 0 || D->mirrors[i].graph2d >= 0) phys |= 32u;
  for (int i = 0; i < D->nnodes; i++)
    if (D->nodes[i].overlap) phys |= 64u;  // RB_PH_OVERLAP
}

static int scene_depth_needed(const SceneBuilder& B) {
  int d = 0;
  for (int v : B.shape_depth) d = std::max(d, v);
  return d;
}

static void validate_desc(const rbg_scene_desc* D) {
  if (!D) throw Invalid("null scene description");
  if (D->abi_version != RBG_ABI_VERSION) throw Invalid("scene description ABI version mismatch");
  if (D->top_volume >= D->nvolumes) throw Invalid("top volume out of range");
  for (int i = 0; i < D->nvolumes; i++) {
    const rbg_volume& v = D->volumes[i];
    if (v.shape < 0 || v.shape >= D->nshapes) throw Invalid("volume with bad shape id");
    if (v.first_node < 0 || v.first_node + v.nnodes > D->nnodes) throw Invalid("volume with bad node slice");
    if (v.first_border < 0 || v.first_border + v.nborders > D->nborders) throw Invalid("volume with bad border slice");
    if (v.index >= D->nindices || v.mirror >= D->nmirrors || v.focal >= D->nfocals) throw Invalid("volume with bad table id");
  }
  for (int i = 0; i < D->nmultilayers; i++) {
    const rbg_multilayer& m = D->multilayers[i];
    if (m.n < 2 || m.n > RB_MAX_TMM_LAYERS * 64 || m.first < 0 || m.first + m.n > D->nlayers) throw Invalid("bad multilayer");
  }
  for (int i = 0; i < D->nmirrors; i++)
    if (D->mirrors[i].graph2d >= D->ngraph2d) throw Invalid("mirror with bad TGraph2D id");
  for (int i = 0; i < D->ngraph2d; i++) {
    const rbg_graph2d& g = D->graph2d[i];
    if (g.first_tri < 0 || g.ntri < 0 || g.first_tri + g.ntri > D->ntri) throw Invalid("bad TGraph2D triangle slice");
  }
  for (int i = 0; i < 3 * D->ntri; i++)
    if (D->tri[i] < 0 || D->tri[i] >= D->ng2pts) throw Invalid("TGraph2D triangle vertex out of range");
}

}  // namespace
#endif
