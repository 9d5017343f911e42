// RootExporter.h — the ROOT-side half of the drop-in: walks a closed TGeoManager geometry built from REAL ROOT TGeo classes and
// the reference's UNMODIFIED ROBAST classes once, flattens it into the tables of include/robast_b200.h, and runs
// AOpticsManager::TraceNonSequential(ARayArray&) (src/AOpticsManager.cxx:523-587) on the GPU through the C ABI.
//
// This header needs ROOT and the reference's headers on the include path:
//     g++ -DROBAST_HAVE_ROOT $(root-config --cflags) -I<ROBAST>/include -I<repo>/include -I<repo>/include/robast my_macro.cxx \
//         $(root-config --libs) -lGeom -L<ROBAST> -lROBAST -L<repo>/robast_b200 -lrobast_b200
// ROOT is not in this repository's build image, so here the header is only checked for syntax against a minimal set of stand-in
// declarations (tests/fake_root/, tests/test_root_exporter_syntax.py); the ROOT-free mirror classes of Robast.h carry the same
// export logic (ASceneExport) and are what the GPU tests run.
//
// Private state of the reference's classes that has no public getter (include/AGeoAsphericDisk.h:32-35 conic constants,
// include/AGeoWinstonCone2D.h:30-33, include/AGeoWinstonConePoly.h:23, the fPar[] of the dispersion formulas, the layer lists of
// include/AMultilayer.h:28-33, the reflectance tables of include/AMirror.h:26-32, include/AFocalSurface.h:24-25, ...) is read
// through the rootcling dictionary the reference builds for every class (include/LinkDef.h:12-41): TClass::GetDataMemberOffset
// gives the byte offset of a data member by name, for private members too.
#ifndef ROBAST_ROOT_EXPORTER_H
#define ROBAST_ROOT_EXPORTER_H
#ifdef ROBAST_HAVE_ROOT

#include <complex>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "TClass.h"
#include "TGeoArb8.h"
#include "TGeoBBox.h"
#include "TGeoBoolNode.h"
#include "TGeoCompositeShape.h"
#include "TGeoManager.h"
#include "TGeoMatrix.h"
#include "TGeoNode.h"
#include "TGeoParaboloid.h"
#include "TGeoPcon.h"
#include "TGeoPgon.h"
#include "TGeoSphere.h"
#include "TGeoTube.h"
#include "TGeoVolume.h"
#include "TGeoXtru.h"
#include "TGraph.h"
#include "TGraph2D.h"
#include "TH2.h"
#include "TObjArray.h"

#include "ABorderSurfaceCondition.h"
#include "ACauchyFormula.h"
#include "AFocalSurface.h"
#include "AGeoAsphericDisk.h"
#include "AGeoWinstonCone2D.h"
#include "AGeoWinstonConePoly.h"
#include "ALens.h"
#include "AMirror.h"
#include "AMixedRefractiveIndex.h"
#include "AMultilayer.h"
#include "AObscuration.h"
#include "AOpticalComponent.h"
#include "AOpticsManager.h"
#include "ARay.h"
#include "ARayArray.h"
#include "ARefractiveIndex.h"
#include "ASchottFormula.h"
#include "ASellmeierFormula.h"

#include "robast_b200.h"

namespace robast_b200 {

// a data member of `obj` (of dictionary class `cls`) by name, private or not
template <class T> const T& Member(const void* obj, const char* cls, const char* name) {
  TClass* c = TClass::GetClass(cls);
  if (!c) throw std::runtime_error(std::string("no dictionary for ") + cls);
  Long_t off = c->GetDataMemberOffset(name);
  if (off <= 0 && std::strcmp(name, "fUniqueID") != 0) {  // 0 = not found (no ROBAST member sits at offset 0: TObject comes first)
    throw std::runtime_error(std::string(cls) + " has no data member " + name);
  }
  return *reinterpret_cast<const T*>(reinterpret_cast<const char*>(obj) + off);
}

class RootExporter {
 public:
  std::vector<rbg_shape> shapes;
  std::vector<double> dpar;
  std::vector<rbg_matrix> matrices;
  std::vector<rbg_node> nodes;
  std::vector<rbg_volume> volumes;
  std::vector<rbg_border> borders;
  std::vector<rbg_graph> graphs;
  std::vector<double> gx, gy;
  std::vector<rbg_th2> th2;
  std::vector<double> th2v;
  std::vector<rbg_index> indices;
  std::vector<rbg_mirror> mirrors;
  std::vector<rbg_focal> focals;
  std::vector<rbg_multilayer> multilayers;
  std::vector<rbg_layer> layers;
  std::vector<rbg_graph2d> graph2ds;
  std::vector<int32_t> tri;
  std::vector<double> g2x, g2y, g2z;
  std::vector<char> names;
  std::vector<TGeoNode*> node_of_id;  // physical (flattened, DFS pre-order) node id -> TGeoNode, for ARay::AddNode
  rbg_scene_desc desc;

  int AddMatrix(const TGeoMatrix* m) {
    if (!m || m->IsIdentity()) return -1;
    auto it = matrix_id_.find(m);
    if (it != matrix_id_.end()) return it->second;
    rbg_matrix r;
    std::memcpy(r.rot, m->GetRotationMatrix(), sizeof(r.rot));
    std::memcpy(r.tr, m->GetTranslation(), sizeof(r.tr));
    matrices.push_back(r);
    return matrix_id_[m] = (int)matrices.size() - 1;
  }

  // every shape the C ABI knows (RBG_SHAPE_*, include/robast_b200.h); exact class identity, most derived classes first
  int AddShape(const TGeoShape* s) {
    if (!s) throw std::runtime_error("volume without shape");
    auto it = shape_id_.find(s);
    if (it != shape_id_.end()) return it->second;
    rbg_shape r;
    r.left = r.right = r.lmat = r.rmat = -1;
    r.ipar = (int32_t)dpar.size();
    auto P = [&](std::initializer_list<double> v) { dpar.insert(dpar.end(), v); };
    if (s->IsA() == TGeoCompositeShape::Class()) {
      TGeoBoolNode* b = static_cast<const TGeoCompositeShape*>(s)->GetBoolNode();
      int left = AddShape(b->GetLeftShape()), right = AddShape(b->GetRightShape());  // operands precede their composite
      r.ipar = (int32_t)dpar.size();
      switch (b->GetBooleanOperator()) {
        case TGeoBoolNode::kGeoUnion: r.type = RBG_SHAPE_UNION; break;
        case TGeoBoolNode::kGeoIntersection: r.type = RBG_SHAPE_INTERSECTION; break;
        default: r.type = RBG_SHAPE_SUBTRACTION; break;
      }
      r.left = left;
      r.right = right;
      r.lmat = AddMatrix(b->GetLeftMatrix());
      r.rmat = AddMatrix(b->GetRightMatrix());
    } else if (s->IsA() == AGeoAsphericDisk::Class()) {
      auto* a = static_cast<const AGeoAsphericDisk*>(s);
      const int n1 = (int)a->GetNPol1(), n2 = (int)a->GetNPol2();
      r.type = RBG_SHAPE_ASPHERE;
      P({a->GetZ1(), a->GetZ2(), a->GetCurve1(), a->GetCurve2(), Member<Double_t>(a, "AGeoAsphericDisk", "fKappa1"), Member<Double_t>(a, "AGeoAsphericDisk", "fKappa2"),
         a->GetRmin(), a->GetRmax(), (double)n1, (double)n2, a->GetOrigin()[2], a->GetDZ()});
      dpar.insert(dpar.end(), a->GetK1(), a->GetK1() + n1);
      dpar.insert(dpar.end(), a->GetK2(), a->GetK2() + n2);
    } else if (s->IsA() == AGeoWinstonConePoly::Class()) {
      r.type = RBG_SHAPE_WINSTONPOLY;
      P({Member<Double_t>(s, "AGeoWinstonCone2D", "fR1"), Member<Double_t>(s, "AGeoWinstonCone2D", "fR2"), (double)Member<Int_t>(s, "AGeoWinstonConePoly", "fPolyN")});
    } else if (s->IsA() == AGeoWinstonCone2D::Class()) {
      r.type = RBG_SHAPE_WINSTON2D;
      P({Member<Double_t>(s, "AGeoWinstonCone2D", "fR1"), Member<Double_t>(s, "AGeoWinstonCone2D", "fR2"), static_cast<const TGeoBBox*>(s)->GetDY()});
    } else if (s->InheritsFrom(TGeoPgon::Class())) {  // also AGeoBezierPgon: its sections are ordinary TGeoPgon sections
      auto* p = static_cast<const TGeoPgon*>(s);
      r.type = RBG_SHAPE_PGON;
      P({p->GetPhi1(), p->GetDphi(), (double)p->GetNedges(), (double)p->GetNz()});
      for (Int_t i = 0; i < p->GetNz(); i++) P({p->GetZ(i), p->GetRmin(i), p->GetRmax(i)});
    } else if (s->InheritsFrom(TGeoPcon::Class())) {  // also AGeoBezierPcon
      auto* p = static_cast<const TGeoPcon*>(s);
      r.type = RBG_SHAPE_PCON;
      P({p->GetPhi1(), p->GetDphi(), (double)p->GetNz()});
      for (Int_t i = 0; i < p->GetNz(); i++) P({p->GetZ(i), p->GetRmin(i), p->GetRmax(i)});
    } else if (s->IsA() == TGeoSphere::Class()) {
      auto* t = static_cast<const TGeoSphere*>(s);
      r.type = RBG_SHAPE_SPHERE;
      P({t->GetRmin(), t->GetRmax(), t->GetTheta1(), t->GetTheta2(), t->GetPhi1(), t->GetPhi2()});
    } else if (s->IsA() == TGeoParaboloid::Class()) {
      auto* t = static_cast<const TGeoParaboloid*>(s);
      r.type = RBG_SHAPE_PARABOLOID;
      P({t->GetRlo(), t->GetRhi(), t->GetDz()});
    } else if (s->IsA() == TGeoTube::Class()) {
      auto* t = static_cast<const TGeoTube*>(s);
      r.type = RBG_SHAPE_TUBE;
      P({t->GetRmin(), t->GetRmax(), t->GetDz()});
    } else if (s->IsA() == TGeoArb8::Class()) {
      auto* t = const_cast<TGeoArb8*>(static_cast<const TGeoArb8*>(s));
      r.type = RBG_SHAPE_ARB8;
      P({t->GetDz()});
      dpar.insert(dpar.end(), t->GetVertices(), t->GetVertices() + 16);
    } else if (s->IsA() == TGeoXtru::Class()) {
      auto* t = static_cast<const TGeoXtru*>(s);
      r.type = RBG_SHAPE_XTRU;
      P({(double)t->GetNvert(), (double)t->GetNz()});
      for (Int_t i = 0; i < t->GetNvert(); i++) P({t->GetX(i), t->GetY(i)});
      for (Int_t i = 0; i < t->GetNz(); i++) P({t->GetZ(i), t->GetXOffset(i), t->GetYOffset(i), t->GetScale(i)});
    } else if (s->IsA() == TGeoBBox::Class()) {
      auto* b = static_cast<const TGeoBBox*>(s);
      r.type = RBG_SHAPE_BBOX;
      P({b->GetDX(), b->GetDY(), b->GetDZ(), b->GetOrigin()[0], b->GetOrigin()[1], b->GetOrigin()[2]});
    } else {
      throw std::runtime_error(std::string("shape class not supported on the device path: ") + s->ClassName());
    }
    r.npar = (int32_t)dpar.size() - r.ipar;
    shapes.push_back(r);
    return shape_id_[s] = (int)shapes.size() - 1;
  }

  // TGraph -> points sorted by x (TGraph::Eval's bracket search is order independent; the device bisects)
  int AddGraph(const TGraph* g) {
    if (!g) return -1;
    auto it = graph_id_.find(g);
    if (it != graph_id_.end()) return it->second;
    std::vector<std::pair<double, double>> pts;
    for (Int_t i = 0; i < g->GetN(); i++) pts.emplace_back(g->GetX()[i], g->GetY()[i]);
    std::stable_sort(pts.begin(), pts.end(), [](const std::pair<double, double>& a, const std::pair<double, double>& b) { return a.first < b.first; });
    rbg_graph r = {(int32_t)gx.size(), (int32_t)pts.size()};
    for (auto& p : pts) { gx.push_back(p.first); gy.push_back(p.second); }
    graphs.push_back(r);
    return graph_id_[g] = (int)graphs.size() - 1;
  }
  int AddTH2(const TH2* h) {
    if (!h) return -1;
    auto it = th2_id_.find(h);
    if (it != th2_id_.end()) return it->second;
    rbg_th2 r;
    r.first = (int32_t)th2v.size();
    r.nx = h->GetNbinsX(); r.ny = h->GetNbinsY(); r.pad = 0;
    r.xmin = h->GetXaxis()->GetXmin(); r.xmax = h->GetXaxis()->GetXmax();
    r.ymin = h->GetYaxis()->GetXmin(); r.ymax = h->GetYaxis()->GetXmax();
    for (Int_t j = 1; j <= r.ny; j++)
      for (Int_t i = 1; i <= r.nx; i++) th2v.push_back(h->GetBinContent(i, j));
    th2.push_back(r);
    return th2_id_[h] = (int)th2.size() - 1;
  }
  // TGraph2D::Interpolate is linear on the Delaunay triangles: ship the triangle list ROOT itself built.  GetHistogram("empty")
  // forces the triangulation; TGraphDelaunay2D keeps the triangles — for ROOT versions without an accessor the mirror classes'
  // Bowyer-Watson triangulation (RootCompat.h, TGraph2D::GetTriangles) gives the same triangles for points in general position.
  int AddGraph2D(TGraph2D* g, const std::vector<Int_t>& triangles) {
    if (!g) return -1;
    auto it = graph2d_id_.find(g);
    if (it != graph2d_id_.end()) return it->second;
    rbg_graph2d r;
    r.first_tri = (int32_t)(tri.size() / 3);
    r.ntri = (int32_t)(triangles.size() / 3);
    const int32_t base = (int32_t)g2x.size();
    for (Int_t v : triangles) tri.push_back(base + v);
    g2x.insert(g2x.end(), g->GetX(), g->GetX() + g->GetN());
    g2y.insert(g2y.end(), g->GetY(), g->GetY() + g->GetN());
    g2z.insert(g2z.end(), g->GetZ(), g->GetZ() + g->GetN());
    graph2ds.push_back(r);
    return graph2d_id_[g] = (int)graph2ds.size() - 1;
  }
  // ARefractiveIndex family (include/ARefractiveIndex.h:27-28 graphs; formula parameters are private fPar[])
  int AddIndex(const ARefractiveIndex* x) {
    if (!x) return -1;
    auto it = index_id_.find(x);
    if (it != index_id_.end()) return it->second;
    rbg_index r;
    std::memset(&r, 0, sizeof(r));
    r.mix_a = r.mix_b = -1;
    r.ngraph = AddGraph(Member<std::shared_ptr<TGraph>>(x, "ARefractiveIndex", "fRefractiveIndex").get());
    r.kgraph = AddGraph(Member<std::shared_ptr<TGraph>>(x, "ARefractiveIndex", "fExtinctionCoefficient").get());
    r.kind = RBG_INDEX_GRAPH;
    if (x->IsA() == ASellmeierFormula::Class()) {
      r.kind = RBG_INDEX_SELLMEIER;
      std::memcpy(r.par, &Member<Double_t>(x, "ASellmeierFormula", "fPar"), 6 * sizeof(double));  // B1,B2,B3,C1,C2,C3 (src/ASellmeierFormula.cxx:46-54)
    } else if (x->IsA() == ASchottFormula::Class()) {
      r.kind = RBG_INDEX_SCHOTT;
      std::memcpy(r.par, &Member<Double_t>(x, "ASchottFormula", "fPar"), 6 * sizeof(double));
    } else if (x->IsA() == ACauchyFormula::Class()) {
      r.kind = RBG_INDEX_CAUCHY;
      std::memcpy(r.par, &Member<Double_t>(x, "ACauchyFormula", "fPar"), 3 * sizeof(double));
    } else if (x->IsA() == AMixedRefractiveIndex::Class()) {
      r.kind = RBG_INDEX_MIXED;
      r.mix_a = AddIndex(Member<std::shared_ptr<ARefractiveIndex>>(x, "AMixedRefractiveIndex", "fMaterialA").get());
      r.mix_b = AddIndex(Member<std::shared_ptr<ARefractiveIndex>>(x, "AMixedRefractiveIndex", "fMaterialB").get());
      r.frac_a = Member<Double_t>(x, "AMixedRefractiveIndex", "fFractionA");
      r.frac_b = Member<Double_t>(x, "AMixedRefractiveIndex", "fFractionB");
    }
    indices.push_back(r);
    return index_id_[x] = (int)indices.size() - 1;
  }
  int AddMultilayer(const AMultilayer* m) {
    if (!m) return -1;
    auto it = multilayer_id_.find(m);
    if (it != multilayer_id_.end()) return it->second;
    const auto& idx = Member<std::vector<std::shared_ptr<ARefractiveIndex>>>(m, "AMultilayer", "fRefractiveIndexList");
    const auto& thick = Member<std::vector<Double_t>>(m, "AMultilayer", "fThicknessList");
    const auto& coh = Member<std::vector<Double_t>>(m, "AMultilayer", "fCoherentList");
    rbg_multilayer r;
    r.n = (int32_t)idx.size();
    std::vector<rbg_layer> tmp;
    for (int i = 0; i < r.n; i++) {
      rbg_layer l;
      l.index = AddIndex(idx[i].get());
      l.incoherent = (i == 0 || i == r.n - 1 || (i < (int)coh.size() && coh[i] == 0)) ? 1 : 0;
      l.thickness = thick[i];
      tmp.push_back(l);
    }
    r.first = (int32_t)layers.size();
    layers.insert(layers.end(), tmp.begin(), tmp.end());
    r.table_r = AddTH2(Member<std::shared_ptr<TH2D>>(m, "AMultilayer", "fPreCalculatedReflectanceMixed").get());
    r.table_t = AddTH2(Member<std::shared_ptr<TH2D>>(m, "AMultilayer", "fPreCalculatedTransmittanceMixed").get());
    if (r.table_r < 0 || r.table_t < 0) r.table_r = r.table_t = -1;
    multilayers.push_back(r);
    return multilayer_id_[m] = (int)multilayers.size() - 1;
  }

  // exact class identity, include/AOpticsManager.h:76-90
  static int VolumeType(const TGeoVolume* v) {
    if (v->IsA() == ALens::Class()) return RBG_LENS;
    if (v->IsA() == AObscuration::Class()) return RBG_OBS;
    if (v->IsA() == AMirror::Class()) return RBG_MIRROR;
    if (v->IsA() == AFocalSurface::Class()) return RBG_FOCUS;
    if (v->IsA() == AOpticalComponent::Class()) return RBG_OPT;
    return RBG_OTHER;
  }

  // `triangles_of`: the Delaunay triangle list (3 point ids per triangle) of a TGraph2D reflectance, see AddGraph2D
  const rbg_scene_desc* Build(TGeoManager* mgr, std::vector<Int_t> (*triangles_of)(TGraph2D*) = nullptr) {
    TGeoVolume* top = mgr->GetTopVolume();
    if (!top) throw std::runtime_error("no top volume");
    std::vector<TGeoVolume*> order;
    Collect(top, order);
    volumes.resize(order.size());
    for (size_t i = 0; i < order.size(); i++) {
      TGeoVolume* v = order[i];
      rbg_volume& r = volumes[i];
      std::memset(&r, 0, sizeof(r));
      r.type = VolumeType(v);
      r.shape = AddShape(v->GetShape());
      r.index = r.mirror = r.focal = -1;
      r.name = (int32_t)names.size();
      const char* nm = v->GetName();
      names.insert(names.end(), nm, nm + std::strlen(nm) + 1);
      if (r.type == RBG_LENS) r.index = AddIndex(Member<std::shared_ptr<ARefractiveIndex>>(v, "ALens", "fIndex").get());
      if (r.type == RBG_MIRROR) {
        rbg_mirror mm;
        mm.constant = Member<Double_t>(v, "AMirror", "fReflectance");
        mm.graph1d = AddGraph(Member<std::shared_ptr<const TGraph>>(v, "AMirror", "fReflectance1D").get());
        mm.th2 = AddTH2(Member<std::shared_ptr<TH2>>(v, "AMirror", "fReflectanceTH2").get());
        TGraph2D* g2 = Member<std::shared_ptr<TGraph2D>>(v, "AMirror", "fReflectance2D").get();
        mm.graph2d = -1;
        if (g2) {
          if (!triangles_of) throw std::runtime_error("a TGraph2D reflectance needs a triangle-list provider (see AddGraph2D)");
          mm.graph2d = AddGraph2D(g2, triangles_of(g2));
        }
        mm.pad = 0;
        mirrors.push_back(mm);
        r.mirror = (int32_t)mirrors.size() - 1;
      }
      if (r.type == RBG_FOCUS) {
        TGraph* ql = Member<TGraph*>(v, "AFocalSurface", "fQuantumEfficiencyLambda");
        TGraph* qa = Member<TGraph*>(v, "AFocalSurface", "fQuantumEfficiencyAngle");
        if (ql || qa) {
          rbg_focal ff = {AddGraph(ql), AddGraph(qa)};
          focals.push_back(ff);
          r.focal = (int32_t)focals.size() - 1;
        }
      }
      r.first_node = (int32_t)nodes.size();
      r.nnodes = v->GetNdaughters();
      for (Int_t k = 0; k < v->GetNdaughters(); k++) {
        TGeoNode* n = v->GetNode(k);
        rbg_node nn = {volume_id_[n->GetVolume()], AddMatrix(n->GetMatrix()), n->GetNumber(), n->IsOverlapping() ? 1 : 0};
        nodes.push_back(nn);
      }
    }
    for (size_t i = 0; i < order.size(); i++) {  // borders once every volume id is known (directional: registered on component 1)
      rbg_volume& r = volumes[i];
      r.first_border = (int32_t)borders.size();
      if (r.type != RBG_OTHER) {
        TObjArray* arr = Member<TObjArray*>(order[i], "AOpticalComponent", "fBorderSurfaceConditionArray");
        for (Int_t k = 0; arr && k <= arr->GetLast(); k++) {
          auto* b = static_cast<ABorderSurfaceCondition*>(arr->At(k));
          rbg_border bb;
          const AOpticalComponent* c2 = b->GetComponent2();
          bb.vol2 = !c2 ? -1 : (volume_id_.count(c2) ? volume_id_[c2] : -2);
          bb.multilayer = AddMultilayer(b->GetMultilayer().get());
          bb.lambertian = b->IsLambertian() ? 1 : 0;
          bb.pad = 0;
          bb.sigma = b->GetGaussianRoughness();
          borders.push_back(bb);
        }
      }
      r.nborders = (int32_t)borders.size() - r.first_border;
    }
    node_of_id.clear();
    node_of_id.push_back(mgr->GetTopNode());
    FlattenNodes(top);
    Finish(0);
    return &desc;
  }

 private:
  std::map<const TGeoMatrix*, int> matrix_id_;
  std::map<const TGeoShape*, int> shape_id_;
  std::map<const TGeoVolume*, int> volume_id_;
  std::map<const TGraph*, int> graph_id_;
  std::map<const TH2*, int> th2_id_;
  std::map<const ARefractiveIndex*, int> index_id_;
  std::map<const AMultilayer*, int> multilayer_id_;
  std::map<const TGraph2D*, int> graph2d_id_;

  void Collect(TGeoVolume* v, std::vector<TGeoVolume*>& order) {
    if (volume_id_.count(v)) return;
    volume_id_[v] = (int)order.size();
    order.push_back(v);
    for (Int_t i = 0; i < v->GetNdaughters(); i++) Collect(v->GetNode(i)->GetVolume(), order);
  }
  void FlattenNodes(TGeoVolume* v) {  // the order rbg_scene_create flattens placed nodes in (DFS pre-order)
    for (Int_t i = 0; i < v->GetNdaughters(); i++) {
      node_of_id.push_back(v->GetNode(i));
      FlattenNodes(v->GetNode(i)->GetVolume());
    }
  }
  void Finish(int top) {
    std::memset(&desc, 0, sizeof(desc));
    desc.abi_version = RBG_ABI_VERSION;
    desc.top_volume = top;
#define RB_SET(field, cnt, vec) desc.cnt = (int32_t)vec.size(); desc.field = vec.empty() ? nullptr : vec.data();
    RB_SET(shapes, nshapes, shapes) RB_SET(dpar, ndpar, dpar) RB_SET(matrices, nmatrices, matrices) RB_SET(nodes, nnodes, nodes)
    RB_SET(volumes, nvolumes, volumes) RB_SET(borders, nborders, borders) RB_SET(graphs, ngraphs, graphs) RB_SET(gx, ngpts, gx)
    RB_SET(th2, nth2, th2) RB_SET(th2v, nth2v, th2v) RB_SET(indices, nindices, indices) RB_SET(mirrors, nmirrors, mirrors)
    RB_SET(focals, nfocals, focals) RB_SET(multilayers, nmultilayers, multilayers) RB_SET(layers, nlayers, layers) RB_SET(names, nnames, names)
    RB_SET(graph2d, ngraph2d, graph2ds) RB_SET(g2x, ng2pts, g2x)
#undef RB_SET
    desc.gy = gy.empty() ? nullptr : gy.data();
    desc.ntri = (int32_t)(tri.size() / 3);
    desc.tri = tri.empty() ? nullptr : tri.data();
    desc.g2y = g2y.empty() ? nullptr : g2y.data();
    desc.g2z = g2z.empty() ? nullptr : g2z.data();
  }
};

// AOpticsManager::TraceNonSequential(ARayArray&) on the GPU for an unmodified reference manager:
//     robast_b200::GpuTracer gpu(manager);          // once per closed geometry
//     gpu.TraceNonSequential(*array);                // instead of manager->TraceNonSequential(*array)
// Packs the running bucket into SoA host arrays, calls rbg_trace (or rbg_multi_trace on `ngpu` devices), writes every ray back
// (last point, direction, status, last node) and re-buckets in order as src/AOpticsManager.cxx:571-582 does.
class GpuTracer {
 public:
  explicit GpuTracer(AOpticsManager* mgr, int ngpu = 1, std::vector<Int_t> (*triangles_of)(TGraph2D*) = nullptr) : fManager(mgr) {
    const rbg_scene_desc* d = fExport.Build(mgr, triangles_of);
    if (ngpu > 1) {
      if (rbg_multi_create(d, ngpu, nullptr, &fMulti) != RBG_OK) throw std::runtime_error(rbg_last_error());
    } else if (rbg_scene_create(d, 0, &fScene) != RBG_OK) throw std::runtime_error(rbg_last_error());
  }
  ~GpuTracer() {
    if (fScene) rbg_scene_destroy(fScene);
    if (fMulti) rbg_multi_destroy(fMulti);
  }
  GpuTracer(const GpuTracer&) = delete;
  GpuTracer& operator=(const GpuTracer&) = delete;
  void SetSeed(ULong64_t seed) { fSeed = seed; fRayCounter = 0; }

  void TraceNonSequential(ARayArray& array) {
    TObjArray* running = array.GetRunning();
    const Int_t ntot = running->GetLast() + 1;
    if (ntot <= 0) return;
    // The reference suspends a ray when ray->GetNpoints() >= fLimit (src/AOpticsManager.cxx:515-517), counting the points it held
    // before the call; the device counts from 1.  Rays are therefore traced in groups of equal point count (one group — fresh
    // rays — in every tutorial), each with the limit lowered by what its rays already hold.
    std::vector<ARay*> all((size_t)ntot);
    std::map<Int_t, std::vector<size_t>> groups;
    for (Int_t i = 0; i < ntot; i++) {
      all[i] = static_cast<ARay*>(running->At(i));
      groups[all[i]->GetNpoints()].push_back((size_t)i);
    }
    std::vector<int32_t> status((size_t)ntot, RBG_RUN);
    const Int_t limit = Member<Int_t>(fManager, "AOpticsManager", "fLimit");
    for (auto& grp : groups) {
      const std::vector<size_t>& ids = grp.second;
      const size_t n = ids.size();
      std::vector<double> col(8 * n);
      std::vector<int32_t> icol(3 * n);
      for (size_t j = 0; j < n; j++) {
        ARay* ray = all[ids[j]];
        Double_t x[4], d[3];
        ray->GetLastPoint(x);
        ray->GetDirection(d);
        for (int k = 0; k < 4; k++) col[(size_t)k * n + j] = x[k];
        for (int k = 0; k < 3; k++) col[(size_t)(4 + k) * n + j] = d[k];
        col[(size_t)7 * n + j] = ray->GetLambda();
      }
      rbg_rays r;
      std::memset(&r, 0, sizeof(r));
      r.n = (int64_t)n;
      r.on_device = 0;
      double* c = col.data();
      r.x = c; r.y = c + n; r.z = c + 2 * n; r.t = c + 3 * n;
      r.dx = c + 4 * n; r.dy = c + 5 * n; r.dz = c + 6 * n; r.lambda = c + 7 * n;
      r.ox = c; r.oy = c + n; r.oz = c + 2 * n; r.ot = c + 3 * n;  // in place
      r.odx = c + 4 * n; r.ody = c + 5 * n; r.odz = c + 6 * n;
      r.status = icol.data(); r.last_node = icol.data() + n; r.npoints = icol.data() + 2 * n;
      rbg_trace_opts o;
      std::memset(&o, 0, sizeof(o));
      o.limit = grp.first > 1 ? std::max<Int_t>(2, limit - (grp.first - 1)) : limit;
      o.disable_fresnel = Member<Bool_t>(fManager, "AOpticsManager", "fDisableFresnelReflection") ? 1 : 0;
      o.quirks = RBG_QUIRKS_DEFAULT;
      o.seed = fSeed;
      o.ray_id_offset = fRayCounter;
      fRayCounter += (ULong64_t)n;
      const int rc = fMulti ? rbg_multi_trace(fMulti, &o, &r) : rbg_trace(fScene, &o, &r, nullptr);
      if (rc != RBG_OK) throw std::runtime_error(std::string("GpuTracer: ") + rbg_last_error());
      for (size_t j = 0; j < n; j++) {
        ARay* ray = all[ids[j]];
        if (icol[2 * n + j] > 1) {  // the last point (the polyline in between is kept only by rbg_trace_history)
          ray->AddPoint(col[j], col[n + j], col[2 * n + j], col[3 * n + j]);
          const int32_t node = icol[n + j];
          if (node >= 0 && node < (int32_t)fExport.node_of_id.size()) ray->AddNode(fExport.node_of_id[node]);
        }
        ray->SetDirection(col[4 * n + j], col[5 * n + j], col[6 * n + j]);
        status[ids[j]] = icol[j];
      }
    }
    // every ray leaves the running bucket; ARayArray::Add routes it by status and keeps the relative order (:571-582)
    for (Int_t i = 0; i < ntot; i++) running->RemoveAt(i);
    running->Expand(0);
    for (Int_t i = 0; i < ntot; i++) {
      ARay* ray = all[i];
      switch (status[i]) {
        case RBG_STOP: ray->Stop(); break;
        case RBG_EXIT: ray->Exit(); break;
        case RBG_FOCUSED: ray->Focus(); break;
        case RBG_SUSPEND: ray->Suspend(); break;
        case RBG_ABSORB: ray->Absorb(); break;
        default: break;
      }
      array.Add(ray);
    }
  }

 private:
  AOpticsManager* fManager;
  RootExporter fExport;
  rbg_scene* fScene = nullptr;
  rbg_multi* fMulti = nullptr;
  ULong64_t fSeed = 20180601ULL;
  ULong64_t fRayCounter = 0;
};

}  // namespace robast_b200

#endif  // ROBAST_HAVE_ROOT
#endif  // ROBAST_ROOT_EXPORTER_H
