// pybind.cpp — Python bindings of the host-side ROBAST mirror (include/robast/Robast.h), so that
// scripts written for PyROOT + ROBAST (e.g. tutorials/unittest_robast.py, SimpleParabolicTelescope.py
// of the reference) run with `import robast_b200 as ROOT`.  Geometry objects follow ROOT ownership
// (never deleted from Python); optical data use shared_ptr like the reference API.
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/complex.h>
#include <pybind11/stl.h>

#include "../../include/robast/Robast.h"

namespace py = pybind11;
template <class T> using Raw = std::unique_ptr<T, py::nodelete>;

static py::array_t<double> view_d(std::vector<double>& v, py::object owner) { return py::array_t<double>({(py::ssize_t)v.size()}, {sizeof(double)}, v.data(), owner); }
static py::array_t<int32_t> view_i(std::vector<int32_t>& v, py::object owner) { return py::array_t<int32_t>({(py::ssize_t)v.size()}, {sizeof(int32_t)}, v.data(), owner); }


PYBIND11_MODULE(_robast, m) {
  m.doc() = "robast_b200 host layer (ROOT-compat subset + ROBAST classes) over the CUDA C ABI";

  // ---- TMath-ish helpers and units live on AOpticsManager like the reference
  py::class_<TObject, Raw<TObject>>(m, "TObject").def("GetName", &TObject::GetName);
  py::class_<TNamed, TObject, Raw<TNamed>>(m, "TNamed").def("GetTitle", &TNamed::GetTitle).def("SetName", &TNamed::SetName);

  py::class_<TVector3>(m, "TVector3")
      .def(py::init<double, double, double>(), py::arg("x") = 0., py::arg("y") = 0., py::arg("z") = 0.)
      .def("X", &TVector3::X).def("Y", &TVector3::Y).def("Z", &TVector3::Z)
      .def("SetXYZ", &TVector3::SetXYZ).def("SetMagThetaPhi", &TVector3::SetMagThetaPhi)
      .def("Theta", &TVector3::Theta).def("Phi", &TVector3::Phi).def("Mag", &TVector3::Mag).def("Unit", &TVector3::Unit)
      .def("RotateZ", &TVector3::RotateZ).def("RotateX", &TVector3::RotateX).def("RotateY", &TVector3::RotateY)
      .def("Angle", &TVector3::Angle).def("Dot", &TVector3::Dot)
      .def("__getitem__", [](const TVector3& v, int i) { return v[i]; });

  py::class_<TObjArray, TObject, Raw<TObjArray>>(m, "TObjArray")
      .def(py::init<>())
      .def("Add", [](TObjArray& a, TObject* o) { a.Add(o); }, py::keep_alive<1, 2>())
      .def("GetLast", &TObjArray::GetLast).def("GetEntries", &TObjArray::GetEntries).def("GetEntriesFast", &TObjArray::GetEntriesFast)
      .def("At", &TObjArray::At, py::return_value_policy::reference)
      .def("__getitem__", &TObjArray::At, py::return_value_policy::reference)
      .def("__len__", [](const TObjArray& a) { return a.GetLast() + 1; });

  py::class_<TRandom, TObject, Raw<TRandom>>(m, "TRandom")
      .def("SetSeed", &TRandom::SetSeed, py::arg("seed") = 0)
      .def("Rndm", &TRandom::Rndm)
      .def("Uniform", (Double_t(TRandom::*)(Double_t)) & TRandom::Uniform, py::arg("x1") = 1.)
      .def("Uniform", (Double_t(TRandom::*)(Double_t, Double_t)) & TRandom::Uniform)
      .def("Gaus", &TRandom::Gaus, py::arg("mean") = 0., py::arg("sigma") = 1.)
      .def("Exp", &TRandom::Exp);
  m.attr("gRandom") = py::cast(gRandom, py::return_value_policy::reference);

  // ---- matrices
  py::class_<TGeoMatrix, TNamed, Raw<TGeoMatrix>>(m, "TGeoMatrix")
      .def("RegisterYourself", &TGeoMatrix::RegisterYourself)
      .def("GetRotationMatrix", [](const TGeoMatrix& t) { return std::vector<double>(t.GetRotationMatrix(), t.GetRotationMatrix() + 9); })
      .def("GetTranslation", [](const TGeoMatrix& t) { return std::vector<double>(t.GetTranslation(), t.GetTranslation() + 3); })
      .def("LocalToMaster", [](const TGeoMatrix& t, std::array<double, 3> l) { std::array<double, 3> o; t.LocalToMaster(l.data(), o.data()); return o; })
      .def("MasterToLocal", [](const TGeoMatrix& t, std::array<double, 3> l) { std::array<double, 3> o; t.MasterToLocal(l.data(), o.data()); return o; })
      .def("LocalToMasterVect", [](const TGeoMatrix& t, std::array<double, 3> l) { std::array<double, 3> o; t.LocalToMasterVect(l.data(), o.data()); return o; });
  py::class_<TGeoTranslation, TGeoMatrix, Raw<TGeoTranslation>>(m, "TGeoTranslation")
      .def(py::init<double, double, double>())
      .def(py::init<const char*, double, double, double>())
      .def("SetTranslation", &TGeoTranslation::SetTranslation);
  py::class_<TGeoRotation, TGeoMatrix, Raw<TGeoRotation>>(m, "TGeoRotation")
      .def(py::init<>())
      .def(py::init<const char*>())
      .def(py::init<const char*, double, double, double>())
      .def("SetAngles", &TGeoRotation::SetAngles)
      .def("MultiplyBy", &TGeoRotation::MultiplyBy, py::arg("rot"), py::arg("after") = true)
      .def("RotateX", &TGeoRotation::RotateX).def("RotateY", &TGeoRotation::RotateY).def("RotateZ", &TGeoRotation::RotateZ);
  py::class_<TGeoCombiTrans, TGeoMatrix, Raw<TGeoCombiTrans>>(m, "TGeoCombiTrans")
      .def(py::init<const TGeoTranslation&, const TGeoRotation&>())
      .def(py::init<double, double, double, TGeoRotation*>())
      .def(py::init<const char*, double, double, double, TGeoRotation*>());
  py::class_<TGeoHMatrix, TGeoMatrix, Raw<TGeoHMatrix>>(m, "TGeoHMatrix")
      .def(py::init<>())
      .def(py::init<const TGeoMatrix&>())
      .def("__mul__", [](const TGeoHMatrix& a, const TGeoMatrix& b) { return new TGeoHMatrix(a * b); }, py::return_value_policy::reference);

  // ---- shapes
  py::class_<TGeoShape, TNamed, Raw<TGeoShape>>(m, "TGeoShape").def_static("Big", &TGeoShape::Big).def_static("Tolerance", &TGeoShape::Tolerance);
  py::class_<TGeoBBox, TGeoShape, Raw<TGeoBBox>>(m, "TGeoBBox")
      .def(py::init<const char*, double, double, double>())
      .def(py::init<double, double, double>())
      .def(py::init([](const char* n, double dx, double dy, double dz, std::array<double, 3> o) { return new TGeoBBox(n, dx, dy, dz, o.data()); }))
      .def("GetDX", &TGeoBBox::GetDX).def("GetDY", &TGeoBBox::GetDY).def("GetDZ", &TGeoBBox::GetDZ)
      .def("GetOrigin", [](const TGeoBBox& b) { return std::vector<double>(b.GetOrigin(), b.GetOrigin() + 3); });
  py::class_<TGeoTube, TGeoBBox, Raw<TGeoTube>>(m, "TGeoTube").def(py::init<const char*, double, double, double>()).def(py::init<double, double, double>());
  py::class_<TGeoSphere, TGeoBBox, Raw<TGeoSphere>>(m, "TGeoSphere")
      .def(py::init<const char*, double, double, double, double, double, double>(), py::arg("name"), py::arg("rmin"), py::arg("rmax"), py::arg("theta1") = 0.,
           py::arg("theta2") = 180., py::arg("phi1") = 0., py::arg("phi2") = 360.)
      .def(py::init<double, double, double, double, double, double>(), py::arg("rmin"), py::arg("rmax"), py::arg("theta1") = 0., py::arg("theta2") = 180.,
           py::arg("phi1") = 0., py::arg("phi2") = 360.);
  py::class_<TGeoParaboloid, TGeoBBox, Raw<TGeoParaboloid>>(m, "TGeoParaboloid").def(py::init<const char*, double, double, double>());
  py::class_<TGeoPcon, TGeoBBox, Raw<TGeoPcon>>(m, "TGeoPcon").def(py::init<const char*, double, double, int>()).def("DefineSection", &TGeoPcon::DefineSection);
  py::class_<TGeoPgon, TGeoPcon, Raw<TGeoPgon>>(m, "TGeoPgon").def(py::init<const char*, double, double, int, int>());
  py::class_<TGeoArb8, TGeoBBox, Raw<TGeoArb8>>(m, "TGeoArb8")
      .def(py::init([](const char* name, double dz, std::vector<double> v) {
             if (!v.empty() && v.size() != 16) throw std::runtime_error("TGeoArb8: 16 vertex coordinates expected");
             return new TGeoArb8(name, dz, v.empty() ? nullptr : v.data());
           }), py::arg("name"), py::arg("dz"), py::arg("vertices") = std::vector<double>())
      .def("SetVertex", &TGeoArb8::SetVertex);
  py::class_<TGeoXtru, TGeoBBox, Raw<TGeoXtru>>(m, "TGeoXtru")
      .def(py::init<int>())
      .def("SetName", [](TGeoXtru& x, const char* n) { x.SetName(n); })
      .def("DefinePolygon", [](TGeoXtru& x, std::vector<double> xv, std::vector<double> yv) {
        if (xv.size() != yv.size()) throw std::runtime_error("TGeoXtru::DefinePolygon: x and y differ in length");
        return (bool)x.DefinePolygon((int)xv.size(), xv.data(), yv.data());
      })
      .def("DefineSection", &TGeoXtru::DefineSection, py::arg("snum"), py::arg("z"), py::arg("x0") = 0., py::arg("y0") = 0., py::arg("scale") = 1.);
  py::class_<TGeoCompositeShape, TGeoBBox, Raw<TGeoCompositeShape>>(m, "TGeoCompositeShape").def(py::init<const char*, const char*>());
  py::class_<AGeoAsphericDisk, TGeoBBox, Raw<AGeoAsphericDisk>>(m, "AGeoAsphericDisk")
      .def(py::init<const char*, double, double, double, double, double, double>(), py::arg("name"), py::arg("z1"), py::arg("curve1"), py::arg("z2"),
           py::arg("curve2"), py::arg("rmax"), py::arg("rmin") = 0.)
      .def("SetPolynomials", [](AGeoAsphericDisk& a, int n1, std::vector<double> k1, int n2, std::vector<double> k2) {
        a.SetPolynomials(n1, k1.data(), n2, k2.data());
      })
      .def("SetConicConstants", &AGeoAsphericDisk::SetConicConstants).def("SetFineness", &AGeoAsphericDisk::SetFineness)
      .def("CalcF1", &AGeoAsphericDisk::CalcF1).def("CalcF2", &AGeoAsphericDisk::CalcF2)
      .def("CalcdF1dr", &AGeoAsphericDisk::CalcdF1dr).def("CalcdF2dr", &AGeoAsphericDisk::CalcdF2dr)
      .def("GetRmax", &AGeoAsphericDisk::GetRmax).def("GetRmin", &AGeoAsphericDisk::GetRmin)
      .def("GetZ1", &AGeoAsphericDisk::GetZ1).def("GetZ2", &AGeoAsphericDisk::GetZ2);
  py::class_<AGeoWinstonCone2D, TGeoBBox, Raw<AGeoWinstonCone2D>>(m, "AGeoWinstonCone2D")
      .def(py::init<const char*, double, double, double>())
      .def("CalcR", &AGeoWinstonCone2D::CalcR).def("CalcdRdZ", &AGeoWinstonCone2D::CalcdRdZ)
      .def("GetTheta", &AGeoWinstonCone2D::GetTheta).def("GetR1", &AGeoWinstonCone2D::GetR1).def("GetR2", &AGeoWinstonCone2D::GetR2).def("GetF", &AGeoWinstonCone2D::GetF);
  py::class_<AGeoWinstonConePoly, AGeoWinstonCone2D, Raw<AGeoWinstonConePoly>>(m, "AGeoWinstonConePoly").def(py::init<const char*, double, double, int>());
  py::class_<AGeoBezierPgon, TGeoPgon, Raw<AGeoBezierPgon>>(m, "AGeoBezierPgon")
      .def(py::init<const char*, double, double, int, int, double, double, double>())
      .def("SetControlPoints", (void(AGeoBezierPgon::*)(double, double)) & AGeoBezierPgon::SetControlPoints)
      .def("SetControlPoints", (void(AGeoBezierPgon::*)(double, double, double, double)) & AGeoBezierPgon::SetControlPoints);
  py::class_<AGeoBezierPcon, TGeoPcon, Raw<AGeoBezierPcon>>(m, "AGeoBezierPcon")
      .def(py::init<const char*, double, double, int, double, double, double>())
      .def("SetControlPoints", (void(AGeoBezierPcon::*)(double, double)) & AGeoBezierPcon::SetControlPoints)
      .def("SetControlPoints", (void(AGeoBezierPcon::*)(double, double, double, double)) & AGeoBezierPcon::SetControlPoints);
  m.def("RbGeomEpoch", []() { return (unsigned long long)RbGeomEpoch().load(); },
        "geometry epoch: bumped by every mutator of a shape / matrix / volume / table / optical property (scene cache of AOpticsManager)");
  m.def("ContainmentRadius", [](TH2* h, double fraction) {
    double r, x, y;
    AGeoUtil::ContainmentRadius(h, fraction, r, x, y);
    return py::make_tuple(r, x, y);
  });
  m.def("MakeArb8FromPoints", [](const char* name, TVector3 v1, TVector3 v2, TVector3 v3, TVector3 v4, TVector3 v5) {
    TGeoArb8* a;
    TGeoCombiTrans* c;
    AGeoUtil::MakeArb8FromPoints(name, v1, v2, v3, v4, v5, &a, &c);
    return py::make_tuple(py::cast(a, py::return_value_policy::reference), py::cast(c, py::return_value_policy::reference));
  });
  m.def("MakeXtruFromPoints", [](const char* name, std::vector<TVector3> vecs) {
    if (vecs.size() < 4) throw std::runtime_error("MakeXtruFromPoints: at least three top corners and the bottom corner");
    TGeoXtru* x;
    TGeoCombiTrans* c;
    AGeoUtil::MakeXtruFromPoints(name, vecs, &x, &c);
    return py::make_tuple(py::cast(x, py::return_value_policy::reference), py::cast(c, py::return_value_policy::reference));
  });
  m.def("MakePointToPointTube", [](const char* name, TVector3 v1, TVector3 v2, double radius) {
    TGeoTube* t;
    TGeoCombiTrans* c;
    AGeoUtil::MakePointToPointTube(name, v1, v2, radius, &t, &c);
    return py::make_tuple(py::cast(t, py::return_value_policy::reference), py::cast(c, py::return_value_policy::reference));
  });

  // ---- graphs / histograms
  py::class_<TGraph, std::shared_ptr<TGraph>>(m, "TGraph")
      .def(py::init<>())
      .def("SetPoint", &TGraph::SetPoint).def("GetN", &TGraph::GetN).def("Eval", &TGraph::Eval);
  py::class_<TGraph2D, std::shared_ptr<TGraph2D>>(m, "TGraph2D").def(py::init<>()).def("SetPoint", &TGraph2D::SetPoint).def("GetN", &TGraph2D::GetN)
      .def("Interpolate", &TGraph2D::Interpolate).def("GetTriangles", [](const TGraph2D& g) { return g.GetTriangles(); });
  py::class_<TH1, std::shared_ptr<TH1>>(m, "TH1").def("GetEntries", &TH1::GetEntries);
  py::class_<TH1D, TH1, std::shared_ptr<TH1D>>(m, "TH1D")
      .def(py::init<const char*, const char*, int, double, double>())
      .def("Fill", &TH1D::Fill, py::arg("x"), py::arg("w") = 1.)
      .def("GetMean", &TH1D::GetMean, py::arg("axis") = 1).def("GetRMS", &TH1D::GetRMS, py::arg("axis") = 1).def("GetStdDev", &TH1D::GetStdDev, py::arg("axis") = 1)
      .def("GetBinContent", &TH1D::GetBinContent).def("GetNbinsX", &TH1D::GetNbinsX).def("GetBinCenter", &TH1D::GetBinCenter);
  py::class_<TH2, TH1, std::shared_ptr<TH2>>(m, "TH2")
      .def("Fill", &TH2::Fill, py::arg("x"), py::arg("y"), py::arg("w") = 1.)
      .def("GetMean", &TH2::GetMean, py::arg("axis") = 1).def("GetRMS", &TH2::GetRMS, py::arg("axis") = 1).def("GetStdDev", &TH2::GetStdDev, py::arg("axis") = 1)
      .def("GetBinContent", &TH2::GetBinContent).def("SetBinContent", &TH2::SetBinContent).def("Interpolate", &TH2::Interpolate)
      .def("GetNbinsX", &TH2::GetNbinsX).def("GetNbinsY", &TH2::GetNbinsY);
  py::class_<TH2D, TH2, std::shared_ptr<TH2D>>(m, "TH2D").def(py::init<const char*, const char*, int, double, double, int, double, double>());

  // ---- refractive indices
  py::class_<ARefractiveIndex, std::shared_ptr<ARefractiveIndex>>(m, "ARefractiveIndex")
      .def(py::init<>())
      .def(py::init<double, double>(), py::arg("n"), py::arg("k") = 0.)
      .def("GetAbbeNumber", &ARefractiveIndex::GetAbbeNumber)
      .def("GetRefractiveIndex", &ARefractiveIndex::GetRefractiveIndex)
      .def("GetExtinctionCoefficient", &ARefractiveIndex::GetExtinctionCoefficient)
      .def("GetAbsorptionLength", &ARefractiveIndex::GetAbsorptionLength)
      .def("SetRefractiveIndex", &ARefractiveIndex::SetRefractiveIndex)
      .def("SetExtinctionCoefficient", &ARefractiveIndex::SetExtinctionCoefficient)
      .def_static("AbsorptionLengthToExtinctionCoefficient", &ARefractiveIndex::AbsorptionLengthToExtinctionCoefficient)
      .def_static("ExtinctionCoefficientToAbsorptionLength", &ARefractiveIndex::ExtinctionCoefficientToAbsorptionLength);
  py::class_<ASellmeierFormula, ARefractiveIndex, std::shared_ptr<ASellmeierFormula>>(m, "ASellmeierFormula").def(py::init<double, double, double, double, double, double>());
  py::class_<ASchottFormula, ARefractiveIndex, std::shared_ptr<ASchottFormula>>(m, "ASchottFormula").def(py::init<double, double, double, double, double, double>());
  py::class_<ACauchyFormula, ARefractiveIndex, std::shared_ptr<ACauchyFormula>>(m, "ACauchyFormula").def(py::init<double, double, double>());
  py::class_<AMixedRefractiveIndex, ARefractiveIndex, std::shared_ptr<AMixedRefractiveIndex>>(m, "AMixedRefractiveIndex")
      .def(py::init<std::shared_ptr<ARefractiveIndex>, std::shared_ptr<ARefractiveIndex>, double, double>())
      .def("SetFraction", &AMixedRefractiveIndex::SetFraction);
  py::class_<AFilmetrixDotCom, ARefractiveIndex, std::shared_ptr<AFilmetrixDotCom>>(m, "AFilmetrixDotCom").def(py::init<const char*>());
  py::class_<ARefractiveIndexDotInfo, ARefractiveIndex, std::shared_ptr<ARefractiveIndexDotInfo>>(m, "ARefractiveIndexDotInfo").def(py::init<const char*>());
  py::class_<AGlassCatalog>(m, "AGlassCatalog").def(py::init<const std::string&>()).def("GetRefractiveIndex", &AGlassCatalog::GetRefractiveIndex);

  py::class_<AMultilayer, std::shared_ptr<AMultilayer>> ml(m, "AMultilayer");
  ml.def(py::init<std::shared_ptr<ARefractiveIndex>, std::shared_ptr<ARefractiveIndex>>())
      .def("AddLayer", &AMultilayer::AddLayer, py::arg("idx"), py::arg("thickness"), py::arg("coherent") = true)
      .def("InsertLayer", &AMultilayer::InsertLayer, py::arg("idx"), py::arg("thickness"), py::arg("coherent") = true)
      .def("ChangeThickness", &AMultilayer::ChangeThickness).def("GetThickness", &AMultilayer::GetThickness)
      .def("CoherentTMMMixed", [](const AMultilayer& a, double th, double lam) { double r, t; a.CoherentTMMMixed(th, lam, r, t); return py::make_tuple(r, t); })
      .def("PreCalculateCoherentTMM", &AMultilayer::PreCalculateCoherentTMM)
      .def("PreCalculateIncoherentTMM", &AMultilayer::PreCalculateIncoherentTMM)
      .def("CoherentTMM", [](const AMultilayer& a, int pol, std::complex<double> th, double lam, bool reverse) {
        double r, t;
        a.CoherentTMM(pol == 0 ? AMultilayer::kS : AMultilayer::kP, th, lam, r, t, reverse);
        return py::make_tuple(r, t);
      }, py::arg("pol"), py::arg("th_0"), py::arg("lam_vac"), py::arg("reverse") = false)
      .def("IncoherentTMM", [](const AMultilayer& a, int pol, std::complex<double> th, double lam) {
        double r, t;
        a.IncoherentTMM(pol == 0 ? AMultilayer::kS : AMultilayer::kP, th, lam, r, t);
        return py::make_tuple(r, t);
      })
      .def("IncoherentTMMMixed", [](const AMultilayer& a, std::complex<double> th, double lam) {
        double r, t;
        a.IncoherentTMMMixed(th, lam, r, t);
        return py::make_tuple(r, t);
      })
      .def_property_readonly_static("kS", [](py::object) { return 0; })
      .def_property_readonly_static("kP", [](py::object) { return 1; });

  // ---- volumes
  py::class_<TGeoNode, TNamed, Raw<TGeoNode>>(m, "TGeoNode");
  py::class_<TGeoVolume, TNamed, Raw<TGeoVolume>>(m, "TGeoVolume")
      .def("AddNode", [](TGeoVolume& v, TGeoVolume* d, int copy, TGeoMatrix* mat) { v.AddNode(d, copy, mat); }, py::arg("vol"), py::arg("copy_no"), py::arg("mat") = nullptr)
      .def("AddNodeOverlap", [](TGeoVolume& v, TGeoVolume* d, int copy, TGeoMatrix* mat) { v.AddNodeOverlap(d, copy, mat); }, py::arg("vol"), py::arg("copy_no"),
           py::arg("mat") = nullptr)
      .def("GetNdaughters", &TGeoVolume::GetNdaughters);
  py::class_<AOpticalComponent, TGeoVolume, Raw<AOpticalComponent>>(m, "AOpticalComponent").def(py::init<const char*, const TGeoShape*>());
  py::class_<ALens, AOpticalComponent, Raw<ALens>>(m, "ALens")
      .def(py::init<const char*, const TGeoShape*>())
      .def("SetRefractiveIndex", &ALens::SetRefractiveIndex)
      .def("GetRefractiveIndex", &ALens::GetRefractiveIndex).def("GetAbsorptionLength", &ALens::GetAbsorptionLength);
  py::class_<AMirror, AOpticalComponent, Raw<AMirror>>(m, "AMirror")
      .def(py::init<const char*, const TGeoShape*>())
      .def("SetReflectance", (void(AMirror::*)(double)) & AMirror::SetReflectance)
      .def("SetReflectance", (void(AMirror::*)(std::shared_ptr<TGraph>)) & AMirror::SetReflectance)
      .def("SetReflectance", (void(AMirror::*)(std::shared_ptr<TH2>)) & AMirror::SetReflectance)
      .def("SetReflectance", (void(AMirror::*)(std::shared_ptr<TGraph2D>)) & AMirror::SetReflectance)
      .def("GetReflectance", &AMirror::GetReflectance);
  py::class_<AFocalSurface, AOpticalComponent, Raw<AFocalSurface>>(m, "AFocalSurface")
      .def(py::init<const char*, const TGeoShape*>())
      .def("SetQuantumEfficiency", [](AFocalSurface& f, std::shared_ptr<TGraph> g) { f.SetQuantumEfficiency(g.get()); }, py::keep_alive<1, 2>())
      .def("SetQuantumEfficiencyAngle", [](AFocalSurface& f, std::shared_ptr<TGraph> g) { f.SetQuantumEfficiencyAngle(g.get()); }, py::keep_alive<1, 2>());
  py::class_<AObscuration, AOpticalComponent, Raw<AObscuration>>(m, "AObscuration").def(py::init<const char*, const TGeoShape*>());
  py::class_<ABorderSurfaceCondition, TObject, Raw<ABorderSurfaceCondition>>(m, "ABorderSurfaceCondition")
      .def(py::init<AOpticalComponent*, AOpticalComponent*>())
      .def("SetGaussianRoughness", &ABorderSurfaceCondition::SetGaussianRoughness).def("GetGaussianRoughness", &ABorderSurfaceCondition::GetGaussianRoughness)
      .def("SetMultilayer", &ABorderSurfaceCondition::SetMultilayer).def("EnableLambertian", &ABorderSurfaceCondition::EnableLambertian)
      .def("IsLambertian", &ABorderSurfaceCondition::IsLambertian);

  // ---- rays
  py::class_<ARay, TObject, Raw<ARay>>(m, "ARay")
      .def(py::init<int, double, double, double, double, double, double, double, double>())
      .def("GetDirection", [](const ARay& r) { std::array<double, 3> d; r.GetDirection(d.data()); return d; })
      .def("GetLastPoint", [](const ARay& r) { std::array<double, 4> p; r.GetLastPoint(p.data()); return p; })
      .def("GetFirstPoint", [](const ARay& r) { return std::vector<double>(r.GetFirstPoint(), r.GetFirstPoint() + 4); })
      .def("GetNpoints", &ARay::GetNpoints).def("GetLambda", &ARay::GetLambda).def("GetStatus", &ARay::GetStatus)
      .def("AddPoint", &ARay::AddPoint)
      .def("GetLastNodeName", &ARay::GetLastNodeName)
      .def("GetPoint", [](const ARay& r, int i) { const double* p = r.GetPoint(i); return std::vector<double>(p, p + 4); })
      .def("GetNrecorded", &ARay::GetNrecorded)
      .def("FindNodeNumberStartWith", &ARay::FindNodeNumberStartWith)
      .def("GetNodeHistoryNames", [](const ARay& r) {
        std::vector<std::string> v;
        const TObjArray* h = r.GetNodeHistory();
        Int_t n = std::max(h->GetLast() + 1, r.GetNnodesRecorded());
        for (Int_t i = 0; i < n; i++) v.push_back(h->At(i) ? h->At(i)->GetName() : "");
        return v;
      })
      .def("IsAbsorbed", &ARay::IsAbsorbed).def("IsExited", &ARay::IsExited).def("IsFocused", &ARay::IsFocused).def("IsRunning", &ARay::IsRunning)
      .def("IsStopped", &ARay::IsStopped).def("IsSuspended", &ARay::IsSuspended);
  py::class_<ARayArray>(m, "ARayArray")
      .def(py::init<>())
      .def("Add", [](ARayArray& a, ARay* r) { a.Add(r); })
      .def("AddRaw", &ARayArray::AddRaw)
      .def("AddRays", [](ARayArray& a, py::array_t<double, py::array::c_style | py::array::forcecast> rays) {
        // rays: (n, 8) = x,y,z,t,dx,dy,dz,lambda
        auto r = rays.unchecked<2>();
        if (r.shape(1) != 8) throw std::runtime_error("AddRays expects an (n, 8) array");
        a.Reserve(a.GetN() + r.shape(0));
        for (py::ssize_t i = 0; i < r.shape(0); i++) a.AddRaw(r(i, 0), r(i, 1), r(i, 2), r(i, 3), r(i, 4), r(i, 5), r(i, 6), r(i, 7));
      })
      .def("Merge", &ARayArray::Merge)
      .def("GetAbsorbed", &ARayArray::GetAbsorbed, py::return_value_policy::reference_internal)
      .def("GetExited", &ARayArray::GetExited, py::return_value_policy::reference_internal)
      .def("GetFocused", &ARayArray::GetFocused, py::return_value_policy::reference_internal)
      .def("GetRunning", &ARayArray::GetRunning, py::return_value_policy::reference_internal)
      .def("GetStopped", &ARayArray::GetStopped, py::return_value_policy::reference_internal)
      .def("GetSuspended", &ARayArray::GetSuspended, py::return_value_policy::reference_internal)
      .def("GetN", &ARayArray::GetN).def("Count", &ARayArray::Count)
      .def("columns", [](py::object self) {
        ARayArray& a = self.cast<ARayArray&>();
        ARayArray::Table& T = a.GetTable();
        py::dict d;
        d["x0"] = view_d(T.x0, self); d["y0"] = view_d(T.y0, self); d["z0"] = view_d(T.z0, self); d["t0"] = view_d(T.t0, self);
        d["x"] = view_d(T.x, self); d["y"] = view_d(T.y, self); d["z"] = view_d(T.z, self); d["t"] = view_d(T.t, self);
        d["dx"] = view_d(T.dx, self); d["dy"] = view_d(T.dy, self); d["dz"] = view_d(T.dz, self); d["lambda"] = view_d(T.lambda, self);
        d["status"] = view_i(T.status, self); d["npoints"] = view_i(T.npoints, self); d["last_node"] = view_i(T.last_node, self);
        return d;
      });
  py::class_<ARayShooter>(m, "ARayShooter")
      .def_static("Circle", &ARayShooter::Circle, py::arg("lambda"), py::arg("rmax"), py::arg("nr"), py::arg("nphi"), py::arg("rot") = nullptr, py::arg("tr") = nullptr,
                  py::arg("v") = nullptr, py::return_value_policy::take_ownership)
      .def_static("RandomCircle", &ARayShooter::RandomCircle, py::arg("lambda"), py::arg("rmax"), py::arg("n"), py::arg("rot") = nullptr, py::arg("tr") = nullptr,
                  py::arg("v") = nullptr, py::return_value_policy::take_ownership)
      .def_static("RandomCone", &ARayShooter::RandomCone, py::arg("lambda"), py::arg("r"), py::arg("d"), py::arg("n"), py::arg("rot") = nullptr, py::arg("tr") = nullptr,
                  py::return_value_policy::take_ownership)
      .def_static("RandomRectangle", &ARayShooter::RandomRectangle, py::arg("lambda"), py::arg("dx"), py::arg("dy"), py::arg("n"), py::arg("rot") = nullptr,
                  py::arg("tr") = nullptr, py::arg("v") = nullptr, py::return_value_policy::take_ownership)
      .def_static("RandomSphere", &ARayShooter::RandomSphere, py::arg("lambda"), py::arg("n"), py::arg("tr") = nullptr, py::return_value_policy::take_ownership)
      .def_static("RandomSphericalCone", &ARayShooter::RandomSphericalCone, py::arg("lambda"), py::arg("n"), py::arg("theta"), py::arg("rot") = nullptr,
                  py::arg("tr") = nullptr, py::return_value_policy::take_ownership)
      .def_static("RandomSquare", &ARayShooter::RandomSquare, py::arg("lambda"), py::arg("d"), py::arg("n"), py::arg("rot") = nullptr, py::arg("tr") = nullptr,
                  py::arg("v") = nullptr, py::return_value_policy::take_ownership)
      .def_static("Rectangle", &ARayShooter::Rectangle, py::arg("lambda"), py::arg("dx"), py::arg("dy"), py::arg("nx"), py::arg("ny"), py::arg("rot") = nullptr,
                  py::arg("tr") = nullptr, py::arg("v") = nullptr, py::return_value_policy::take_ownership)
      .def_static("Square", &ARayShooter::Square, py::arg("lambda"), py::arg("d"), py::arg("n"), py::arg("rot") = nullptr, py::arg("tr") = nullptr,
                  py::arg("v") = nullptr, py::return_value_policy::take_ownership);

  // ---- scene export handle (address of the flat rbg_scene_desc for ctypes callers)
  py::class_<ASceneExport, std::shared_ptr<ASceneExport>>(m, "ASceneExport")
      .def("desc_ptr", [](ASceneExport& e) { return (uintptr_t)&e.desc; })
      .def("num_shapes", [](ASceneExport& e) { return e.shapes.size(); })
      .def("num_volumes", [](ASceneExport& e) { return e.volumes.size(); })
      .def("num_nodes", [](ASceneExport& e) { return e.nodes.size(); })
      .def("shape_of_volume", [](ASceneExport& e, int v) { return e.volumes.at(v).shape; })
      .def("matrix", [](ASceneExport& e, int i) {
        const rbg_matrix& mm = e.matrices.at(i);
        return py::make_tuple(std::vector<double>(mm.rot, mm.rot + 9), std::vector<double>(mm.tr, mm.tr + 3));
      })
      .def("node", [](ASceneExport& e, int i) { const rbg_node& n = e.nodes.at(i); return py::make_tuple(n.volume, n.matrix, n.copy_no, n.overlap); });
  m.def("export_multilayer", [](std::shared_ptr<AMultilayer> ml) {
    auto e = std::make_shared<ASceneExport>();
    int id = e->AddMultilayer(ml.get());
    e->Finish(-1);
    return py::make_tuple(e, id);
  });
  m.def("export_index", [](std::shared_ptr<ARefractiveIndex> ix) {
    auto e = std::make_shared<ASceneExport>();
    int id = e->AddIndex(ix.get());
    e->Finish(-1);
    return py::make_tuple(e, id);
  });
  m.def("export_graph", [](std::shared_ptr<TGraph> g) {
    auto e = std::make_shared<ASceneExport>();
    int id = e->AddGraph(g.get());
    e->Finish(-1);
    return py::make_tuple(e, id);
  });
  m.def("export_th2", [](std::shared_ptr<TH2> h) {
    auto e = std::make_shared<ASceneExport>();
    int id = e->AddTH2(h.get());
    e->Finish(-1);
    return py::make_tuple(e, id);
  });

  // ---- manager
  py::class_<TGeoManager, TNamed, Raw<TGeoManager>>(m, "TGeoManager")
      .def("SetTopVolume", &TGeoManager::SetTopVolume).def("GetTopVolume", &TGeoManager::GetTopVolume, py::return_value_policy::reference)
      .def("CloseGeometry", [](TGeoManager& g) { g.CloseGeometry(); })
      .def("SetNsegments", &TGeoManager::SetNsegments).def("SetMaxThreads", &TGeoManager::SetMaxThreads).def("GetMaxThreads", &TGeoManager::GetMaxThreads)
      .def("SetMultiThread", &TGeoManager::SetMultiThread, py::arg("flag") = true).def("IsMultiThread", &TGeoManager::IsMultiThread);
  py::class_<AOpticsManager, TGeoManager, Raw<AOpticsManager>>(m, "AOpticsManager")
      .def(py::init<const char*, const char*>())
      .def_static("km", &AOpticsManager::km).def_static("m", &AOpticsManager::m).def_static("cm", &AOpticsManager::cm).def_static("mm", &AOpticsManager::mm)
      .def_static("um", &AOpticsManager::um).def_static("nm", &AOpticsManager::nm).def_static("inch", &AOpticsManager::inch).def_static("s", &AOpticsManager::s)
      .def_static("ms", &AOpticsManager::ms).def_static("us", &AOpticsManager::us).def_static("ns", &AOpticsManager::ns).def_static("deg", &AOpticsManager::deg)
      .def_static("rad", &AOpticsManager::rad)
      .def("DisableFresnelReflection", &AOpticsManager::DisableFresnelReflection)
      .def("SetLimit", &AOpticsManager::SetLimit).def("GetLimit", &AOpticsManager::GetLimit)
      .def("SetSeed", &AOpticsManager::SetSeed).def("SetQuirks", &AOpticsManager::SetQuirks).def("SetDevice", &AOpticsManager::SetDevice)
      .def("SetHistoryDepth", &AOpticsManager::SetHistoryDepth).def("GetHistoryDepth", &AOpticsManager::GetHistoryDepth)
      .def("ExportScene", &AOpticsManager::ExportScene)
      .def("InvalidateScene", &AOpticsManager::InvalidateScene).def("GetNumberOfGPUs", &AOpticsManager::GetNumberOfGPUs)
      .def("TraceNonSequential", [](AOpticsManager& mg, TObjArray* a) { mg.TraceNonSequential(a); })
      .def("TraceNonSequential", [](AOpticsManager& mg, ARayArray& a) { py::gil_scoped_release rel; mg.TraceNonSequential(a); })
      .def("TraceNonSequential", [](AOpticsManager& mg, ARay& r) { mg.TraceNonSequential(r); });
}
