// Stand-in DECLARATIONS (no bodies, nothing to link) of the few ROOT and ROBAST classes include/robast/RootExporter.h touches, with
// the signatures it relies on: enough for `g++ -fsyntax-only`, so that the exporter — which can only be built for real on a host
// with CERN ROOT — at least parses and type-checks here.  Written from the public ROOT 6 API and the reference's headers
// (/root/reference/include/*.h: class names, method names and argument types only); not a ROOT replacement.
#ifndef FAKE_ROOT_DECLS_H
#define FAKE_ROOT_DECLS_H
#include <memory>
#include <vector>
typedef double Double_t;
typedef int Int_t;
typedef bool Bool_t;
typedef long Long_t;
typedef unsigned long long ULong64_t;
typedef const char Option_t;
class TClass {
 public:
  static TClass* GetClass(const char* name);
  Long_t GetDataMemberOffset(const char* name) const;
  Bool_t InheritsFrom(const TClass* c) const;
};
#define FAKE_CLASSDEF            \
  static TClass* Class();        \
  virtual TClass* IsA() const;   \
  virtual const char* ClassName() const;
class TObject {
 public:
  virtual ~TObject();
  FAKE_CLASSDEF
  Bool_t InheritsFrom(const TClass* c) const;
};
class TNamed : public TObject {
 public:
  const char* GetName() const;
};
class TObjArray : public TObject {
 public:
  Int_t GetLast() const;
  TObject* At(Int_t i) const;
  TObject* RemoveAt(Int_t i);
  void Expand(Int_t n);
  void Add(TObject* o);
};
class TAxis {
 public:
  Double_t GetXmin() const;
  Double_t GetXmax() const;
};
class TGraph : public TNamed {
 public:
  Int_t GetN() const;
  Double_t* GetX() const;
  Double_t* GetY() const;
};
class TGraph2D : public TNamed {
 public:
  Int_t GetN() const;
  Double_t* GetX() const;
  Double_t* GetY() const;
  Double_t* GetZ() const;
};
class TH2 : public TNamed {
 public:
  Int_t GetNbinsX() const;
  Int_t GetNbinsY() const;
  TAxis* GetXaxis() const;
  TAxis* GetYaxis() const;
  Double_t GetBinContent(Int_t i, Int_t j) const;
};
class TH2D : public TH2 {};
class TGeoMatrix : public TNamed {
 public:
  Bool_t IsIdentity() const;
  const Double_t* GetRotationMatrix() const;
  const Double_t* GetTranslation() const;
};
class TGeoShape : public TNamed {
 public:
  FAKE_CLASSDEF
};
class TGeoBBox : public TGeoShape {
 public:
  FAKE_CLASSDEF
  Double_t GetDX() const;
  Double_t GetDY() const;
  Double_t GetDZ() const;
  const Double_t* GetOrigin() const;
};
class TGeoTube : public TGeoBBox {
 public:
  FAKE_CLASSDEF
  Double_t GetRmin() const;
  Double_t GetRmax() const;
  Double_t GetDz() const;
};
class TGeoSphere : public TGeoBBox {
 public:
  FAKE_CLASSDEF
  Double_t GetRmin() const;
  Double_t GetRmax() const;
  Double_t GetTheta1() const;
  Double_t GetTheta2() const;
  Double_t GetPhi1() const;
  Double_t GetPhi2() const;
};
class TGeoParaboloid : public TGeoBBox {
 public:
  FAKE_CLASSDEF
  Double_t GetRlo() const;
  Double_t GetRhi() const;
  Double_t GetDz() const;
};
class TGeoPcon : public TGeoBBox {
 public:
  FAKE_CLASSDEF
  Double_t GetPhi1() const;
  Double_t GetDphi() const;
  Int_t GetNz() const;
  Double_t GetZ(Int_t i) const;
  Double_t GetRmin(Int_t i) const;
  Double_t GetRmax(Int_t i) const;
};
class TGeoPgon : public TGeoPcon {
 public:
  FAKE_CLASSDEF
  Int_t GetNedges() const;
};
class TGeoArb8 : public TGeoBBox {
 public:
  FAKE_CLASSDEF
  Double_t GetDz() const;
  Double_t* GetVertices();
};
class TGeoXtru : public TGeoBBox {
 public:
  FAKE_CLASSDEF
  Int_t GetNvert() const;
  Int_t GetNz() const;
  Double_t GetX(Int_t i) const;
  Double_t GetY(Int_t i) const;
  Double_t GetZ(Int_t i) const;
  Double_t GetXOffset(Int_t i) const;
  Double_t GetYOffset(Int_t i) const;
  Double_t GetScale(Int_t i) const;
};
class TGeoBoolNode : public TObject {
 public:
  enum EGeoBoolType { kGeoUnion, kGeoIntersection, kGeoSubtraction };
  virtual EGeoBoolType GetBooleanOperator() const;
  TGeoShape* GetLeftShape() const;
  TGeoShape* GetRightShape() const;
  TGeoMatrix* GetLeftMatrix() const;
  TGeoMatrix* GetRightMatrix() const;
};
class TGeoCompositeShape : public TGeoBBox {
 public:
  FAKE_CLASSDEF
  TGeoBoolNode* GetBoolNode() const;
};
class TGeoVolume;
class TGeoNode : public TNamed {
 public:
  TGeoVolume* GetVolume() const;
  TGeoMatrix* GetMatrix() const;
  Int_t GetNumber() const;
  Bool_t IsOverlapping() const;
};
class TGeoMedium;
class TGeoVolume : public TNamed {
 public:
  FAKE_CLASSDEF
  TGeoShape* GetShape() const;
  Int_t GetNdaughters() const;
  TGeoNode* GetNode(Int_t i) const;
};
class TGeoManager : public TNamed {
 public:
  TGeoVolume* GetTopVolume() const;
  TGeoNode* GetTopNode() const;
};
// ---- the reference's classes (include/*.h), public surface used by the exporter
class AOpticalComponent : public TGeoVolume {
 public:
  FAKE_CLASSDEF
};
class ALens : public AOpticalComponent { public: FAKE_CLASSDEF };
class AMirror : public AOpticalComponent { public: FAKE_CLASSDEF };
class AObscuration : public AOpticalComponent { public: FAKE_CLASSDEF };
class AFocalSurface : public AOpticalComponent { public: FAKE_CLASSDEF };
class ARefractiveIndex : public TObject { public: FAKE_CLASSDEF };
class ASellmeierFormula : public ARefractiveIndex { public: FAKE_CLASSDEF };
class ASchottFormula : public ARefractiveIndex { public: FAKE_CLASSDEF };
class ACauchyFormula : public ARefractiveIndex { public: FAKE_CLASSDEF };
class AMixedRefractiveIndex : public ARefractiveIndex { public: FAKE_CLASSDEF };
class AMultilayer : public TObject { public: FAKE_CLASSDEF };
class ABorderSurfaceCondition : public TObject {
 public:
  const AOpticalComponent* GetComponent1() const;
  const AOpticalComponent* GetComponent2() const;
  Double_t GetGaussianRoughness() const;
  std::shared_ptr<AMultilayer> GetMultilayer() const;
  Bool_t IsLambertian() const;
};
class AGeoAsphericDisk : public TGeoBBox {
 public:
  FAKE_CLASSDEF
  Double_t GetCurve1() const;
  Double_t GetCurve2() const;
  Double_t* GetK1() const;
  Double_t* GetK2() const;
  Double_t GetNPol1() const;
  Double_t GetNPol2() const;
  Double_t GetRmax() const;
  Double_t GetRmin() const;
  Double_t GetZ1() const;
  Double_t GetZ2() const;
};
class AGeoWinstonCone2D : public TGeoBBox { public: FAKE_CLASSDEF };
class AGeoWinstonConePoly : public AGeoWinstonCone2D { public: FAKE_CLASSDEF };
class ARay : public TObject {
 public:
  void GetLastPoint(Double_t* x) const;
  void GetDirection(Double_t* d) const;
  Double_t GetLambda() const;
  void AddPoint(Double_t x, Double_t y, Double_t z, Double_t t);  // TGeoTrack (ARay's base, include/ARay.h:24)
  Int_t GetNpoints() const;                                       // TGeoTrack::GetNpoints, used at src/AOpticsManager.cxx:515
  void AddNode(TGeoNode* node);
  void SetDirection(Double_t dx, Double_t dy, Double_t dz);
  void Stop();
  void Exit();
  void Focus();
  void Suspend();
  void Absorb();
};
class ARayArray : public TObject {
 public:
  TObjArray* GetRunning() const;
  void Add(ARay* ray);
};
class AOpticsManager : public TGeoManager {
 public:
  FAKE_CLASSDEF
};
#endif
