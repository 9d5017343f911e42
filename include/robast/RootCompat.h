// RootCompat.h — the small subset of CERN ROOT that ROBAST scripts on the TraceNonSequential
// path actually touch (census: SURVEY.md §7 item 4), re-implemented without ROOT so that the
// reference's macros compile against this repo.  ROOT itself is not vendored by the reference
// and is absent from this environment; behaviour documented in SURVEY.md Appendix B.
// Only geometry *description* lives here.  All ray/shape arithmetic of the traced path runs in
// the CUDA library behind include/robast_b200.h.
#ifndef ROBAST_ROOTCOMPAT_H
#define ROBAST_ROOTCOMPAT_H

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <ctime>
#include <complex>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

typedef double Double_t;
typedef float Float_t;
typedef int Int_t;
typedef unsigned int UInt_t;
typedef long Long_t;
typedef long long Long64_t;
typedef unsigned long long ULong64_t;
typedef bool Bool_t;
typedef char Char_t;
typedef const char Option_t;
const Bool_t kTRUE = true;
const Bool_t kFALSE = false;
#ifndef ClassDef
#define ClassDef(name, id)
#define ClassImp(name)
#endif
#define ROOT_VERSION(a, b, c) (((a) << 16) + ((b) << 8) + (c))
#define ROOT_VERSION_CODE ROOT_VERSION(6, 30, 0)

inline const char* Form(const char* fmt, ...) {
  static thread_local char buf[8][2048];
  static thread_local int idx = 0;
  idx = (idx + 1) & 7;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf[idx], sizeof(buf[idx]), fmt, ap);
  va_end(ap);
  return buf[idx];
}
inline void Error(const char* where, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "Error in <%s>: ", where);
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}
inline void Warning(const char* where, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "Warning in <%s>: ", where);
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}

namespace TMath {
inline constexpr Double_t Pi() { return 3.14159265358979323846; }
inline constexpr Double_t TwoPi() { return 2.0 * Pi(); }
inline constexpr Double_t PiOver2() { return Pi() / 2.0; }
inline constexpr Double_t PiOver4() { return Pi() / 4.0; }
inline constexpr Double_t DegToRad() { return Pi() / 180.0; }
inline constexpr Double_t RadToDeg() { return 180.0 / Pi(); }
inline constexpr Double_t Sqrt2() { return 1.4142135623730950488016887242097; }
inline constexpr Double_t C() { return 2.99792458e8; }  // m/s
inline Double_t Infinity() { return std::numeric_limits<Double_t>::infinity(); }
inline Double_t Sqrt(Double_t x) { return std::sqrt(x); }
inline Double_t Sin(Double_t x) { return std::sin(x); }
inline Double_t Cos(Double_t x) { return std::cos(x); }
inline Double_t Tan(Double_t x) { return std::tan(x); }
inline Double_t ASin(Double_t x) { return x < -1. ? -Pi() / 2 : (x > 1. ? Pi() / 2 : std::asin(x)); }
inline Double_t ACos(Double_t x) { return x < -1. ? Pi() : (x > 1. ? 0 : std::acos(x)); }
inline Double_t ATan(Double_t x) { return std::atan(x); }
inline Double_t ATan2(Double_t y, Double_t x) {
  if (x != 0) return std::atan2(y, x);
  if (y == 0) return 0;
  return y > 0 ? Pi() / 2 : -Pi() / 2;
}
inline Double_t Exp(Double_t x) { return std::exp(x); }
inline Double_t Log(Double_t x) { return std::log(x); }
inline Double_t Log10(Double_t x) { return std::log10(x); }
inline Double_t Power(Double_t x, Double_t y) { return std::pow(x, y); }
inline Double_t Power(Double_t x, Int_t y) { return std::pow(x, y); }
inline Double_t Floor(Double_t x) { return std::floor(x); }
inline Double_t Ceil(Double_t x) { return std::ceil(x); }
inline Double_t Hypot(Double_t x, Double_t y) { return std::hypot(x, y); }
template <class T> inline T Abs(T x) { return x < 0 ? -x : x; }
template <class T> inline T Min(T a, T b) { return a <= b ? a : b; }
template <class T> inline T Max(T a, T b) { return a >= b ? a : b; }
inline Double_t Min(Double_t a, Int_t b) { return a <= b ? a : b; }
inline Double_t Max(Double_t a, Int_t b) { return a >= b ? a : b; }
template <class T> inline T Sign(T a, T b) { return b >= 0 ? Abs(a) : -Abs(a); }
template <class T> inline Long64_t LocMin(Long64_t n, const T* a) {
  Long64_t loc = 0;
  for (Long64_t i = 1; i < n; i++)
    if (a[i] < a[loc]) loc = i;
  return loc;
}
template <class T> inline Long64_t LocMax(Long64_t n, const T* a) {
  Long64_t loc = 0;
  for (Long64_t i = 1; i < n; i++)
    if (a[i] > a[loc]) loc = i;
  return loc;
}
}  // namespace TMath

// ---------------------------------------------------------------------------- TObject & co
// Geometry epoch: every mutator of a shape, matrix, volume, table or optical property bumps it, so that AOpticsManager can keep
// its flattened scene (export + device tables) across TraceNonSequential calls for as long as nothing was touched.
inline std::atomic<unsigned long long>& RbGeomEpoch() {
  static std::atomic<unsigned long long> e{1};
  return e;
}
// a constructor that fills its members through its own setters is not a change of the geometry: the object is new, it gets
// into a scene only through AddNode / RegisterYourself / a Set... call on something else, and those bump the epoch
struct RbGeomQuiet {
  static int& Depth() { thread_local int d = 0; return d; }
  RbGeomQuiet() { ++Depth(); }
  ~RbGeomQuiet() { --Depth(); }
};
inline void RbGeomTouch() {
  if (RbGeomQuiet::Depth() == 0) RbGeomEpoch().fetch_add(1, std::memory_order_relaxed);
}

class TObject {
 public:
  virtual ~TObject() {}
  virtual const char* GetName() const { return ""; }
  virtual void Draw(Option_t* = "") {}
};

class TNamed : public TObject {
 protected:
  std::string fName, fTitle;

 public:
  TNamed() {}
  TNamed(const char* name, const char* title = "") : fName(name ? name : ""), fTitle(title ? title : "") {}
  const char* GetName() const override { return fName.c_str(); }
  const char* GetTitle() const { return fTitle.c_str(); }
  virtual void SetName(const char* n) { fName = n ? n : ""; }
  virtual void SetTitle(const char* t) { fTitle = t ? t : ""; }
};

class TObjArray : public TObject {
  std::vector<TObject*> fA;
  Bool_t fOwner = kFALSE;

 public:
  TObjArray() {}
  ~TObjArray() override {
    if (fOwner) Delete();
  }
  void SetOwner(Bool_t o = kTRUE) { fOwner = o; }
  void Add(TObject* o) { fA.push_back(o); }
  TObject* At(Int_t i) const { return (i >= 0 && i < (Int_t)fA.size()) ? fA[i] : nullptr; }
  TObject* operator[](Int_t i) const { return At(i); }
  TObject* UncheckedAt(Int_t i) const { return fA[i]; }
  Int_t GetLast() const {
    for (Int_t i = (Int_t)fA.size() - 1; i >= 0; i--)
      if (fA[i]) return i;
    return -1;
  }
  Int_t GetEntries() const {
    Int_t n = 0;
    for (auto* p : fA)
      if (p) n++;
    return n;
  }
  Int_t GetEntriesFast() const { return GetLast() + 1; }
  TObject* RemoveAt(Int_t i) {
    TObject* o = At(i);
    if (o) fA[i] = nullptr;
    return o;
  }
  void Expand(Int_t n) { fA.resize(n); }
  void Clear() { fA.clear(); }
  void Delete() {
    for (auto* p : fA) delete p;
    fA.clear();
  }
  void Reserve(size_t n) { fA.reserve(n); }
};

// ---------------------------------------------------------------------------- TVector3
class TVector3 {
  Double_t fX, fY, fZ;

 public:
  TVector3(Double_t x = 0, Double_t y = 0, Double_t z = 0) : fX(x), fY(y), fZ(z) {}
  TVector3(const Double_t* a) : fX(a[0]), fY(a[1]), fZ(a[2]) {}
  Double_t X() const { return fX; }
  Double_t Y() const { return fY; }
  Double_t Z() const { return fZ; }
  Double_t x() const { return fX; }
  Double_t y() const { return fY; }
  Double_t z() const { return fZ; }
  Double_t operator[](int i) const { return i == 0 ? fX : (i == 1 ? fY : fZ); }
  Double_t& operator[](int i) { return i == 0 ? fX : (i == 1 ? fY : fZ); }
  void SetXYZ(Double_t x, Double_t y, Double_t z) { fX = x; fY = y; fZ = z; }
  void SetX(Double_t v) { fX = v; }
  void SetY(Double_t v) { fY = v; }
  void SetZ(Double_t v) { fZ = v; }
  void GetXYZ(Double_t* c) const { c[0] = fX; c[1] = fY; c[2] = fZ; }
  Double_t Mag2() const { return fX * fX + fY * fY + fZ * fZ; }
  Double_t Mag() const { return std::sqrt(Mag2()); }
  Double_t Perp() const { return std::sqrt(fX * fX + fY * fY); }
  Double_t Theta() const { return fX == 0 && fY == 0 && fZ == 0 ? 0 : std::atan2(Perp(), fZ); }
  Double_t Phi() const { return fX == 0 && fY == 0 ? 0 : std::atan2(fY, fX); }
  void SetMagThetaPhi(Double_t mag, Double_t theta, Double_t phi) {
    Double_t amag = std::fabs(mag);
    fX = amag * std::sin(theta) * std::cos(phi);
    fY = amag * std::sin(theta) * std::sin(phi);
    fZ = amag * std::cos(theta);
  }
  void SetMag(Double_t ma) {
    Double_t factor = Mag();
    if (factor == 0) return;
    factor = ma / factor;
    fX *= factor; fY *= factor; fZ *= factor;
  }
  TVector3 Unit() const {
    Double_t tot2 = Mag2();
    Double_t tot = (tot2 > 0) ? 1.0 / std::sqrt(tot2) : 1.0;
    return TVector3(fX * tot, fY * tot, fZ * tot);
  }
  Double_t Dot(const TVector3& p) const { return fX * p.fX + fY * p.fY + fZ * p.fZ; }
  TVector3 Cross(const TVector3& p) const {
    return TVector3(fY * p.fZ - p.fY * fZ, fZ * p.fX - p.fZ * fX, fX * p.fY - p.fX * fY);
  }
  Double_t Angle(const TVector3& q) const {
    Double_t ptot2 = Mag2() * q.Mag2();
    if (ptot2 <= 0) return 0;
    Double_t arg = Dot(q) / std::sqrt(ptot2);
    if (arg > 1.0) arg = 1.0;
    if (arg < -1.0) arg = -1.0;
    return std::acos(arg);
  }
  void RotateZ(Double_t angle) {
    Double_t s = std::sin(angle), c = std::cos(angle), xx = fX;
    fX = c * xx - s * fY;
    fY = s * xx + c * fY;
  }
  void RotateX(Double_t angle) {
    Double_t s = std::sin(angle), c = std::cos(angle), yy = fY;
    fY = c * yy - s * fZ;
    fZ = s * yy + c * fZ;
  }
  void RotateY(Double_t angle) {
    Double_t s = std::sin(angle), c = std::cos(angle), zz = fZ;
    fZ = c * zz - s * fX;
    fX = s * zz + c * fX;
  }
  void RotateUz(const TVector3& NewUzVector) {
    // NewUzVector must be normalized
    Double_t u1 = NewUzVector.fX, u2 = NewUzVector.fY, u3 = NewUzVector.fZ;
    Double_t up = u1 * u1 + u2 * u2;
    if (up) {
      up = std::sqrt(up);
      Double_t px = fX, py = fY, pz = fZ;
      fX = (u1 * u3 * px - u2 * py + u1 * up * pz) / up;
      fY = (u2 * u3 * px + u1 * py + u2 * up * pz) / up;
      fZ = (u3 * u3 * px - px + u3 * up * pz) / up;
    } else if (u3 < 0.) {
      fX = -fX;
      fZ = -fZ;
    }
  }
  TVector3& operator*=(Double_t a) { fX *= a; fY *= a; fZ *= a; return *this; }
  TVector3& operator+=(const TVector3& p) { fX += p.fX; fY += p.fY; fZ += p.fZ; return *this; }
  TVector3& operator-=(const TVector3& p) { fX -= p.fX; fY -= p.fY; fZ -= p.fZ; return *this; }
  TVector3 operator-() const { return TVector3(-fX, -fY, -fZ); }
};
inline TVector3 operator+(const TVector3& a, const TVector3& b) { return TVector3(a.X() + b.X(), a.Y() + b.Y(), a.Z() + b.Z()); }
inline TVector3 operator-(const TVector3& a, const TVector3& b) { return TVector3(a.X() - b.X(), a.Y() - b.Y(), a.Z() - b.Z()); }
inline Double_t operator*(const TVector3& a, const TVector3& b) { return a.Dot(b); }
inline TVector3 operator*(const TVector3& p, Double_t a) { return TVector3(a * p.X(), a * p.Y(), a * p.Z()); }
inline TVector3 operator*(Double_t a, const TVector3& p) { return TVector3(a * p.X(), a * p.Y(), a * p.Z()); }

// ---------------------------------------------------------------------------- TRandom (MT19937 = TRandom3)
class TRandom : public TObject {
  std::mt19937 fGen;

 public:
  TRandom(UInt_t seed = 4357) : fGen(seed) {}
  void SetSeed(UInt_t seed = 0) { fGen.seed(seed ? seed : (UInt_t)std::random_device{}()); }
  Double_t Rndm() {
    UInt_t y;
    do { y = (UInt_t)fGen(); } while (!y);
    return y * 2.3283064365386963e-10;  // (0,1)
  }
  Double_t Uniform(Double_t x1 = 1) { return x1 * Rndm(); }
  Double_t Uniform(Double_t x1, Double_t x2) { return x1 + (x2 - x1) * Rndm(); }
  Double_t Exp(Double_t tau) { return -tau * std::log(Rndm()); }
  Double_t Gaus(Double_t mean = 0, Double_t sigma = 1) {
    Double_t u1 = Rndm(), u2 = Rndm();
    return mean + sigma * std::sqrt(-2 * std::log(u1)) * std::cos(TMath::TwoPi() * u2);
  }
  void Sphere(Double_t& x, Double_t& y, Double_t& z, Double_t r) {
    Double_t a = 0, b = 0, r2 = 1;
    while (r2 > 0.25) {
      a = Rndm() - 0.5;
      b = Rndm() - 0.5;
      r2 = a * a + b * b;
    }
    z = r * (-1. + 8.0 * r2);
    Double_t scale = 8.0 * r * std::sqrt(0.25 - r2);
    x = a * scale;
    y = b * scale;
  }
};
typedef TRandom TRandom3;
inline TRandom*& gRandomRef() {
  static TRandom* g = new TRandom();
  return g;
}
#define gRandom (gRandomRef())

// ---------------------------------------------------------------------------- TGeoMatrix family
class TGeoMatrix;
class TGeoShape;
struct RobastRegistry {  // name lookups used by TGeoCompositeShape expressions (gGeoManager lists in ROOT)
  std::map<std::string, TGeoMatrix*> matrices;
  std::map<std::string, TGeoShape*> shapes;
  static RobastRegistry& Get() {
    static RobastRegistry r;
    return r;
  }
};

class TGeoMatrix : public TNamed {
 protected:
  Double_t fRot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  Double_t fTr[3] = {0, 0, 0};

 public:
  TGeoMatrix() {}
  TGeoMatrix(const char* name) : TNamed(name) {}
  const Double_t* GetRotationMatrix() const { return fRot; }
  const Double_t* GetTranslation() const { return fTr; }
  // ROOT keeps every registered matrix in gGeoManager's list, where composite-shape expressions find them by name
  virtual void RegisterYourself() { RbGeomTouch();
    if (!fName.empty()) RobastRegistry::Get().matrices[fName] = this;
  }
  virtual ~TGeoMatrix() {}
  Bool_t IsIdentity() const {
    static const Double_t id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    return !memcmp(fRot, id, sizeof(id)) && fTr[0] == 0 && fTr[1] == 0 && fTr[2] == 0;
  }
  void LocalToMaster(const Double_t* l, Double_t* m) const {
    Double_t r[3];
    for (int i = 0; i < 3; i++) r[i] = fTr[i] + l[0] * fRot[3 * i] + l[1] * fRot[3 * i + 1] + l[2] * fRot[3 * i + 2];
    m[0] = r[0]; m[1] = r[1]; m[2] = r[2];
  }
  void LocalToMasterVect(const Double_t* l, Double_t* m) const {
    Double_t r[3];
    for (int i = 0; i < 3; i++) r[i] = l[0] * fRot[3 * i] + l[1] * fRot[3 * i + 1] + l[2] * fRot[3 * i + 2];
    m[0] = r[0]; m[1] = r[1]; m[2] = r[2];
  }
  void MasterToLocal(const Double_t* m, Double_t* l) const {
    Double_t mt0 = m[0] - fTr[0], mt1 = m[1] - fTr[1], mt2 = m[2] - fTr[2];
    Double_t r[3];
    for (int i = 0; i < 3; i++) r[i] = mt0 * fRot[i] + mt1 * fRot[i + 3] + mt2 * fRot[i + 6];
    l[0] = r[0]; l[1] = r[1]; l[2] = r[2];
  }
  void MasterToLocalVect(const Double_t* m, Double_t* l) const {
    Double_t r[3];
    for (int i = 0; i < 3; i++) r[i] = m[0] * fRot[i] + m[1] * fRot[i + 3] + m[2] * fRot[i + 6];
    l[0] = r[0]; l[1] = r[1]; l[2] = r[2];
  }
  // this = this * right  (apply `right` first, then the old `this`)
  void MultiplyRight(const TGeoMatrix& b) { RbGeomTouch();
    Double_t r[9], t[3];
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) r[3 * i + j] = fRot[3 * i] * b.fRot[j] + fRot[3 * i + 1] * b.fRot[3 + j] + fRot[3 * i + 2] * b.fRot[6 + j];
      t[i] = fTr[i] + fRot[3 * i] * b.fTr[0] + fRot[3 * i + 1] * b.fTr[1] + fRot[3 * i + 2] * b.fTr[2];
    }
    memcpy(fRot, r, sizeof(r));
    memcpy(fTr, t, sizeof(t));
  }
  void CopyFrom(const TGeoMatrix& o) {
    memcpy(fRot, o.fRot, sizeof(fRot));
    memcpy(fTr, o.fTr, sizeof(fTr));
  }
  void SetRotationArray(const Double_t* r) { RbGeomTouch(); memcpy(fRot, r, sizeof(fRot)); }
  void SetTranslationArray(const Double_t* t) { RbGeomTouch(); memcpy(fTr, t, sizeof(fTr)); }
};

class TGeoTranslation : public TGeoMatrix {
 public:
  TGeoTranslation() {}
  TGeoTranslation(Double_t dx, Double_t dy, Double_t dz) { RbGeomQuiet q; SetTranslation(dx, dy, dz); }
  TGeoTranslation(const char* name, Double_t dx, Double_t dy, Double_t dz) : TGeoMatrix(name) { RbGeomQuiet q; SetTranslation(dx, dy, dz); }
  void SetTranslation(Double_t dx, Double_t dy, Double_t dz) { RbGeomTouch(); fTr[0] = dx; fTr[1] = dy; fTr[2] = dz; }
};

class TGeoRotation : public TGeoMatrix {
 public:
  TGeoRotation() {}
  TGeoRotation(const char* name) : TGeoMatrix(name) {}
  TGeoRotation(const char* name, Double_t phi, Double_t theta, Double_t psi) : TGeoMatrix(name) { RbGeomQuiet q; SetAngles(phi, theta, psi); }
  // Euler Z-X-Z in degrees: R = Rz(phi) Rx(theta) Rz(psi)   
  void SetAngles(Double_t phi, Double_t theta, Double_t psi) { RbGeomTouch();
    const Double_t degrad = TMath::Pi() / 180.;
    Double_t sinphi = std::sin(degrad * phi), cosphi = std::cos(degrad * phi);
    Double_t sinthe = std::sin(degrad * theta), costhe = std::cos(degrad * theta);
    Double_t sinpsi = std::sin(degrad * psi), cospsi = std::cos(degrad * psi);
    fRot[0] = cospsi * cosphi - costhe * sinphi * sinpsi;
    fRot[1] = -sinpsi * cosphi - costhe * sinphi * cospsi;
    fRot[2] = sinthe * sinphi;
    fRot[3] = cospsi * sinphi + costhe * cosphi * sinpsi;
    fRot[4] = -sinpsi * sinphi + costhe * cosphi * cospsi;
    fRot[5] = -sinthe * cosphi;
    fRot[6] = sinpsi * sinthe;
    fRot[7] = cospsi * sinthe;
    fRot[8] = costhe;
  }
  void RotateX(Double_t angle) { RbGeomTouch(); TGeoRotation r("", 0, angle, 0); MultiplyBy(&r, kFALSE); }
  void RotateZ(Double_t angle) { RbGeomTouch(); TGeoRotation r("", angle, 0, 0); MultiplyBy(&r, kFALSE); }
  void RotateY(Double_t angle) { RbGeomTouch();
    Double_t a = angle * TMath::DegToRad(), c = std::cos(a), s = std::sin(a);
    TGeoRotation r;
    Double_t m[9] = {c, 0, s, 0, 1, 0, -s, 0, c};
    r.SetRotationArray(m);
    MultiplyBy(&r, kFALSE);
  }
  // after=true: this = this*rot ; after=false: this = rot*this
  void MultiplyBy(const TGeoRotation* rot, Bool_t after = kTRUE) { RbGeomTouch();
    const Double_t *a = after ? fRot : rot->fRot, *b = after ? rot->fRot : fRot;
    Double_t r[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
    memcpy(fRot, r, sizeof(r));
  }
};

class TGeoCombiTrans : public TGeoMatrix {
 public:
  TGeoCombiTrans() {}
  TGeoCombiTrans(const char* name) : TGeoMatrix(name) {}
  TGeoCombiTrans(const TGeoTranslation& tr, const TGeoRotation& rot) {
    memcpy(fTr, tr.GetTranslation(), sizeof(fTr));
    memcpy(fRot, rot.GetRotationMatrix(), sizeof(fRot));
  }
  TGeoCombiTrans(Double_t dx, Double_t dy, Double_t dz, TGeoRotation* rot) : fRotation(rot) {
    fTr[0] = dx; fTr[1] = dy; fTr[2] = dz;
    if (rot) memcpy(fRot, rot->GetRotationMatrix(), sizeof(fRot));
  }
  TGeoCombiTrans(const char* name, Double_t dx, Double_t dy, Double_t dz, TGeoRotation* rot) : TGeoMatrix(name), fRotation(rot) {
    fTr[0] = dx; fTr[1] = dy; fTr[2] = dz;
    if (rot) memcpy(fRot, rot->GetRotationMatrix(), sizeof(fRot));
  }
  // TGeoCombiTrans::RegisterYourself also registers the rotation it was built from (tutorials/AshraOptics.C:865-866 relies on
  // it: "30_rot3" is only ever handed to a TGeoCombiTrans that AddNode registers, and is then named in a composite expression)
  void RegisterYourself() override { RbGeomTouch();
    TGeoMatrix::RegisterYourself();
    if (fRotation) fRotation->RegisterYourself();
  }

 private:
  TGeoRotation* fRotation = nullptr;
};

class TGeoHMatrix : public TGeoMatrix {
 public:
  TGeoHMatrix() {}
  TGeoHMatrix(const char* name) : TGeoMatrix(name) {}
  TGeoHMatrix(const TGeoMatrix& m) : TGeoMatrix(m.GetName()) { CopyFrom(m); }
  TGeoHMatrix operator*(const TGeoMatrix& right) const {
    TGeoHMatrix h(*this);
    h.MultiplyRight(right);
    return h;
  }
  TGeoHMatrix& operator*=(const TGeoMatrix& right) { MultiplyRight(right); return *this; }
  void Multiply(const TGeoMatrix* right) { RbGeomTouch(); MultiplyRight(*right); }
};
inline TGeoHMatrix operator*(const TGeoMatrix& a, const TGeoMatrix& b) {
  TGeoHMatrix h(a);
  h.MultiplyRight(b);
  return h;
}

// ---------------------------------------------------------------------------- shapes (description only)
class TGeoShape : public TNamed {
 public:
  enum EKind { kBBox, kTube, kSphere, kParaboloid, kPgon, kPcon, kAsphere, kWinston2D, kWinstonPoly, kComposite, kArb8, kXtru };
  TGeoShape() {}
  TGeoShape(const char* name) : TNamed(name) {
    if (name && *name) RobastRegistry::Get().shapes[fName] = this;
  }
  void SetName(const char* n) override {
    TNamed::SetName(n);
    if (n && *n) RobastRegistry::Get().shapes[fName] = this;
  }
  virtual EKind Kind() const = 0;
  static Double_t Big() { return 1.E30; }
  static Double_t Tolerance() { return 1.E-10; }
};

class TGeoBBox : public TGeoShape {
 protected:
  Double_t fDX = 0, fDY = 0, fDZ = 0, fOrigin[3] = {0, 0, 0};

 public:
  TGeoBBox() {}
  TGeoBBox(Double_t dx, Double_t dy, Double_t dz, Double_t* origin = nullptr) { SetBoxDimensions(dx, dy, dz, origin); }
  TGeoBBox(const char* name, Double_t dx, Double_t dy, Double_t dz, Double_t* origin = nullptr) : TGeoShape(name) {
    SetBoxDimensions(dx, dy, dz, origin);
  }
  void SetBoxDimensions(Double_t dx, Double_t dy, Double_t dz, Double_t* origin = nullptr) { RbGeomTouch();
    fDX = dx; fDY = dy; fDZ = dz;
    if (origin) memcpy(fOrigin, origin, sizeof(fOrigin));
  }
  EKind Kind() const override { return kBBox; }
  virtual Double_t GetDX() const { return fDX; }
  virtual Double_t GetDY() const { return fDY; }
  virtual Double_t GetDZ() const { return fDZ; }
  virtual const Double_t* GetOrigin() const { return fOrigin; }
};

class TGeoTube : public TGeoBBox {
 protected:
  Double_t fRmin = 0, fRmax = 0, fDz = 0;

 public:
  TGeoTube() {}
  TGeoTube(Double_t rmin, Double_t rmax, Double_t dz) { Set(rmin, rmax, dz); }
  TGeoTube(const char* name, Double_t rmin, Double_t rmax, Double_t dz) { SetName(name); Set(rmin, rmax, dz); }
  void Set(Double_t rmin, Double_t rmax, Double_t dz) { RbGeomTouch();
    fRmin = rmin; fRmax = rmax; fDz = dz;
    fDX = fDY = rmax; fDZ = dz;
  }
  EKind Kind() const override { return kTube; }
  Double_t GetRmin() const { return fRmin; }
  Double_t GetRmax() const { return fRmax; }
  Double_t GetDz() const { return fDz; }
};

class TGeoSphere : public TGeoBBox {
  Double_t fRmin, fRmax, fTheta1, fTheta2, fPhi1, fPhi2;

 public:
  TGeoSphere(Double_t rmin, Double_t rmax, Double_t theta1 = 0, Double_t theta2 = 180, Double_t phi1 = 0, Double_t phi2 = 360) {
    Set(rmin, rmax, theta1, theta2, phi1, phi2);
  }
  TGeoSphere(const char* name, Double_t rmin, Double_t rmax, Double_t theta1 = 0, Double_t theta2 = 180, Double_t phi1 = 0, Double_t phi2 = 360) {
    SetName(name);
    Set(rmin, rmax, theta1, theta2, phi1, phi2);
  }
  void Set(Double_t rmin, Double_t rmax, Double_t t1, Double_t t2, Double_t p1, Double_t p2) { RbGeomTouch();
    fRmin = rmin; fRmax = rmax; fTheta1 = t1; fTheta2 = t2; fPhi1 = p1; fPhi2 = p2;
    if (fPhi1 < 0) fPhi1 += 360.;
    while (fPhi2 <= fPhi1) fPhi2 += 360.;
    fDX = fDY = fDZ = rmax;
  }
  EKind Kind() const override { return kSphere; }
  Double_t GetRmin() const { return fRmin; }
  Double_t GetRmax() const { return fRmax; }
  Double_t GetTheta1() const { return fTheta1; }
  Double_t GetTheta2() const { return fTheta2; }
  Double_t GetPhi1() const { return fPhi1; }
  Double_t GetPhi2() const { return fPhi2; }
};

class TGeoParaboloid : public TGeoBBox {
  Double_t fRlo, fRhi, fDz;

 public:
  TGeoParaboloid(Double_t rlo, Double_t rhi, Double_t dz) : fRlo(rlo), fRhi(rhi), fDz(dz) { Box(); }
  TGeoParaboloid(const char* name, Double_t rlo, Double_t rhi, Double_t dz) : fRlo(rlo), fRhi(rhi), fDz(dz) {
    SetName(name);
    Box();
  }
  void Box() { fDX = fDY = std::max(fRlo, fRhi); fDZ = fDz; }
  EKind Kind() const override { return kParaboloid; }
  Double_t GetRlo() const { return fRlo; }
  Double_t GetRhi() const { return fRhi; }
  Double_t GetDz() const { return fDz; }
};

class TGeoPcon : public TGeoBBox {
 protected:
  Double_t fPhi1, fDphi;
  Int_t fNz;
  std::vector<Double_t> fZ, fRmin, fRmax;

 public:
  TGeoPcon(Double_t phi, Double_t dphi, Int_t nz) : fPhi1(phi), fDphi(dphi), fNz(nz), fZ(nz), fRmin(nz), fRmax(nz) {}
  TGeoPcon(const char* name, Double_t phi, Double_t dphi, Int_t nz) : fPhi1(phi), fDphi(dphi), fNz(nz), fZ(nz), fRmin(nz), fRmax(nz) { SetName(name); }
  virtual void DefineSection(Int_t snum, Double_t z, Double_t rmin, Double_t rmax) { RbGeomTouch();
    if (snum < 0 || snum >= fNz) return;
    fZ[snum] = z; fRmin[snum] = rmin; fRmax[snum] = rmax;
    if (snum == fNz - 1) {
      if (fZ[0] > fZ[snum]) {  // ROOT reorders descending definitions
        std::reverse(fZ.begin(), fZ.end());
        std::reverse(fRmin.begin(), fRmin.end());
        std::reverse(fRmax.begin(), fRmax.end());
      }
      Double_t rmx = 0;
      for (auto r : fRmax) rmx = std::max(rmx, r);
      fDX = fDY = rmx;  // loose (ignores phi range / polygon corners)
      fDZ = 0.5 * (fZ[fNz - 1] - fZ[0]);
      fOrigin[2] = 0.5 * (fZ[fNz - 1] + fZ[0]);
    }
  }
  EKind Kind() const override { return kPcon; }
  Double_t GetPhi1() const { return fPhi1; }
  Double_t GetDphi() const { return fDphi; }
  Int_t GetNz() const { return fNz; }
  Double_t GetZ(Int_t i) const { return fZ[i]; }
  Double_t GetRmin(Int_t i) const { return fRmin[i]; }
  Double_t GetRmax(Int_t i) const { return fRmax[i]; }
};

class TGeoPgon : public TGeoPcon {
 protected:
  Int_t fNedges;

 public:
  TGeoPgon(Double_t phi, Double_t dphi, Int_t nedges, Int_t nz) : TGeoPcon(phi, dphi, nz), fNedges(nedges) {}
  TGeoPgon(const char* name, Double_t phi, Double_t dphi, Int_t nedges, Int_t nz) : TGeoPcon(name, phi, dphi, nz), fNedges(nedges) {}
  EKind Kind() const override { return kPgon; }
  Int_t GetNedges() const { return fNedges; }
};

// TGeoArb8: 8 vertices on two z planes (0-3 at -dz, 4-7 at +dz), possibly twisted (tutorials/AshraOptics.C:264-284)
class TGeoArb8 : public TGeoBBox {
 protected:
  Double_t fDz = 0, fXY[8][2] = {};
  void Box() {
    Double_t xm = 0, ym = 0;
    for (int i = 0; i < 8; i++) { xm = std::max(xm, std::fabs(fXY[i][0])); ym = std::max(ym, std::fabs(fXY[i][1])); }
    fDX = xm; fDY = ym; fDZ = fDz;  // loose (centred); the scene builder computes the exact box
  }

 public:
  TGeoArb8() {}
  TGeoArb8(Double_t dz, Double_t* vertices = nullptr) : fDz(dz) { Init(vertices); }
  TGeoArb8(const char* name, Double_t dz, Double_t* vertices = nullptr) : fDz(dz) { SetName(name); Init(vertices); }
  void Init(const Double_t* vertices) {
    if (vertices)
      for (int i = 0; i < 8; i++) { fXY[i][0] = vertices[2 * i]; fXY[i][1] = vertices[2 * i + 1]; }
    Box();
  }
  void SetVertex(Int_t vnum, Double_t x, Double_t y) { RbGeomTouch();
    if (vnum < 0 || vnum > 7) return;
    fXY[vnum][0] = x; fXY[vnum][1] = y;
    Box();
  }
  void SetDz(Double_t dz) { RbGeomTouch(); fDz = dz; Box(); }
  EKind Kind() const override { return kArb8; }
  Double_t GetDz() const { return fDz; }
  Double_t* GetVertices() { return &fXY[0][0]; }
  const Double_t* GetVertices() const { return &fXY[0][0]; }
};

// TGeoXtru: polygon extruded along z with a placement (x0,y0) and scale per section (tutorials/AshraOptics.C:791-1021)
class TGeoXtru : public TGeoBBox {
 protected:
  Int_t fNz = 0;
  std::vector<Double_t> fX, fY, fZ, fX0, fY0, fScale;
  void Box() {
    Double_t xm = 0, ym = 0;
    for (Int_t k = 0; k < fNz; k++)
      for (size_t i = 0; i < fX.size(); i++) {
        xm = std::max(xm, std::fabs(fX0[k] + fScale[k] * fX[i]));
        ym = std::max(ym, std::fabs(fY0[k] + fScale[k] * fY[i]));
      }
    fDX = xm; fDY = ym;
    if (fNz > 0) { fDZ = 0.5 * (fZ[fNz - 1] - fZ[0]); fOrigin[2] = 0.5 * (fZ[fNz - 1] + fZ[0]); }
  }

 public:
  TGeoXtru(Int_t nz) : fNz(nz), fZ(nz, 0.), fX0(nz, 0.), fY0(nz, 0.), fScale(nz, 1.) {}
  Bool_t DefinePolygon(Int_t nvert, const Double_t* xv, const Double_t* yv) { RbGeomTouch();
    if (nvert < 3) return kFALSE;
    fX.assign(xv, xv + nvert);
    fY.assign(yv, yv + nvert);
    Box();
    return kTRUE;
  }
  virtual void DefineSection(Int_t snum, Double_t z, Double_t x0 = 0., Double_t y0 = 0., Double_t scale = 1.) { RbGeomTouch();
    if (snum < 0 || snum >= fNz) return;
    fZ[snum] = z; fX0[snum] = x0; fY0[snum] = y0; fScale[snum] = scale;
    Box();
  }
  EKind Kind() const override { return kXtru; }
  Int_t GetNz() const { return fNz; }
  Int_t GetNvert() const { return (Int_t)fX.size(); }
  Double_t GetX(Int_t i) const { return fX[i]; }
  Double_t GetY(Int_t i) const { return fY[i]; }
  Double_t GetZ(Int_t i) const { return fZ[i]; }
  Double_t GetXOffset(Int_t i) const { return fX0[i]; }
  Double_t GetYOffset(Int_t i) const { return fY0[i]; }
  Double_t GetScale(Int_t i) const { return fScale[i]; }
};

// Boolean expression tree of a TGeoCompositeShape
struct TGeoBoolNode {
  enum EOp { kUnion, kIntersection, kSubtraction };
  EOp op;
  TGeoShape* left = nullptr;
  TGeoShape* right = nullptr;
  TGeoMatrix* lmat = nullptr;
  TGeoMatrix* rmat = nullptr;
};

class TGeoCompositeShape : public TGeoBBox {
  TGeoBoolNode* fNode = nullptr;

  static std::string Strip(const std::string& s) {
    std::string r;
    for (char c : s)
      if (c != ' ' && c != '\t' && c != '\n') r += c;
    return r;
  }
  // Split at the LAST lowest-precedence operator at parenthesis level 0 ('+','-' lowest, then '*'),
  // which yields ROOT's left-associative trees ("A+B+C" -> (A+B)+C).
  static TGeoShape* Parse(const std::string& e, TGeoMatrix** mat, const char* owner) {
    *mat = nullptr;
    if (e.empty()) throw std::runtime_error(std::string("TGeoCompositeShape ") + owner + ": empty operand");
    int level = 0, pos = -1, pos_mul = -1;
    for (int i = 0; i < (int)e.size(); i++) {
      char c = e[i];
      if (c == '(') level++;
      else if (c == ')') level--;
      else if (level == 0 && i > 0) {
        if (c == '+' || c == '-') pos = i;
        else if (c == '*') pos_mul = i;
      }
    }
    if (pos < 0) pos = pos_mul;
    if (pos >= 0) {
      auto* node = new TGeoBoolNode;
      node->op = e[pos] == '+' ? TGeoBoolNode::kUnion : (e[pos] == '*' ? TGeoBoolNode::kIntersection : TGeoBoolNode::kSubtraction);
      node->left = Parse(e.substr(0, pos), &node->lmat, owner);
      node->right = Parse(e.substr(pos + 1), &node->rmat, owner);
      auto* comp = new TGeoCompositeShape();
      comp->fNode = node;
      return comp;
    }
    // a single operand: "(expr)", "(expr):mat", "name" or "name:mat"
    if (e[0] == '(') {
      int lev = 0, close = -1;
      for (int i = 0; i < (int)e.size(); i++) {
        if (e[i] == '(') lev++;
        if (e[i] == ')' && --lev == 0) { close = i; break; }
      }
      if (close < 0) throw std::runtime_error(std::string("TGeoCompositeShape ") + owner + ": unbalanced parentheses");
      TGeoMatrix* inner = nullptr;
      TGeoShape* s = Parse(e.substr(1, close - 1), &inner, owner);
      std::string rest = e.substr(close + 1);
      if (!rest.empty()) {
        if (rest[0] != ':') throw std::runtime_error(std::string("TGeoCompositeShape ") + owner + ": bad token after ')'");
        TGeoMatrix* outer = Lookup(rest.substr(1), owner);
        if (inner) {  // (name:m1):m2  -> m2*m1
          auto* h = new TGeoHMatrix(*outer);
          h->MultiplyRight(*inner);
          *mat = h;
        } else *mat = outer;
      } else *mat = inner;
      return s;
    }
    size_t colon = e.find(':');
    std::string sname = e.substr(0, colon);
    auto& reg = RobastRegistry::Get();
    auto it = reg.shapes.find(sname);
    if (it == reg.shapes.end()) throw std::runtime_error(std::string("TGeoCompositeShape ") + owner + ": shape '" + sname + "' not found");
    if (colon != std::string::npos) *mat = Lookup(e.substr(colon + 1), owner);
    return it->second;
  }
  static TGeoMatrix* Lookup(const std::string& name, const char* owner) {
    auto& reg = RobastRegistry::Get();
    auto it = reg.matrices.find(name);
    if (it == reg.matrices.end())
      throw std::runtime_error(std::string("TGeoCompositeShape ") + owner + ": matrix '" + name + "' not registered (RegisterYourself)");
    return it->second;
  }
  TGeoCompositeShape() {}

 public:
  TGeoCompositeShape(const char* name, const char* expression) {
    TGeoMatrix* m = nullptr;
    TGeoShape* s = Parse(Strip(expression), &m, name);
    auto* c = dynamic_cast<TGeoCompositeShape*>(s);
    if (!c || c->fNode == nullptr) throw std::runtime_error(std::string("TGeoCompositeShape ") + name + ": expression has no boolean operator");
    // "(A*B):m" at the top level (tutorials/AshraOptics.C:303-304): ROOT's TGeoCompositeShape::MakeNode warns and drops the matrix
    if (m) fprintf(stderr, "Warning in <TGeoCompositeShape::MakeNode>: %s: no geometrical transformation allowed at this level\n", name);
    fNode = c->fNode;
    SetName(name);
  }
  TGeoCompositeShape(const char* name, TGeoBoolNode* node) : fNode(node) { SetName(name); }
  EKind Kind() const override { return kComposite; }
  TGeoBoolNode* GetBoolNode() const { return fNode; }
};

// ---------------------------------------------------------------------------- volumes / nodes / manager
class TGeoMedium;
class TGeoVolume;
class TGeoNode : public TNamed {
  TGeoVolume* fVolume;
  TGeoMatrix* fMatrix;
  Int_t fCopy;
  Bool_t fOverlap;

 public:
  TGeoNode(TGeoVolume* vol, Int_t copy, TGeoMatrix* mat, Bool_t ovl);
  TGeoVolume* GetVolume() const { return fVolume; }
  TGeoMatrix* GetMatrix() const { return fMatrix; }
  Int_t GetNumber() const { return fCopy; }
  Bool_t IsOverlapping() const { return fOverlap; }
};

class TGeoVolume : public TNamed {
 protected:
  TGeoShape* fShape = nullptr;
  std::vector<TGeoNode*> fNodes;

 public:
  TGeoVolume() {}
  TGeoVolume(const char* name, const TGeoShape* shape, const TGeoMedium* = nullptr) : TNamed(name), fShape(const_cast<TGeoShape*>(shape)) {}
  TGeoShape* GetShape() const { return fShape; }
  virtual void AddNode(TGeoVolume* vol, Int_t copy_no, TGeoMatrix* mat = nullptr, Option_t* = "") { RbGeomTouch();
    if (mat) mat->RegisterYourself();  // TGeoVolume::AddNode registers the placement matrix
    fNodes.push_back(new TGeoNode(vol, copy_no, mat, kFALSE));
  }
  virtual void AddNodeOverlap(TGeoVolume* vol, Int_t copy_no, TGeoMatrix* mat = nullptr, Option_t* = "") { RbGeomTouch();
    if (mat) mat->RegisterYourself();
    fNodes.push_back(new TGeoNode(vol, copy_no, mat, kTRUE));
  }
  Int_t GetNdaughters() const { return (Int_t)fNodes.size(); }
  TGeoNode* GetNode(Int_t i) const { return fNodes[i]; }
  void SetLineColor(Int_t) {}
  void SetLineWidth(Int_t) {}
  void SetTransparency(Int_t) {}
  void SetVisibility(Bool_t) {}
  void Draw(Option_t* = "") override {}
};
inline TGeoNode::TGeoNode(TGeoVolume* vol, Int_t copy, TGeoMatrix* mat, Bool_t ovl)
    : TNamed(Form("%s_%d", vol->GetName(), copy)), fVolume(vol), fMatrix(mat), fCopy(copy), fOverlap(ovl) {}

class TGeoManager : public TNamed {
 protected:
  TGeoVolume* fTopVolume = nullptr;
  Int_t fMaxThreads = 1;
  Bool_t fMultiThread = kFALSE;
  Bool_t fClosed = kFALSE;

 public:
  TGeoManager();
  TGeoManager(const char* name, const char* title);
  virtual ~TGeoManager();
  void SetTopVolume(TGeoVolume* v) { RbGeomTouch(); fTopVolume = v; }
  TGeoVolume* GetTopVolume() const { return fTopVolume; }
  virtual void CloseGeometry(Option_t* = "d") { RbGeomTouch(); fClosed = kTRUE; }
  Bool_t IsClosed() const { return fClosed; }
  void SetNsegments(Int_t) {}
  void SetMaxThreads(Int_t n) { fMaxThreads = n; fMultiThread = n >= 1; }
  Int_t GetMaxThreads() const { return fMaxThreads; }
  void SetMultiThread(Bool_t f = kTRUE) { fMultiThread = f; }
  Bool_t IsMultiThread() const { return fMultiThread; }
  void SetVisLevel(Int_t) {}
  static void SetVerboseLevel(Int_t) {}
};
inline TGeoManager*& gGeoManagerRef() {
  static TGeoManager* g = nullptr;
  return g;
}
#define gGeoManager (gGeoManagerRef())
inline TGeoManager::TGeoManager() { gGeoManager = this; }
inline TGeoManager::TGeoManager(const char* name, const char* title) : TNamed(name, title) { gGeoManager = this; }
inline TGeoManager::~TGeoManager() {
  if (gGeoManager == this) gGeoManager = nullptr;
}

struct TThread {
  static void Initialize() {}
};

// wall-clock / CPU timer (tutorials/multithread.C:32-35)
class TStopwatch {
  std::chrono::steady_clock::time_point fT0 = std::chrono::steady_clock::now();
  std::clock_t fC0 = std::clock();
  Double_t fReal = 0, fCpu = 0;
  Bool_t fRunning = kFALSE;

 public:
  void Start(Bool_t reset = kTRUE) {
    if (reset) fReal = fCpu = 0;
    fT0 = std::chrono::steady_clock::now();
    fC0 = std::clock();
    fRunning = kTRUE;
  }
  void Stop() {
    if (!fRunning) return;
    fReal += std::chrono::duration<Double_t>(std::chrono::steady_clock::now() - fT0).count();
    fCpu += Double_t(std::clock() - fC0) / CLOCKS_PER_SEC;
    fRunning = kFALSE;
  }
  Double_t RealTime() { Stop(); return fReal; }
  Double_t CpuTime() { Stop(); return fCpu; }
  void Print(const char* = "") { Stop(); printf("Real time %.6f s, CP time %.3f s\n", fReal, fCpu); }
};

// ---------------------------------------------------------------------------- TGraph / TGraph2D
struct TAxisStub {  // display-only axis handle
  void SetTitle(const char*) {}
  void SetLimits(Double_t, Double_t) {}
  void SetRangeUser(Double_t, Double_t) {}
};
class TGraph : public TNamed {
  std::vector<Double_t> fX, fY;
  TAxisStub fAxis;

 public:
  TGraph() {}
  TGraph(Int_t n) : fX(n), fY(n) {}
  TGraph(Int_t n, const Double_t* x, const Double_t* y) : fX(x, x + n), fY(y, y + n) {}
  void SetPoint(Int_t i, Double_t x, Double_t y) { RbGeomTouch();
    if (i >= (Int_t)fX.size()) { fX.resize(i + 1); fY.resize(i + 1); }
    fX[i] = x; fY[i] = y;
  }
  void AddPoint(Double_t x, Double_t y) { RbGeomTouch(); SetPoint(GetN(), x, y); }
  // drawing is out of scope; with ROBAST_DRAW_SUMMARY set, Draw() prints the points the plot would have shown
  void Draw(Option_t* = "") override {
    if (!getenv("ROBAST_DRAW_SUMMARY")) return;
    printf("TGraph name=\"%s\" n=%d points=", GetName(), GetN());
    for (Int_t i = 0; i < GetN(); i++) printf("%s%.9g:%.9g", i ? "," : "", fX[i], fY[i]);
    printf("\n");
  }
  Int_t GetN() const { return (Int_t)fX.size(); }
  const Double_t* GetX() const { return fX.data(); }
  const Double_t* GetY() const { return fY.data(); }
  Double_t GetPointX(Int_t i) const { return fX[i]; }
  Double_t GetPointY(Int_t i) const { return fY[i]; }
  // TGraph::Eval without spline: linear interpolation, linear extrapolation
  Double_t Eval(Double_t x) const {
    Int_t n = GetN();
    if (n == 0) return 0;
    if (n == 1) return fY[0];
    Int_t low = -1, up = -1, low2 = -1, up2 = -1;
    for (Int_t i = 0; i < n; ++i) {
      if (fX[i] < x) {
        if (low == -1 || fX[i] > fX[low]) { low2 = low; low = i; }
        else if (low2 == -1 || fX[i] > fX[low2]) low2 = i;
      } else if (fX[i] > x) {
        if (up == -1 || fX[i] < fX[up]) { up2 = up; up = i; }
        else if (up2 == -1 || fX[i] < fX[up2]) up2 = i;
      } else return fY[i];
    }
    if (up == -1) { up = low; low = low2; }
    if (low == -1) { low = up; up = up2; }
    if (fX[low] == fX[up]) return fY[low];
    return fY[up] + (x - fX[up]) * (fY[low] - fY[up]) / (fX[low] - fX[up]);
  }
  void SetLineStyle(Int_t) {}
  void SetMarkerStyle(Int_t) {}
  void SetMarkerColor(Int_t) {}
  void SetMarkerSize(Double_t) {}
  void SetLineColor(Int_t) {}
  void SetLineWidth(Int_t) {}
  TAxisStub* GetXaxis() { return &fAxis; }
  TAxisStub* GetYaxis() { return &fAxis; }
  void SetTitle(const char* t) override { TNamed::SetTitle(t); }
};

// TGraph2D::Interpolate = linear interpolation over the Delaunay triangulation of the (x, y) points, which ROOT
// builds on coordinates normalised to the data range (TGraphDelaunay / ROOT::Math::Delaunay2D); outside the convex
// hull it returns 0 (TGraph2D's default fZout).  The triangulation is done here, on the host, by Bowyer-Watson
// insertion; the exporter ships the triangle list to the device (rbg_graph2d).
class TGraph2D : public TNamed {
  std::vector<Double_t> fX, fY, fZ;
  mutable std::vector<Int_t> fTri;  // 3 vertex ids per triangle
  mutable Bool_t fTriValid = kFALSE;

  void Triangulate() const {
    fTri.clear();
    fTriValid = kTRUE;
    const Int_t n = GetN();
    if (n < 3) return;
    Double_t xmin = fX[0], xmax = fX[0], ymin = fY[0], ymax = fY[0];
    for (Int_t i = 1; i < n; i++) {
      xmin = std::min(xmin, fX[i]); xmax = std::max(xmax, fX[i]);
      ymin = std::min(ymin, fY[i]); ymax = std::max(ymax, fY[i]);
    }
    const Double_t sx = xmax > xmin ? 1. / (xmax - xmin) : 1., sy = ymax > ymin ? 1. / (ymax - ymin) : 1.;
    std::vector<Double_t> px(n + 3), py(n + 3);
    for (Int_t i = 0; i < n; i++) { px[i] = (fX[i] - xmin) * sx; py[i] = (fY[i] - ymin) * sy; }
    px[n] = -1000.; py[n] = -1000.; px[n + 1] = 1001.; py[n + 1] = -1000.; px[n + 2] = 0.5; py[n + 2] = 2000.;  // super-triangle
    struct T { Int_t v[3]; };
    std::vector<T> tris;
    tris.push_back(T{{n, n + 1, n + 2}});
    auto orient = [&](Int_t a, Int_t b, Int_t c) { return (px[b] - px[a]) * (py[c] - py[a]) - (py[b] - py[a]) * (px[c] - px[a]); };
    auto in_circle = [&](const T& t, Int_t p) {  // strictly inside the circumcircle of the (counter-clockwise) triangle
      Double_t ax = px[t.v[0]] - px[p], ay = py[t.v[0]] - py[p], bx = px[t.v[1]] - px[p], by = py[t.v[1]] - py[p], cx = px[t.v[2]] - px[p],
               cy = py[t.v[2]] - py[p];
      Double_t det = (ax * ax + ay * ay) * (bx * cy - cx * by) - (bx * bx + by * by) * (ax * cy - cx * ay) + (cx * cx + cy * cy) * (ax * by - bx * ay);
      return det > 1e-14;
    };
    for (Int_t p = 0; p < n; p++) {
      Bool_t dup = kFALSE;
      for (Int_t q = 0; q < p && !dup; q++) dup = px[q] == px[p] && py[q] == py[p];
      if (dup) continue;
      std::vector<T> keep;
      std::vector<std::pair<Int_t, Int_t>> edges;
      for (const T& t : tris) {
        if (in_circle(t, p)) {
          for (int e = 0; e < 3; e++) edges.emplace_back(t.v[e], t.v[(e + 1) % 3]);
        } else keep.push_back(t);
      }
      if (edges.empty()) {  // on a circumcircle of every neighbour (degenerate): fall back to the containing triangle
        for (size_t k = 0; k < keep.size(); k++) {
          const T t = keep[k];
          if (orient(t.v[0], t.v[1], p) >= 0 && orient(t.v[1], t.v[2], p) >= 0 && orient(t.v[2], t.v[0], p) >= 0) {
            for (int e = 0; e < 3; e++) edges.emplace_back(t.v[e], t.v[(e + 1) % 3]);
            keep.erase(keep.begin() + k);
            break;
          }
        }
      }
      for (size_t i = 0; i < edges.size(); i++) {  // boundary of the cavity = edges that appear once
        Bool_t shared = kFALSE;
        for (size_t j = 0; j < edges.size() && !shared; j++)
          shared = i != j && edges[i].first == edges[j].second && edges[i].second == edges[j].first;
        if (shared) continue;
        T t{{edges[i].first, edges[i].second, p}};
        if (orient(t.v[0], t.v[1], t.v[2]) < 0) std::swap(t.v[0], t.v[1]);
        if (orient(t.v[0], t.v[1], t.v[2]) > 0) keep.push_back(t);
      }
      tris.swap(keep);
    }
    for (const T& t : tris)
      if (t.v[0] < n && t.v[1] < n && t.v[2] < n) { fTri.push_back(t.v[0]); fTri.push_back(t.v[1]); fTri.push_back(t.v[2]); }
  }

 public:
  TGraph2D() {}
  TGraph2D(Int_t n, const Double_t* x, const Double_t* y, const Double_t* z) : fX(x, x + n), fY(y, y + n), fZ(z, z + n) {}
  void SetPoint(Int_t i, Double_t x, Double_t y, Double_t z) { RbGeomTouch();
    if (i >= (Int_t)fX.size()) { fX.resize(i + 1); fY.resize(i + 1); fZ.resize(i + 1); }
    fX[i] = x; fY[i] = y; fZ[i] = z;
    fTriValid = kFALSE;
  }
  void AddPoint(Double_t x, Double_t y, Double_t z) { RbGeomTouch(); SetPoint(GetN(), x, y, z); }
  Int_t GetN() const { return (Int_t)fX.size(); }
  const Double_t* GetX() const { return fX.data(); }
  const Double_t* GetY() const { return fY.data(); }
  const Double_t* GetZ() const { return fZ.data(); }
  const std::vector<Int_t>& GetTriangles() const {
    if (!fTriValid) Triangulate();
    return fTri;
  }
  // same arithmetic as the device/oracle evaluation of rbg_graph2d: first triangle (in list order) whose three
  // barycentric coordinates are >= -1e-9, linear interpolation inside it, 0 outside the hull
  Double_t Interpolate(Double_t x, Double_t y) const {
    const std::vector<Int_t>& t = GetTriangles();
    for (size_t k = 0; k + 2 < t.size(); k += 3) {
      Double_t x0 = fX[t[k]], y0 = fY[t[k]], x1 = fX[t[k + 1]], y1 = fY[t[k + 1]], x2 = fX[t[k + 2]], y2 = fY[t[k + 2]];
      Double_t den = (y1 - y2) * (x0 - x2) + (x2 - x1) * (y0 - y2);
      if (den == 0) continue;
      Double_t l0 = ((y1 - y2) * (x - x2) + (x2 - x1) * (y - y2)) / den, l1 = ((y2 - y0) * (x - x2) + (x0 - x2) * (y - y2)) / den, l2 = 1. - l0 - l1;
      if (l0 < -1e-9 || l1 < -1e-9 || l2 < -1e-9) continue;
      return l0 * fZ[t[k]] + l1 * fZ[t[k + 1]] + l2 * fZ[t[k + 2]];
    }
    return 0.;
  }
};

// ---------------------------------------------------------------------------- histograms
class TAxis {
 public:
  Int_t fN = 1;
  Double_t fMin = 0, fMax = 1;
  Int_t GetNbins() const { return fN; }
  Double_t GetXmin() const { return fMin; }
  Double_t GetXmax() const { return fMax; }
  Double_t GetBinWidth(Int_t) const { return (fMax - fMin) / fN; }
  Double_t GetBinCenter(Int_t bin) const { return fMin + (bin - 0.5) * (fMax - fMin) / fN; }
  Double_t GetBinLowEdge(Int_t bin) const { return fMin + (bin - 1) * (fMax - fMin) / fN; }
  Double_t GetBinUpEdge(Int_t bin) const { return fMin + bin * (fMax - fMin) / fN; }
  Int_t FindFixBin(Double_t x) const {
    if (x < fMin) return 0;
    if (!(x < fMax)) return fN + 1;
    return 1 + Int_t(fN * (x - fMin) / (fMax - fMin));
  }
  Int_t FindBin(Double_t x) const { return FindFixBin(x); }
  void SetTitle(const char*) {}
  void SetLimits(Double_t, Double_t) {}
  void SetRangeUser(Double_t, Double_t) {}
  void SetNdivisions(Int_t, Bool_t = kTRUE) {}
  void SetTitleOffset(Double_t = 1.) {}
};

class TH1 : public TNamed {
 protected:
  Double_t fEntries = 0;

 public:
  TH1() {}
  TH1(const char* n, const char* t) : TNamed(n, t) {}
  Double_t GetEntries() const { return fEntries; }
  void SetLineColor(Int_t) {}
  void SetMaximum(Double_t = -1111) {}
  void SetMinimum(Double_t = -1111) {}
};

class TH1D : public TH1 {
  TAxis fXaxis;
  std::vector<Double_t> fC;
  Double_t fSw = 0, fSwx = 0, fSwx2 = 0;

 public:
  TH1D() {}
  TH1D(const char* name, const char* title, Int_t n, Double_t lo, Double_t hi) : TH1(name, title), fC(n + 2, 0.) {
    fXaxis.fN = n; fXaxis.fMin = lo; fXaxis.fMax = hi;
  }
  Int_t Fill(Double_t x, Double_t w = 1) { RbGeomTouch();
    Int_t b = fXaxis.FindFixBin(x);
    fC[b] += w;
    fEntries += 1;
    if (b >= 1 && b <= fXaxis.fN) { fSw += w; fSwx += w * x; fSwx2 += w * x * x; }
    return b;
  }
  Int_t GetNbinsX() const { return fXaxis.fN; }
  TAxis* GetXaxis() { return &fXaxis; }
  Double_t GetBinContent(Int_t b) const { return fC[b]; }
  void SetBinContent(Int_t b, Double_t v) { RbGeomTouch(); fC[b] = v; }
  Double_t GetBinCenter(Int_t b) const { return fXaxis.GetBinCenter(b); }
  Double_t GetMean(Int_t = 1) const { return fSw ? fSwx / fSw : 0; }
  Double_t GetStdDev(Int_t = 1) const {
    if (!fSw) return 0;
    Double_t m = fSwx / fSw;
    return std::sqrt(std::fabs(fSwx2 / fSw - m * m));
  }
  Double_t GetRMS(Int_t a = 1) const { return GetStdDev(a); }
  void Draw(Option_t* = "") override {
    if (getenv("ROBAST_DRAW_SUMMARY"))
      printf("TH1 name=\"%s\" title=\"%s\" entries=%.0f inrange=%.9g mean=%.9g rms=%.9g\n", GetName(), GetTitle(), fEntries, fSw, GetMean(), GetStdDev());
  }
  Double_t Integral() const {
    Double_t s = 0;
    for (Int_t i = 1; i <= fXaxis.fN; i++) s += fC[i];
    return s;
  }
};

class TH2 : public TH1 {
 protected:
  TAxis fXaxis, fYaxis;
  std::vector<Double_t> fC;
  Double_t fSw = 0, fSwx = 0, fSwx2 = 0, fSwy = 0, fSwy2 = 0;

 public:
  TH2() {}
  TH2(const char* name, const char* title, Int_t nx, Double_t xlo, Double_t xhi, Int_t ny, Double_t ylo, Double_t yhi)
      : TH1(name, title), fC(size_t(nx + 2) * (ny + 2), 0.) {
    fXaxis.fN = nx; fXaxis.fMin = xlo; fXaxis.fMax = xhi;
    fYaxis.fN = ny; fYaxis.fMin = ylo; fYaxis.fMax = yhi;
  }
  Int_t GetBin(Int_t bx, Int_t by) const { return bx + (fXaxis.fN + 2) * by; }
  Int_t Fill(Double_t x, Double_t y, Double_t w = 1) { RbGeomTouch();
    Int_t bx = fXaxis.FindFixBin(x), by = fYaxis.FindFixBin(y);
    fC[GetBin(bx, by)] += w;
    fEntries += 1;
    if (bx >= 1 && bx <= fXaxis.fN && by >= 1 && by <= fYaxis.fN) {
      fSw += w; fSwx += w * x; fSwx2 += w * x * x; fSwy += w * y; fSwy2 += w * y * y;
    }
    return GetBin(bx, by);
  }
  Int_t GetNbinsX() const { return fXaxis.fN; }
  Int_t GetNbinsY() const { return fYaxis.fN; }
  TAxis* GetXaxis() { return &fXaxis; }
  TAxis* GetYaxis() { return &fYaxis; }
  const TAxis* GetXaxis() const { return &fXaxis; }
  const TAxis* GetYaxis() const { return &fYaxis; }
  Double_t GetBinContent(Int_t bx, Int_t by) const { return fC[GetBin(bx, by)]; }
  void SetBinContent(Int_t bx, Int_t by, Double_t v) { RbGeomTouch(); fC[GetBin(bx, by)] = v; }
  Double_t GetMean(Int_t axis = 1) const { return !fSw ? 0 : (axis == 1 ? fSwx / fSw : fSwy / fSw); }
  Double_t GetStdDev(Int_t axis = 1) const {
    if (!fSw) return 0;
    Double_t m = GetMean(axis), s2 = (axis == 1 ? fSwx2 : fSwy2) / fSw;
    return std::sqrt(std::fabs(s2 - m * m));
  }
  Double_t GetRMS(Int_t axis = 1) const { return GetStdDev(axis); }
  // sum w, wx, wy, wx^2, wy^2 of the in-range fills (ROOT's fTsumw...), for the device reducers
  void GetStats5(Double_t* s) const { s[0] = fSw; s[1] = fSwx; s[2] = fSwy; s[3] = fSwx2; s[4] = fSwy2; }
  Double_t Integral() const {
    Double_t s = 0;
    for (Int_t j = 1; j <= fYaxis.fN; j++)
      for (Int_t i = 1; i <= fXaxis.fN; i++) s += fC[GetBin(i, j)];
    return s;
  }
  // drawing is out of scope; with ROBAST_DRAW_SUMMARY set, Draw() prints what the plot would have shown
  void Draw(Option_t* = "") override {
    if (getenv("ROBAST_DRAW_SUMMARY"))
      printf("TH2 name=\"%s\" title=\"%s\" entries=%.0f inrange=%.0f meanx=%.9g meany=%.9g rmsx=%.9g rmsy=%.9g\n", GetName(), GetTitle(), fEntries, fSw,
             GetMean(1), GetMean(2), GetStdDev(1), GetStdDev(2));
  }
  // TH2::Interpolate: bilinear between the four surrounding bin centres
  Double_t Interpolate(Double_t x, Double_t y) const {
    Int_t bin_x = fXaxis.FindFixBin(x), bin_y = fYaxis.FindFixBin(y);
    if (bin_x < 1 || bin_x > fXaxis.fN || bin_y < 1 || bin_y > fYaxis.fN) {
      Error("Interpolate", "Cannot interpolate outside histogram domain.");
      return 0;
    }
    Double_t dx = fXaxis.GetBinUpEdge(bin_x) - x, dy = fYaxis.GetBinUpEdge(bin_y) - y;
    Double_t hx = fXaxis.GetBinWidth(bin_x) / 2, hy = fYaxis.GetBinWidth(bin_y) / 2;
    Int_t ix1 = dx <= hx ? bin_x : bin_x - 1, iy1 = dy <= hy ? bin_y : bin_y - 1;
    Double_t x1 = fXaxis.GetBinCenter(ix1), x2 = fXaxis.GetBinCenter(ix1 + 1);
    Double_t y1 = fYaxis.GetBinCenter(iy1), y2 = fYaxis.GetBinCenter(iy1 + 1);
    Int_t bx1 = std::max(ix1, 1), bx2 = std::min(ix1 + 1, fXaxis.fN);
    Int_t by1 = std::max(iy1, 1), by2 = std::min(iy1 + 1, fYaxis.fN);
    Double_t q11 = GetBinContent(bx1, by1), q12 = GetBinContent(bx1, by2), q21 = GetBinContent(bx2, by1), q22 = GetBinContent(bx2, by2);
    Double_t d = 1.0 * (x2 - x1) * (y2 - y1);
    return 1.0 * q11 / d * (x2 - x) * (y2 - y) + 1.0 * q21 / d * (x - x1) * (y2 - y) + 1.0 * q12 / d * (x2 - x) * (y - y1) +
           1.0 * q22 / d * (x - x1) * (y - y1);
  }
};
class TH2D : public TH2 {
 public:
  using TH2::TH2;
};

#endif  // ROBAST_ROOTCOMPAT_H
