// Robast.h — host-side mirror of the ROBAST classes on the TraceNonSequential path
// (AOpticsManager / ARay / ARayArray / ARayShooter / AOpticalComponent & subclasses /
// ABorderSurfaceCondition / AMultilayer / ARefractiveIndex family / AGlassCatalog / AGeo* shapes /
// AGeoUtil), same names, argument meaning and error behaviour as the reference headers
// (/root/reference/include/*.h, cited per class).  These classes only DESCRIBE the scene and own
// ray buffers; AOpticsManager::TraceNonSequential flattens the scene to rbg_scene_desc and calls
// the C ABI (include/robast_b200.h) — all tracing arithmetic runs in the CUDA library.
#ifndef ROBAST_ROBAST_H
#define ROBAST_ROBAST_H

#include <complex>
#include <fstream>
#include <iostream>
#include <functional>
#include <sstream>

#include "../robast_b200.h"
#include "RootCompat.h"

// ============================================================================ refractive indices
// reference include/ARefractiveIndex.h:25-67, src/ARefractiveIndex.cxx:19-40
class AMixedRefractiveIndex;
class ARefractiveIndex : public TObject {
 protected:
  std::shared_ptr<TGraph> fRefractiveIndex;
  std::shared_ptr<TGraph> fExtinctionCoefficient;

 public:
  ARefractiveIndex() {}
  ARefractiveIndex(Double_t n, Double_t k = 0.) {
    fRefractiveIndex = std::make_shared<TGraph>();
    fRefractiveIndex->SetPoint(0, 0, n);
    if (k > 0) {
      fExtinctionCoefficient = std::make_shared<TGraph>();
      fExtinctionCoefficient->SetPoint(0, 0, k);
    }
  }
  virtual ~ARefractiveIndex() {}
  virtual Int_t Kind() const { return RBG_INDEX_GRAPH; }
  virtual const Double_t* Par() const { return nullptr; }
  virtual Double_t GetAbbeNumber() const {
    const Double_t nm = 1e-7;
    Double_t nC = GetRefractiveIndex(656.2725 * nm), nD = GetRefractiveIndex(589.2938 * nm), nF = GetRefractiveIndex(486.1327 * nm);
    return (nD - 1.) / (nF - nC);
  }
  virtual Double_t GetRefractiveIndex(Double_t lambda) const { return fRefractiveIndex ? fRefractiveIndex->Eval(lambda) : 1.; }
  virtual Double_t GetExtinctionCoefficient(Double_t lambda) const { return fExtinctionCoefficient ? fExtinctionCoefficient->Eval(lambda) : 0.; }
  virtual Double_t GetAbsorptionLength(Double_t lambda) const {
    Double_t k = GetExtinctionCoefficient(lambda);
    return k <= 0. ? std::numeric_limits<Double_t>::infinity() : ExtinctionCoefficientToAbsorptionLength(k, lambda);
  }
  virtual std::complex<Double_t> GetComplexRefractiveIndex(Double_t lambda) const {
    return std::complex<Double_t>(GetRefractiveIndex(lambda), GetExtinctionCoefficient(lambda));
  }
  virtual void SetExtinctionCoefficient(std::shared_ptr<TGraph> graph) { RbGeomTouch(); fExtinctionCoefficient = graph; }
  virtual void SetRefractiveIndex(std::shared_ptr<TGraph> graph) { RbGeomTouch(); fRefractiveIndex = graph; }
  std::shared_ptr<TGraph> GetRefractiveIndexGraph() const { return fRefractiveIndex; }
  std::shared_ptr<TGraph> GetExtinctionCoefficientGraph() const { return fExtinctionCoefficient; }
  static Double_t AbsorptionLengthToExtinctionCoefficient(Double_t a, Double_t lambda) { return lambda / (4 * TMath::Pi() * a); }
  static Double_t ExtinctionCoefficientToAbsorptionLength(Double_t k, Double_t lambda) { return lambda / (4 * TMath::Pi() * k); }
};

// reference src/ASellmeierFormula.cxx:27-54 (λ in µm inside the formula; 1 µm = 1e-4 cm)
class ASellmeierFormula : public ARefractiveIndex {
  Double_t fPar[6];

 public:
  ASellmeierFormula(Double_t B1, Double_t B2, Double_t B3, Double_t C1, Double_t C2, Double_t C3) : fPar{B1, B2, B3, C1, C2, C3} {}
  ASellmeierFormula(const Double_t* p) { memcpy(fPar, p, sizeof(fPar)); }
  Int_t Kind() const override { return RBG_INDEX_SELLMEIER; }
  const Double_t* Par() const override { return fPar; }
  Double_t GetRefractiveIndex(Double_t lambda) const override {
    lambda /= 1e-4;
    Double_t l2 = lambda * lambda;
    return std::sqrt(1 + fPar[0] * l2 / (l2 - fPar[3]) + fPar[1] * l2 / (l2 - fPar[4]) + fPar[2] * l2 / (l2 - fPar[5]));
  }
};

// reference src/ASchottFormula.cxx:24-55
class ASchottFormula : public ARefractiveIndex {
  Double_t fPar[6];

 public:
  ASchottFormula(Double_t A0, Double_t A1, Double_t A2, Double_t A3, Double_t A4, Double_t A5) : fPar{A0, A1, A2, A3, A4, A5} {}
  ASchottFormula(const Double_t* p) { memcpy(fPar, p, sizeof(fPar)); }
  Int_t Kind() const override { return RBG_INDEX_SCHOTT; }
  const Double_t* Par() const override { return fPar; }
  Double_t GetRefractiveIndex(Double_t lambda) const override {
    lambda /= 1e-4;
    return std::sqrt(fPar[0] + fPar[1] * std::pow(lambda, 2.) + fPar[2] * std::pow(lambda, -2.) + fPar[3] * std::pow(lambda, -4.) +
                     fPar[4] * std::pow(lambda, -6.) + fPar[5] * std::pow(lambda, -8.));
  }
};

// reference src/ACauchyFormula.cxx:24-46
class ACauchyFormula : public ARefractiveIndex {
  Double_t fPar[6] = {0, 0, 0, 0, 0, 0};

 public:
  ACauchyFormula(Double_t A, Double_t B, Double_t C) { fPar[0] = A; fPar[1] = B; fPar[2] = C; }
  ACauchyFormula(const Double_t* p) { memcpy(fPar, p, 3 * sizeof(Double_t)); }
  Int_t Kind() const override { return RBG_INDEX_CAUCHY; }
  const Double_t* Par() const override { return fPar; }
  Double_t GetRefractiveIndex(Double_t lambda) const override {
    lambda /= 1e-4;
    return fPar[0] + fPar[1] * std::pow(lambda, -2) + fPar[2] * std::pow(lambda, -4);
  }
};

// reference include/AMixedRefractiveIndex.h:25-50, src/AMixedRefractiveIndex.cxx
class AMixedRefractiveIndex : public ARefractiveIndex {
  std::shared_ptr<ARefractiveIndex> fMaterialA, fMaterialB;
  Double_t fFractionA = 0.5, fFractionB = 0.5;

 public:
  AMixedRefractiveIndex(std::shared_ptr<ARefractiveIndex> a, std::shared_ptr<ARefractiveIndex> b, Double_t fa, Double_t fb)
      : fMaterialA(a), fMaterialB(b) {
    SetFraction(fa, fb);
  }
  Int_t Kind() const override { return RBG_INDEX_MIXED; }
  std::shared_ptr<ARefractiveIndex> GetA() const { return fMaterialA; }
  std::shared_ptr<ARefractiveIndex> GetB() const { return fMaterialB; }
  Double_t GetFractionA() const { return fFractionA; }
  Double_t GetFractionB() const { return fFractionB; }
  Double_t GetRefractiveIndex(Double_t lambda) const override {
    return fMaterialA->GetRefractiveIndex(lambda) * fFractionA + fMaterialB->GetRefractiveIndex(lambda) * fFractionB;
  }
  Double_t GetExtinctionCoefficient(Double_t lambda) const override {
    return fMaterialA->GetExtinctionCoefficient(lambda) * fFractionA + fMaterialB->GetExtinctionCoefficient(lambda) * fFractionB;
  }
  void SetFraction(Double_t fa, Double_t fb) { RbGeomTouch();
    fFractionA = fa / (fa + fb);
    fFractionB = fb / (fa + fb);
  }
};

// reference src/AFilmetrixDotCom.cxx:24-62 — "Wavelength(nm)\tn\tk" tables (optional UTF-8 BOM / CR)
class AFilmetrixDotCom : public ARefractiveIndex {
 public:
  AFilmetrixDotCom(const char* fname) {
    std::ifstream fin(fname);
    if (!fin.is_open()) {
      Error("AFilmetrixDotCom", "Cannot open %s", fname);
      return;
    }
    std::string head;
    std::getline(fin, head);
    if (head.size() >= 3 && (unsigned char)head[0] == 0xef && (unsigned char)head[1] == 0xbb && (unsigned char)head[2] == 0xbf) head = head.substr(3);
    if (!head.empty() && head.back() == '\r') head.pop_back();
    if (head != "Wavelength(nm)\tn\tk") {
      Error("AFilmetrixDotCom", "Invalid data format");
      return;
    }
    fRefractiveIndex = std::make_shared<TGraph>();
    fExtinctionCoefficient = std::make_shared<TGraph>();
    double wl, n, k;
    while (fin >> wl >> n >> k) {
      wl *= 1e-7;
      fRefractiveIndex->SetPoint(fRefractiveIndex->GetN(), wl, n);
      fExtinctionCoefficient->SetPoint(fExtinctionCoefficient->GetN(), wl, k);
    }
  }
};

// reference src/ARefractiveIndexDotInfo.cxx:24-105 — tables from refractiveindex.info: a "wl,n" (or tab-separated) block in um,
// optionally followed by a "wl,k" block; LF or CRLF line ends
class ARefractiveIndexDotInfo : public ARefractiveIndex {
 public:
  ARefractiveIndexDotInfo(const char* fname) {
    std::ifstream fin(fname);
    if (!fin.is_open()) {
      Error("ARefractiveIndexDotInfo", "Cannot open %s", fname);
      return;
    }
    std::string line;
    std::getline(fin, line);
    if (!line.empty() && line.back() == '\r') line.pop_back();
    char sep;
    if (line == "wl,n") sep = ',';
    else if (line == "wl\tn") sep = '\t';
    else {
      Error("ARefractiveIndexDotInfo", "Invalid data format");
      return;
    }
    fRefractiveIndex = std::make_shared<TGraph>();
    std::shared_ptr<TGraph> cur = fRefractiveIndex;
    while (std::getline(fin, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      size_t pos = line.find(sep);
      if (pos == std::string::npos) break;
      std::string a = line.substr(0, pos), b = line.substr(pos + 1);
      char *e1, *e2;
      double wl = strtod(a.c_str(), &e1), v = strtod(b.c_str(), &e2);
      if (*e1 != '\0' || *e2 != '\0' || a.empty() || b.empty()) {  // a header line: only "wl<sep>k" continues the file
        if (cur == fRefractiveIndex && a == "wl" && b == "k") {
          fExtinctionCoefficient = std::make_shared<TGraph>();
          cur = fExtinctionCoefficient;
          continue;
        }
        break;
      }
      cur->SetPoint(cur->GetN(), wl * 1e-4, v);  // um -> cm
    }
  }
};

// reference src/AGlassCatalog.cxx:27-121 — Zemax AGF parser (formula 2 = Sellmeier only, as there)
class AGlassCatalog : public TObject {
  std::map<std::string, std::shared_ptr<ARefractiveIndex>> fIndexMap;

 public:
  AGlassCatalog() {}
  AGlassCatalog(const std::string& catalog_file) {
    std::ifstream fin(catalog_file.c_str());
    if (!fin.is_open()) {
      Error("AGlassCatalog", "Cannot open %s", catalog_file.c_str());
      return;
    }
    auto ends_with = [&](const char* suf) { return catalog_file.size() >= 4 && catalog_file.compare(catalog_file.size() - 4, 4, suf) == 0; };
    if (!ends_with(".agf") && !ends_with(".AGF")) {
      Error("AGlassCatalog", "Cannot read a non-ZEMAX file");
      return;
    }
    std::string line, glass;
    int formula = 0;
    std::shared_ptr<TGraph> graph;
    while (std::getline(fin, line)) {
      if (line.compare(0, 3, "NM ") == 0) {
        char name[256], product[256];
        double nd, vd;
        if (sscanf(line.c_str(), "NM %255s %d %255s %lf %lf", name, &formula, product, &nd, &vd) != 5)
          Warning("AGlassCatalog", "Bad format line found: %s", line.c_str());
        glass = name;
        graph = std::make_shared<TGraph>();
        fIndexMap.insert(std::make_pair(glass, std::shared_ptr<ARefractiveIndex>()));
      } else if (line.compare(0, 3, "CD ") == 0) {
        double cd[8];
        int ret = sscanf(line.c_str(), "CD %lf %lf %lf %lf %lf %lf %lf %lf", &cd[0], &cd[1], &cd[2], &cd[3], &cd[4], &cd[5], &cd[6], &cd[7]);
        if (formula == 2 && ret >= 6) {
          auto it = fIndexMap.find(glass);
          if (it != fIndexMap.end()) {
            it->second = std::make_shared<ASellmeierFormula>(cd[0], cd[2], cd[4], cd[1], cd[3], cd[5]);
            it->second->SetExtinctionCoefficient(graph);
          }
        }
      } else if (line.compare(0, 3, "IT ") == 0) {
        double wl, T, d;
        int ret = sscanf(line.c_str(), "IT %lf %lf %lf", &wl, &T, &d);
        if (ret == 3) {
          wl *= 1e-4;  // um
          d *= 0.1;    // mm
          Double_t absl = -d / std::log(T);
          if (graph) graph->SetPoint(graph->GetN(), wl, ARefractiveIndex::AbsorptionLengthToExtinctionCoefficient(absl, wl));
        } else if (ret != 2) {
          Warning("AGlassCatalog", "Bad format line found: %s", line.c_str());
        }
      }
    }
  }
  std::shared_ptr<ARefractiveIndex> GetRefractiveIndex(const std::string& name) {
    auto it = fIndexMap.find(name);
    return it == fIndexMap.end() ? std::shared_ptr<ARefractiveIndex>() : it->second;
  }
};

// ============================================================================ AGeo shapes (description + host helpers)
// reference src/AGeoAsphericDisk.cxx:96-243,971-1024,1187-1204
class AGeoAsphericDisk : public TGeoBBox {
  Double_t fZ1, fZ2, fCurve1, fCurve2, fConic1 = 0, fConic2 = 0, fKappa1 = 1, fKappa2 = 1, fRmin, fRmax;
  std::vector<Double_t> fK1, fK2;
  Int_t fSteps = 100, fRepeat = 4;

  static Double_t Sag(Double_t z0, Double_t c, Double_t kappa, const std::vector<Double_t>& K, Double_t r, bool& ok) {
    Double_t p = r * r * c * c * kappa;
    if (1 - p < 0) { ok = false; return 0; }
    Double_t ret = z0 + r * r * c / (1 + std::sqrt(1 - p));
    for (size_t i = 0; i < K.size(); i++) ret += K[i] * std::pow(r, 2 * (Int_t(i) + 1));
    ok = true;
    return ret;
  }
  static Double_t Slope(Double_t c, Double_t kappa, const std::vector<Double_t>& K, Double_t r) {
    Double_t p = r * r * c * c * kappa;
    if (1 - p <= 0) throw std::runtime_error("AGeoAsphericDisk: slope undefined");
    Double_t ret = r * c / std::sqrt(1 - p);
    for (size_t i = 0; i < K.size(); i++) ret += 2 * (Int_t(i) + 1) * K[i] * std::pow(r, 2 * (Int_t(i) + 1) - 1);
    return ret;
  }
  Double_t Extremum(int surf, bool want_max) const {
    const Double_t Big = TGeoShape::Big();
    const auto& K = surf == 1 ? fK1 : fK2;
    auto F = [&](Double_t r, bool& ok) { return surf == 1 ? Sag(fZ1, fCurve1, fKappa1, fK1, r, ok) : Sag(fZ2, fCurve2, fKappa2, fK2, r, ok); };
    if (K.empty()) {
      bool ok1, ok2;
      Double_t f1 = F(fRmin, ok1), f2 = F(fRmax, ok2);
      if (!ok1 || !ok2) return want_max ? Big : -Big;
      return want_max ? std::max(f1, f2) : std::min(f1, f2);
    }
    Double_t best = want_max ? -Big : Big, r1 = fRmin, r2 = fRmax;
    for (Int_t i = 0; i < fRepeat; i++) {
      Double_t step = (r2 - r1) / fSteps, r_ = r1;
      for (Int_t j = 0; j <= fSteps + 1; j++) {
        Double_t r = r1 + j * step;
        bool ok;
        Double_t f = F(r, ok);
        if (!ok) f = want_max ? -Big : Big;
        if (want_max ? f > best : f < best) { best = f; r_ = r; }
      }
      r1 = r_ == fRmin ? fRmin : r_ - step;
      r2 = r_ == fRmax ? fRmax : r_ + step;
    }
    return best;
  }

 public:
  AGeoAsphericDisk(Double_t z1, Double_t curve1, Double_t z2, Double_t curve2, Double_t rmax, Double_t rmin = 0.) {
    SetAsphDimensions(z1, curve1, z2, curve2, rmax, rmin);
    ComputeBBox();
  }
  AGeoAsphericDisk(const char* name, Double_t z1, Double_t curve1, Double_t z2, Double_t curve2, Double_t rmax, Double_t rmin = 0.) {
    SetName(name);
    SetAsphDimensions(z1, curve1, z2, curve2, rmax, rmin);
    ComputeBBox();
  }
  EKind Kind() const override { return kAsphere; }
  void SetAsphDimensions(Double_t z1, Double_t curve1, Double_t z2, Double_t curve2, Double_t rmax, Double_t rmin) { RbGeomTouch();
    if (z1 < z2) { fZ1 = z1; fZ2 = z2; fCurve1 = curve1; fCurve2 = curve2; }
    else { fZ1 = z2; fZ2 = z1; fCurve1 = curve2; fCurve2 = curve1; }
    rmax = std::fabs(rmax); rmin = std::fabs(rmin);
    fRmax = std::max(rmax, rmin); fRmin = std::min(rmax, rmin);
    fK1.clear(); fK2.clear();
  }
  void ComputeBBox() {
    Double_t zmax = Extremum(2, true), zmin = Extremum(1, false);
    fOrigin[0] = fOrigin[1] = 0;
    fOrigin[2] = (zmax + zmin) / 2;
    fDX = fDY = fRmax;
    fDZ = (zmax - zmin) / 2;
  }
  void SetConicConstants(Double_t conic1, Double_t conic2) { RbGeomTouch();
    fConic1 = conic1; fConic2 = conic2; fKappa1 = conic1 + 1; fKappa2 = conic2 + 1;
    ComputeBBox();
  }
  void SetPolynomials(Int_t n1, const Double_t* k1, Int_t n2, const Double_t* k2) { RbGeomTouch();
    fK1.assign(k1 ? k1 : nullptr, k1 ? k1 + (n1 > 0 ? n1 : 0) : nullptr);
    fK2.assign(k2 ? k2 : nullptr, k2 ? k2 + (n2 > 0 ? n2 : 0) : nullptr);
    ComputeBBox();
  }
  void SetFineness(Int_t steps, Int_t repeat) {
    if (steps > 0) fSteps = steps;
    if (repeat > 0) fRepeat = repeat;
  }
  Double_t CalcF1(Double_t r) const {
    bool ok;
    Double_t f = Sag(fZ1, fCurve1, fKappa1, fK1, r, ok);
    if (!ok) throw std::runtime_error("AGeoAsphericDisk::CalcF1: out of domain");
    return f;
  }
  Double_t CalcF2(Double_t r) const {
    bool ok;
    Double_t f = Sag(fZ2, fCurve2, fKappa2, fK2, r, ok);
    if (!ok) throw std::runtime_error("AGeoAsphericDisk::CalcF2: out of domain");
    return f;
  }
  Double_t CalcdF1dr(Double_t r) const { return Slope(fCurve1, fKappa1, fK1, r); }
  Double_t CalcdF2dr(Double_t r) const { return Slope(fCurve2, fKappa2, fK2, r); }
  Double_t GetCurve1() const { return fCurve1; }
  Double_t GetCurve2() const { return fCurve2; }
  Double_t GetConic1() const { return fConic1; }
  Double_t GetConic2() const { return fConic2; }
  Double_t GetKappa1() const { return fKappa1; }
  Double_t GetKappa2() const { return fKappa2; }
  const Double_t* GetK1() const { return fK1.data(); }
  const Double_t* GetK2() const { return fK2.data(); }
  Int_t GetNPol1() const { return (Int_t)fK1.size(); }
  Int_t GetNPol2() const { return (Int_t)fK2.size(); }
  Double_t GetRmax() const { return fRmax; }
  Double_t GetRmin() const { return fRmin; }
  Double_t GetZ1() const { return fZ1; }
  Double_t GetZ2() const { return fZ2; }
};

// reference src/AGeoWinstonCone2D.cxx:59-98,551-567
class AGeoWinstonCone2D : public TGeoBBox {
 protected:
  Double_t fR1, fR2, fF, fTheta;

  void SetBase(Double_t r1, Double_t r2) { RbGeomTouch();
    fR1 = std::max(std::fabs(r1), std::fabs(r2));
    fR2 = std::min(std::fabs(r1), std::fabs(r2));
    fTheta = std::asin(fR2 / fR1);
    fDZ = (fR1 + fR2) / std::tan(fTheta) / 2.;
    fF = fR2 * (1 + std::sin(fTheta));
  }
  AGeoWinstonCone2D() {}

 public:
  AGeoWinstonCone2D(Double_t r1, Double_t r2, Double_t y) { SetBase(r1, r2); fDY = std::fabs(y); fDX = fR1; }
  AGeoWinstonCone2D(const char* name, Double_t r1, Double_t r2, Double_t y) { SetName(name); SetBase(r1, r2); fDY = std::fabs(y); fDX = fR1; }
  EKind Kind() const override { return kWinston2D; }
  Double_t CalcR(Double_t z) const {
    if (std::fabs(z) > fDZ + 1e-10) throw std::runtime_error("AGeoWinstonCone2D::CalcR: |z| > DZ");
    Double_t sint = std::sin(fTheta), cost = std::cos(fTheta), t = z + fDZ;
    Double_t a0 = t * t * sint * sint - 4. * fF * (t * cost + fF), a1 = 2. * t * sint * cost + 4. * fF * sint, a2 = cost * cost;
    return (-a1 + std::sqrt(a1 * a1 - 4. * a0 * a2)) / (2 * a2) - fR2;
  }
  Double_t CalcdRdZ(Double_t z) const {
    if (std::fabs(z) > fDZ + 1e-10) throw std::runtime_error("AGeoWinstonCone2D::CalcdRdZ: |z| > DZ");
    Double_t sint = std::sin(fTheta), cost = std::cos(fTheta), t = z + fDZ;
    Double_t a0 = t * t * sint * sint - 4. * fF * (t * cost + fF), a1 = 2. * t * sint * cost + 4. * fF * sint, a2 = cost * cost;
    Double_t da0dt = 2 * t * sint * sint - 4 * fF * cost, da1dt = 2 * sint * cost;
    return (-da1dt + (a1 * da1dt - 2 * da0dt * a2) / std::sqrt(a1 * a1 - 4 * a0 * a2)) / (2 * a2);
  }
  Double_t GetR1() const { return fR1; }
  Double_t GetR2() const { return fR2; }
  Double_t GetF() const { return fF; }
  Double_t GetTheta() const { return fTheta; }
};

// reference src/AGeoWinstonConePoly.cxx:30-64,356-386
class AGeoWinstonConePoly : public AGeoWinstonCone2D {
  Int_t fPolyN;
  void SetPoly(Double_t r1, Double_t r2, Int_t n) { RbGeomTouch();
    SetBase(r1, r2);
    fPolyN = n >= 3 ? n : 3;
    Double_t r = r1 / std::cos(TMath::Pi() / n);
    fDX = fDY = 0;
    for (Int_t i = 0; i < fPolyN; i++) {
      fDX = std::max(std::fabs(r * std::cos(TMath::Pi() / n * (2 * i + 1))), fDX);
      fDY = std::max(std::fabs(r * std::sin(TMath::Pi() / n * (2 * i + 1))), fDY);
    }
  }

 public:
  AGeoWinstonConePoly(Double_t r1, Double_t r2, Int_t n) { SetPoly(r1, r2, n); }
  AGeoWinstonConePoly(const char* name, Double_t r1, Double_t r2, Int_t n) { SetName(name); SetPoly(r1, r2, n); }
  EKind Kind() const override { return kWinstonPoly; }
  Int_t GetPolyN() const { return fPolyN; }
};

// reference src/AGeoBezierPgon.cxx:44-104 / src/AGeoBezierPcon.cxx:43-103: z-sections sampled on a Bezier profile
template <class Base> class ABezierSections : public Base {
 protected:
  Int_t fNcontrol = 0;
  Double_t fPr[2] = {0, 0}, fPz[2] = {0, 0}, fR1, fR2, fDZb;

 public:
  template <class... A> ABezierSections(Double_t r1, Double_t r2, Double_t dz, A... a) : Base(a...), fR1(r1), fR2(r2), fDZb(dz) { SetSections(); }
  void Bezier(Double_t t, Double_t& r, Double_t& z) const {
    // control points: P0=(r2,-dz) ... P_last=(r1,+dz); fP in units of (r1-r2, 2dz) relative to P0
    Double_t p0r = fR2, p0z = -fDZb, pLr = fR1, pLz = fDZb;
    if (fNcontrol == 0) { r = (1 - t) * p0r + t * pLr; z = (1 - t) * p0z + t * pLz; }
    else if (fNcontrol == 1) {
      Double_t c1r = fPr[0] * (fR1 - fR2) + fR2, c1z = fPz[0] * 2 * fDZb - fDZb;
      r = (1 - t) * (1 - t) * p0r + 2 * (1 - t) * t * c1r + t * t * pLr;
      z = (1 - t) * (1 - t) * p0z + 2 * (1 - t) * t * c1z + t * t * pLz;
    } else {
      Double_t c1r = fPr[0] * (fR1 - fR2) + fR2, c1z = fPz[0] * 2 * fDZb - fDZb, c2r = fPr[1] * (fR1 - fR2) + fR2, c2z = fPz[1] * 2 * fDZb - fDZb;
      Double_t u = 1 - t;
      r = u * u * u * p0r + 3 * u * u * t * c1r + 3 * u * t * t * c2r + t * t * t * pLr;
      z = u * u * u * p0z + 3 * u * u * t * c1z + 3 * u * t * t * c2z + t * t * t * pLz;
    }
  }
  void SetControlPoints(Double_t r1, Double_t z1) { RbGeomTouch(); fNcontrol = 1; fPr[0] = r1; fPz[0] = z1; SetSections(); }
  void SetControlPoints(Double_t r1, Double_t z1, Double_t r2, Double_t z2) { RbGeomTouch();
    fNcontrol = 2; fPr[0] = r1; fPz[0] = z1; fPr[1] = r2; fPz[1] = z2;
    SetSections();
  }
  void SetSections() { RbGeomTouch();
    for (Int_t i = 0; i < this->fNz; i++) {
      Double_t t = Double_t(i) / (this->fNz - 1), r, z;
      Bezier(t, r, z);
      this->DefineSection(i, z, 0, r);
    }
  }
};
class AGeoBezierPgon : public ABezierSections<TGeoPgon> {
 public:
  AGeoBezierPgon(Double_t phi, Double_t dphi, Int_t nedges, Int_t nz, Double_t r1, Double_t r2, Double_t dz)
      : ABezierSections<TGeoPgon>(r1, r2, dz, phi, dphi, nedges, nz) {}
  AGeoBezierPgon(const char* name, Double_t phi, Double_t dphi, Int_t nedges, Int_t nz, Double_t r1, Double_t r2, Double_t dz)
      : ABezierSections<TGeoPgon>(r1, r2, dz, name, phi, dphi, nedges, nz) {}
};
class AGeoBezierPcon : public ABezierSections<TGeoPcon> {
 public:
  AGeoBezierPcon(Double_t phi, Double_t dphi, Int_t nz, Double_t r1, Double_t r2, Double_t dz) : ABezierSections<TGeoPcon>(r1, r2, dz, phi, dphi, nz) {}
  AGeoBezierPcon(const char* name, Double_t phi, Double_t dphi, Int_t nz, Double_t r1, Double_t r2, Double_t dz)
      : ABezierSections<TGeoPcon>(r1, r2, dz, name, phi, dphi, nz) {}
};

// reference src/AGeoUtil.cxx:128-195
namespace AGeoUtil {
inline void MakePointToPointTube(const char* name, const TVector3& v1, const TVector3& v2, Double_t rmin, Double_t rmax, TGeoTube** tube,
                                 TGeoCombiTrans** combi) {
  TVector3 v3 = (v1 + v2) * 0.5, v4 = v3 - v1;
  Double_t theta = v4.Theta() * TMath::RadToDeg(), phi = v4.Phi() * TMath::RadToDeg();
  *tube = new TGeoTube(Form("%stube", name), rmin, rmax, v4.Mag());
  *combi = new TGeoCombiTrans(TGeoTranslation(v3.X(), v3.Y(), v3.Z()), TGeoRotation("", phi + 90, theta, 0));
  (*combi)->SetName(Form("%scombi", name));
  (*combi)->RegisterYourself();
}
inline void MakePointToPointTube(const char* name, const TVector3& v1, const TVector3& v2, Double_t radius, TGeoTube** tube, TGeoCombiTrans** combi) {
  MakePointToPointTube(name, v1, v2, 0, radius, tube, combi);
}
inline void MakePointToPointBBox(const char* name, const TVector3& v1, const TVector3& v2, Double_t dx, Double_t dy, TGeoBBox** box,
                                 TGeoCombiTrans** combi) {
  TVector3 v3 = (v1 + v2) * 0.5, v4 = v3 - v1;
  Double_t theta = v4.Theta() * TMath::RadToDeg(), phi = v4.Phi() * TMath::RadToDeg();
  *box = new TGeoBBox(Form("%sbox", name), dx, dy, v4.Mag());
  *combi = new TGeoCombiTrans(TGeoTranslation(v3.X(), v3.Y(), v3.Z()), TGeoRotation("", phi + 90, theta, 0));
  (*combi)->SetName(Form("%scombi", name));
  (*combi)->RegisterYourself();
}
// reference src/AGeoUtil.cxx:47-82: an Arb8 prism from the four corners of its top face (clockwise seen from the top) and the
// corner of the bottom face below the first one
inline TGeoRotation AxisRotation(Double_t theta, Double_t phi) {  // TGeoRotation("",0,0,phi+90) * TGeoRotation("",0,theta,0), angles in rad
  TGeoRotation rot("", 0, 0, phi * TMath::RadToDeg() + 90), tilt("", 0, theta * TMath::RadToDeg(), 0);
  rot.MultiplyBy(&tilt, kTRUE);
  return rot;
}
inline void MakeArb8FromPoints(const char* name, const TVector3& v1, const TVector3& v2, const TVector3& v3, const TVector3& v4, const TVector3& v5,
                               TGeoArb8** arb8, TGeoCombiTrans** combi) {
  TVector3 normal = v1 - v5;
  Double_t dZ = normal.Mag() / 2., theta = normal.Theta(), phi = normal.Phi();
  TVector3 v[4] = {TVector3(0, 0, 0), v2 - v1, v3 - v1, v4 - v1};
  Double_t vertices[16] = {0};
  for (Int_t i = 1; i <= 3; ++i) {
    v[i].RotateZ(-phi - TMath::Pi() / 2.);
    v[i].RotateX(-theta);
    vertices[2 * i] = vertices[2 * i + 8] = v[i].X();
    vertices[2 * i + 1] = vertices[2 * i + 9] = v[i].Y();
  }
  *arb8 = new TGeoArb8(name, dZ, vertices);
  TGeoTranslation tr(v5.X() + normal.X() / 2., v5.Y() + normal.Y() / 2., v5.Z() + normal.Z() / 2.);
  *combi = new TGeoCombiTrans(tr, AxisRotation(theta, phi));
  (*combi)->SetName(Form("%scombi", name));
  (*combi)->RegisterYourself();
}
// reference src/AGeoUtil.cxx:84-125: an Xtru prism from the corners of its top face and the bottom corner below the first one
inline void MakeXtruFromPoints(const char* name, const std::vector<TVector3>& vecs, TGeoXtru** xtru, TGeoCombiTrans** combi) {
  std::size_t nvert = vecs.size() - 1;
  TVector3 normal = vecs[0] - vecs[nvert];
  Double_t dZ = normal.Mag() / 2., theta = normal.Theta(), phi = normal.Phi();
  std::vector<Double_t> x(nvert), y(nvert);
  for (std::size_t i = 0; i < nvert; ++i) {
    TVector3 v = vecs[i] - vecs[0];
    v.RotateZ(-phi - TMath::Pi() / 2.);
    v.RotateX(-theta);
    x[i] = v.X();
    y[i] = v.Y();
  }
  *xtru = new TGeoXtru(2);
  (*xtru)->SetName(name);
  (*xtru)->DefinePolygon((Int_t)nvert, x.data(), y.data());
  (*xtru)->DefineSection(0, -dZ);
  (*xtru)->DefineSection(1, +dZ);
  TVector3 shift = vecs[0] - normal * .5;
  TGeoTranslation tr(shift.X(), shift.Y(), shift.Z());
  *combi = new TGeoCombiTrans(tr, AxisRotation(theta, phi));
  (*combi)->SetName(Form("%scombi", name));
  (*combi)->RegisterYourself();
}
// reference src/AGeoUtil.cxx:198-308: radius and centre of the circle containing `fraction` of the histogram (D80 for
// fraction = 0.8).  Runs on the GPU (rbg_containment_radius_host: one block, the reference's search step for step).
inline void ContainmentRadius(TH2* h2, Double_t fraction, Double_t& r, Double_t& x, Double_t& y, Int_t device = 0) {
  const Int_t nx = h2->GetNbinsX(), ny = h2->GetNbinsY();
  std::vector<Double_t> bins((size_t)nx * ny);
  for (Int_t j = 1; j <= ny; j++)
    for (Int_t i = 1; i <= nx; i++) bins[(i - 1) + (size_t)nx * (j - 1)] = h2->GetBinContent(i, j);
  Double_t stats[5], out[3];
  h2->GetStats5(stats);
  if (rbg_containment_radius_host(bins.data(), nx, h2->GetXaxis()->GetXmin(), h2->GetXaxis()->GetXmax(), ny, h2->GetYaxis()->GetXmin(),
                                  h2->GetYaxis()->GetXmax(), stats, fraction, out, device) != RBG_OK)
    throw std::runtime_error(std::string("AGeoUtil::ContainmentRadius: ") + rbg_last_error());
  r = out[0]; x = out[1]; y = out[2];
}
}  // namespace AGeoUtil

// ============================================================================ multilayer (description; TMM runs on device)
class AOpticsManager;
struct ASceneExport;
// reference include/AMultilayer.h:23-298, src/AMultilayer.cxx:135-238
class AMultilayer : public TObject {
 public:
  enum EPolarization { kS, kP };

 private:
  std::vector<std::shared_ptr<ARefractiveIndex>> fRefractiveIndexList;  // [0]=top ... [n-1]=bottom
  std::vector<Double_t> fThicknessList;
  std::vector<Bool_t> fCoherentList;  // per layer; the two semi-infinite ends are incoherent (src/AMultilayer.cxx:104-114)
  std::shared_ptr<TH2D> fPreCalculatedReflectanceMixed, fPreCalculatedTransmittanceMixed;
  void DeviceTMM(Int_t n, const Double_t* th, const Double_t* lam, Double_t* R, Double_t* T) const;  // defined below
  // mode 0 = coherent (optionally reversed), 1 = incoherent; pol 0 = s, 1 = p, 2 = mixed
  void DeviceTMMGeneral(Int_t mode, Int_t pol, Bool_t reverse, Int_t n, const std::complex<Double_t>* th, const Double_t* lam, Double_t* R, Double_t* T) const;

 public:
  AMultilayer(std::shared_ptr<ARefractiveIndex> top, std::shared_ptr<ARefractiveIndex> bottom) {
    const Double_t inf = std::numeric_limits<Double_t>::infinity();
    fRefractiveIndexList = {top, bottom};
    fThicknessList = {inf, inf};
    fCoherentList = {kFALSE, kFALSE};
  }
  void AddLayer(std::shared_ptr<ARefractiveIndex> idx, Double_t thickness, Bool_t coherent = kTRUE) { RbGeomTouch();
    fRefractiveIndexList.insert(fRefractiveIndexList.begin() + 1, idx);
    fThicknessList.insert(fThicknessList.begin() + 1, thickness);
    fCoherentList.insert(fCoherentList.begin() + 1, coherent);
  }
  void InsertLayer(std::shared_ptr<ARefractiveIndex> idx, Double_t thickness, Bool_t coherent = kTRUE) { RbGeomTouch();
    fRefractiveIndexList.insert(fRefractiveIndexList.end() - 1, idx);
    fThicknessList.insert(fThicknessList.end() - 1, thickness);
    fCoherentList.insert(fCoherentList.end() - 1, coherent);
  }
  const std::vector<Bool_t>& GetCoherentList() const { return fCoherentList; }
  // reference src/AMultilayer.cxx:240-481 (one polarisation, complex angle, optionally the reversed stack) — on the GPU
  void CoherentTMM(EPolarization pol, std::complex<Double_t> th_0, Double_t lam_vac, Double_t& reflectance, Double_t& transmittance, Bool_t reverse = kFALSE) const {
    DeviceTMMGeneral(0, pol == kS ? 0 : 1, reverse, 1, &th_0, &lam_vac, &reflectance, &transmittance);
  }
  void CoherentTMMP(std::complex<Double_t> th_0, Double_t lam_vac, Double_t& r, Double_t& t) const { CoherentTMM(kP, th_0, lam_vac, r, t); }
  void CoherentTMMS(std::complex<Double_t> th_0, Double_t lam_vac, Double_t& r, Double_t& t) const { CoherentTMM(kS, th_0, lam_vac, r, t); }
  void CoherentTMMMixed(std::complex<Double_t> th_0, Double_t lam_vac, Double_t& reflectance, Double_t& transmittance) const {
    DeviceTMMGeneral(0, 2, kFALSE, 1, &th_0, &lam_vac, &reflectance, &transmittance);
  }
  // reference src/AMultilayer.cxx:484-731 (tmm.inc_tmm): layers inserted with coherent = kFALSE exchange power only
  void IncoherentTMM(EPolarization pol, std::complex<Double_t> th_0, Double_t lam_vac, Double_t& reflectance, Double_t& transmittance) const {
    DeviceTMMGeneral(1, pol == kS ? 0 : 1, kFALSE, 1, &th_0, &lam_vac, &reflectance, &transmittance);
  }
  void IncoherentTMMP(std::complex<Double_t> th_0, Double_t lam_vac, Double_t& r, Double_t& t) const { IncoherentTMM(kP, th_0, lam_vac, r, t); }
  void IncoherentTMMS(std::complex<Double_t> th_0, Double_t lam_vac, Double_t& r, Double_t& t) const { IncoherentTMM(kS, th_0, lam_vac, r, t); }
  void IncoherentTMMMixed(std::complex<Double_t> th_0, Double_t lam_vac, Double_t& reflectance, Double_t& transmittance) const {
    DeviceTMMGeneral(1, 2, kFALSE, 1, &th_0, &lam_vac, &reflectance, &transmittance);
  }
  // Snell's law with complex indices (include/AMultilayer.h Snell): angle in medium 2
  static std::complex<Double_t> Snell(std::complex<Double_t> n_1, std::complex<Double_t> n_2, std::complex<Double_t> th_1) {
    return std::asin(n_1 * std::sin(th_1) / n_2);
  }
  void ChangeThickness(std::size_t i, Double_t thickness) { RbGeomTouch();
    if (i < 1 || i > fThicknessList.size() - 2) Error("ChangeThickness", "Cannot change the thickness of the %luth layer", (unsigned long)i);
    else fThicknessList[i] = thickness;
  }
  Double_t GetThickness(std::size_t i) const { return fThicknessList.at(i); }
  std::size_t GetNLayers() const { return fThicknessList.size(); }
  const std::vector<std::shared_ptr<ARefractiveIndex>>& GetIndexList() const { return fRefractiveIndexList; }
  const std::vector<Double_t>& GetThicknessList() const { return fThicknessList; }
  std::shared_ptr<TH2D> GetPrecalculatedReflectanceMixed() const { return fPreCalculatedReflectanceMixed; }
  std::shared_ptr<TH2D> GetPrecalculatedTransmittanceMixed() const { return fPreCalculatedTransmittanceMixed; }
  // unpolarised R,T at (theta, lambda); evaluated on the GPU through rbg_tmm (no host TMM exists)
  void CoherentTMMMixed(Double_t th_0, Double_t lam_vac, Double_t& reflectance, Double_t& transmittance) const {
    if (fPreCalculatedReflectanceMixed && fPreCalculatedTransmittanceMixed) {
      reflectance = fPreCalculatedReflectanceMixed->Interpolate(lam_vac, th_0);
      transmittance = fPreCalculatedTransmittanceMixed->Interpolate(lam_vac, th_0);
      return;
    }
    DeviceTMM(1, &th_0, &lam_vac, &reflectance, &transmittance);
  }
  void CoherentTMMMixed(const std::vector<Double_t>& th_0, Double_t lam_vac, std::vector<Double_t>& reflectance, std::vector<Double_t>& transmittance) const {
    std::vector<Double_t> lam(th_0.size(), lam_vac);
    reflectance.resize(th_0.size());
    transmittance.resize(th_0.size());
    if (!th_0.empty()) DeviceTMM((Int_t)th_0.size(), th_0.data(), lam.data(), reflectance.data(), transmittance.data());
  }
  void CoherentTMMMixed(Double_t th_0, const std::vector<Double_t>& lam_vac, std::vector<Double_t>& reflectance, std::vector<Double_t>& transmittance) const {
    std::vector<Double_t> th(lam_vac.size(), th_0);
    reflectance.resize(lam_vac.size());
    transmittance.resize(lam_vac.size());
    if (!lam_vac.empty()) DeviceTMM((Int_t)lam_vac.size(), th.data(), lam_vac.data(), reflectance.data(), transmittance.data());
  }
  // reference include/AMultilayer.h:243-262 — λ×θ table at bin centres, one device launch
  void PreCalculateCoherentTMM(Int_t lam_nbins, Double_t lam_min, Double_t lam_max, Int_t th_nbins, Double_t th_min, Double_t th_max) { RbGeomTouch();
    auto R = std::make_shared<TH2D>("", "", lam_nbins, lam_min, lam_max, th_nbins, th_min, th_max);
    auto T = std::make_shared<TH2D>("", "", lam_nbins, lam_min, lam_max, th_nbins, th_min, th_max);
    std::vector<Double_t> th, lam;
    for (Int_t j = 1; j <= th_nbins; ++j)
      for (Int_t i = 1; i <= lam_nbins; ++i) {
        th.push_back(R->GetYaxis()->GetBinCenter(j));
        lam.push_back(R->GetXaxis()->GetBinCenter(i));
      }
    std::vector<Double_t> r(th.size()), t(th.size());
    fPreCalculatedReflectanceMixed.reset();
    fPreCalculatedTransmittanceMixed.reset();
    DeviceTMM((Int_t)th.size(), th.data(), lam.data(), r.data(), t.data());
    size_t k = 0;
    for (Int_t j = 1; j <= th_nbins; ++j)
      for (Int_t i = 1; i <= lam_nbins; ++i, ++k) {
        R->SetBinContent(i, j, r[k]);
        T->SetBinContent(i, j, t[k]);
      }
    fPreCalculatedReflectanceMixed = R;
    fPreCalculatedTransmittanceMixed = T;
  }
  // reference include/AMultilayer.h:263-282: the same table filled by IncoherentTMMMixed
  void PreCalculateIncoherentTMM(Int_t lam_nbins, Double_t lam_min, Double_t lam_max, Int_t th_nbins, Double_t th_min, Double_t th_max) { RbGeomTouch();
    auto R = std::make_shared<TH2D>("", "", lam_nbins, lam_min, lam_max, th_nbins, th_min, th_max);
    auto T = std::make_shared<TH2D>("", "", lam_nbins, lam_min, lam_max, th_nbins, th_min, th_max);
    std::vector<std::complex<Double_t>> th;
    std::vector<Double_t> lam;
    for (Int_t j = 1; j <= th_nbins; ++j)
      for (Int_t i = 1; i <= lam_nbins; ++i) {
        th.push_back(R->GetYaxis()->GetBinCenter(j));
        lam.push_back(R->GetXaxis()->GetBinCenter(i));
      }
    std::vector<Double_t> r(th.size()), t(th.size());
    DeviceTMMGeneral(1, 2, kFALSE, (Int_t)th.size(), th.data(), lam.data(), r.data(), t.data());
    size_t k = 0;
    for (Int_t j = 1; j <= th_nbins; ++j)
      for (Int_t i = 1; i <= lam_nbins; ++i, ++k) {
        R->SetBinContent(i, j, r[k]);
        T->SetBinContent(i, j, t[k]);
      }
    fPreCalculatedReflectanceMixed = R;
    fPreCalculatedTransmittanceMixed = T;
  }
  void SetNthreads(std::size_t) {}
};

// ============================================================================ optical components
class ABorderSurfaceCondition;
// reference include/AOpticalComponent.h:22-43, src/AOpticalComponent.cxx:24-65
class AOpticalComponent : public TGeoVolume {
  std::vector<ABorderSurfaceCondition*> fBorders;

 public:
  AOpticalComponent() {}
  AOpticalComponent(const char* name, const TGeoShape* shape, const TGeoMedium* med = nullptr) : TGeoVolume(name, shape, med) {}
  virtual Int_t OpticalType() const { return RBG_OPT; }
  void AddBorderSurfaceCondition(ABorderSurfaceCondition* c) { RbGeomTouch(); fBorders.push_back(c); }
  ABorderSurfaceCondition* FindBorderSurfaceCondition(AOpticalComponent* component2);
  const std::vector<ABorderSurfaceCondition*>& GetBorders() const { return fBorders; }
};

// reference include/ABorderSurfaceCondition.h:24-48, src/ABorderSurfaceCondition.cxx:20-41
class ABorderSurfaceCondition : public TObject {
  AOpticalComponent* fComponent[2];
  Double_t fSigma = 0;
  std::shared_ptr<AMultilayer> fMultilayer;
  Bool_t fLambertian = false;

 public:
  ABorderSurfaceCondition(AOpticalComponent* component1, AOpticalComponent* component2) {
    fComponent[0] = component1;
    fComponent[1] = component2;
    if (component1) component1->AddBorderSurfaceCondition(this);
  }
  AOpticalComponent* GetComponent1() { return fComponent[0]; }
  AOpticalComponent* GetComponent2() { return fComponent[1]; }
  Double_t GetGaussianRoughness() const { return fSigma; }
  void SetGaussianRoughness(Double_t sigma) { RbGeomTouch(); fSigma = std::fabs(sigma); }
  void SetMultilayer(std::shared_ptr<AMultilayer> layer) { RbGeomTouch(); fMultilayer = layer; }
  std::shared_ptr<AMultilayer> GetMultilayer() const { return fMultilayer; }
  Bool_t IsLambertian() const { return fLambertian; }
  void EnableLambertian(Bool_t mode) { RbGeomTouch(); fLambertian = mode; }
};
inline ABorderSurfaceCondition* AOpticalComponent::FindBorderSurfaceCondition(AOpticalComponent* component2) {
  for (auto* b : fBorders)
    if (b->GetComponent2() == component2) return b;
  return nullptr;
}

// reference include/ALens.h:22-41, src/ALens.cxx:36-60
class ALens : public AOpticalComponent {
  std::shared_ptr<ARefractiveIndex> fIndex;

 public:
  ALens(const char* name, const TGeoShape* shape, const TGeoMedium* med = nullptr) : AOpticalComponent(name, shape, med) {}
  Int_t OpticalType() const override { return RBG_LENS; }
  Double_t GetAbsorptionLength(Double_t lambda) const { return fIndex ? fIndex->GetAbsorptionLength(lambda) : std::numeric_limits<Double_t>::infinity(); }
  Double_t GetExtinctionCoefficient(Double_t lambda) const { return fIndex ? fIndex->GetExtinctionCoefficient(lambda) : 0; }
  Double_t GetRefractiveIndex(Double_t lambda) const { return fIndex ? fIndex->GetRefractiveIndex(lambda) : 1.; }
  void SetRefractiveIndex(std::shared_ptr<ARefractiveIndex> index) { RbGeomTouch(); fIndex = index; }
  std::shared_ptr<ARefractiveIndex> GetIndex() const { return fIndex; }
};

// reference include/AMirror.h:22-47, src/AMirror.cxx:24-60
class AMirror : public AOpticalComponent {
  Double_t fReflectance = 1.0;
  std::shared_ptr<TGraph> fReflectance1D;
  std::shared_ptr<TGraph2D> fReflectance2D;
  std::shared_ptr<TH2> fReflectanceTH2;

 public:
  AMirror(const char* name, const TGeoShape* shape, const TGeoMedium* med = nullptr) : AOpticalComponent(name, shape, med) {}
  Int_t OpticalType() const override { return RBG_MIRROR; }
  void SetReflectance(Double_t ref) { RbGeomTouch(); fReflectance = ref; }
  void SetReflectance(std::shared_ptr<TGraph> ref) { RbGeomTouch(); fReflectance1D = ref; }
  void SetReflectance(std::shared_ptr<TGraph2D> ref) { RbGeomTouch(); fReflectance2D = ref; }
  void SetReflectance(std::shared_ptr<TH2> ref) { RbGeomTouch(); fReflectanceTH2 = ref; }
  // src/AMirror.cxx:39-60: priority TGraph2D > TH2 > TGraph > constant, clamped to [0, 1]
  Double_t GetReflectance(Double_t lambda, Double_t angle) const {
    Double_t ret = fReflectance;
    if (fReflectance2D) ret = fReflectance2D->Interpolate(lambda, angle);
    else if (fReflectanceTH2) ret = fReflectanceTH2->Interpolate(lambda, angle);
    else if (fReflectance1D) ret = fReflectance1D->Eval(lambda);
    return ret > 1 ? 1 : (ret < 0 ? 0 : ret);
  }
  Double_t GetConstantReflectance() const { return fReflectance; }
  std::shared_ptr<TGraph> GetReflectance1D() const { return fReflectance1D; }
  std::shared_ptr<TGraph2D> GetReflectance2D() const { return fReflectance2D; }
  std::shared_ptr<TH2> GetReflectanceTH2() const { return fReflectanceTH2; }
};

// reference include/AFocalSurface.h:22-45, src/AFocalSurface.cxx:35-52 (QE graphs are raw, non-owned)
class AFocalSurface : public AOpticalComponent {
  TGraph* fQuantumEfficiencyLambda = nullptr;
  TGraph* fQuantumEfficiencyAngle = nullptr;

 public:
  AFocalSurface(const char* name, const TGeoShape* shape, const TGeoMedium* med = nullptr) : AOpticalComponent(name, shape, med) {}
  Int_t OpticalType() const override { return RBG_FOCUS; }
  Bool_t HasQEAngle() const { return fQuantumEfficiencyAngle != nullptr; }
  void SetQuantumEfficiency(TGraph* qe) { RbGeomTouch(); fQuantumEfficiencyLambda = qe; }
  void SetQuantumEfficiencyAngle(TGraph* qe) { RbGeomTouch(); fQuantumEfficiencyAngle = qe; }
  TGraph* GetQELambdaGraph() const { return fQuantumEfficiencyLambda; }
  TGraph* GetQEAngleGraph() const { return fQuantumEfficiencyAngle; }
  Double_t GetQuantumEfficiency(Double_t lambda) const { return fQuantumEfficiencyLambda ? fQuantumEfficiencyLambda->Eval(lambda) : 1.; }
  Double_t GetQuantumEfficiency(Double_t lambda, Double_t angle) const {
    Double_t qe = GetQuantumEfficiency(lambda);
    if (HasQEAngle()) qe *= fQuantumEfficiencyAngle->Eval(angle);
    return qe;
  }
};

// reference include/AObscuration.h, src/AObscuration.cxx
class AObscuration : public AOpticalComponent {
 public:
  AObscuration(const char* name, const TGeoShape* shape, const TGeoMedium* med = nullptr) : AOpticalComponent(name, shape, med) {}
  Int_t OpticalType() const override { return RBG_OBS; }
};

// ============================================================================ rays
class TPolyLine3D;
// reference include/ARay.h:24-68, src/ARay.cxx:28-36,66-71,210-223.  Only the first and the last
// track point are kept; GetNpoints() is the true count.
class ARay : public TObject {
 public:
  enum { kRun, kStop, kExit, kFocus, kSuspend, kAbsorb };

 private:
  Double_t fFirst[4], fLast[4];
  Int_t fNpoints = 1;
  Double_t fLambda;
  TVector3 fDirection;
  Int_t fStatus = kRun;
  Int_t fId;
  TObjArray fNodeHistory;
  TNamed fLastNode;
  std::vector<Double_t> fHist;    // recorded polyline, 4 per point (x, y, z, t), point 0 = start; empty = not recorded
  std::vector<TNamed> fNodeObjs;  // node-history entries of the recorded points ("<volume>_<copyNo>")

 public:
  ARay(Int_t id, Double_t lambda, Double_t x, Double_t y, Double_t z, Double_t t, Double_t nx, Double_t ny, Double_t nz) : fLambda(lambda), fId(id) {
    fFirst[0] = fLast[0] = x; fFirst[1] = fLast[1] = y; fFirst[2] = fLast[2] = z; fFirst[3] = fLast[3] = t;
    SetDirection(nx, ny, nz);
  }
  void Absorb() { fStatus = kAbsorb; }
  void Exit() { fStatus = kExit; }
  void Focus() { fStatus = kFocus; }
  void Stop() { fStatus = kStop; }
  void Suspend() { fStatus = kSuspend; }
  Bool_t IsAbsorbed() const { return fStatus == kAbsorb; }
  Bool_t IsExited() const { return fStatus == kExit; }
  Bool_t IsFocused() const { return fStatus == kFocus; }
  Bool_t IsRunning() const { return fStatus == kRun; }
  Bool_t IsStopped() const { return fStatus == kStop; }
  Bool_t IsSuspended() const { return fStatus == kSuspend; }
  Int_t GetStatus() const { return fStatus; }
  Int_t GetId() const { return fId; }
  void GetDirection(Double_t* v) const { fDirection.GetXYZ(v); }
  void GetLastPoint(Double_t* x) const { memcpy(x, fLast, sizeof(fLast)); }
  const Double_t* GetFirstPoint() const { return fFirst; }
  const Double_t* GetLastPoint() const { return fLast; }
  // TGeoTrack::GetPoint: vertex i of the polyline.  Intermediate vertices exist when the trace recorded a history
  // (AOpticsManager::SetHistoryDepth; on by default for small batches); otherwise only the first and last are kept.
  const Double_t* GetPoint(Int_t i) const {
    if (i >= 0 && (size_t)(4 * i + 3) < fHist.size()) return &fHist[4 * i];
    return i == fNpoints - 1 ? fLast : fFirst;
  }
  Int_t GetPoint(Int_t i, Double_t& x, Double_t& y, Double_t& z, Double_t& t) const {
    if (i < 0 || i >= fNpoints) return -1;
    const Double_t* p = GetPoint(i);
    x = p[0]; y = p[1]; z = p[2]; t = p[3];
    return i;
  }
  // TVirtualGeoTrack-style lookup by time of flight (tutorials/AbsLengthTest.C:54 calls GetPoint(0, p0)): the point of the
  // polyline at time `tof`, linearly interpolated between the recorded vertices; clamps to the first / last point
  Int_t GetPoint(Double_t tof, Double_t* point, Int_t istart = 0) const {
    Int_t n = GetNrecorded();
    const Double_t* first = n > 0 ? &fHist[0] : fFirst;
    const Double_t* last = n > 0 && n == fNpoints ? &fHist[4 * (n - 1)] : fLast;
    if (!(tof > first[3]) || fNpoints < 2) { memcpy(point, first, 4 * sizeof(Double_t)); return 0; }
    if (!(tof < last[3])) { memcpy(point, last, 4 * sizeof(Double_t)); return fNpoints - 1; }
    for (Int_t i = std::max(istart, 0); i + 1 < n; i++) {
      const Double_t *a = &fHist[4 * i], *b = &fHist[4 * (i + 1)];
      if (tof >= a[3] && tof <= b[3] && b[3] > a[3]) {
        Double_t u = (tof - a[3]) / (b[3] - a[3]);
        for (int k = 0; k < 3; k++) point[k] = a[k] + u * (b[k] - a[k]);
        point[3] = tof;
        return i;
      }
    }
    Double_t u = (tof - first[3]) / (last[3] - first[3]);
    for (int k = 0; k < 3; k++) point[k] = first[k] + u * (last[k] - first[k]);
    point[3] = tof;
    return 0;
  }
  Int_t GetNrecorded() const { return (Int_t)(fHist.size() / 4); }
  Int_t GetNnodesRecorded() const { return (Int_t)fNodeObjs.size(); }  // node-history slots, a trailing null (world exit) included
  Int_t GetNpoints() const { return fNpoints; }
  Double_t GetLambda() const { return fLambda; }
  void SetLambda(Double_t l) { fLambda = l; }
  void SetDirection(Double_t dx, Double_t dy, Double_t dz) {
    Double_t mag = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (mag > 0) fDirection.SetXYZ(dx / mag, dy / mag, dz / mag);
  }
  void SetDirection(Double_t* d) { SetDirection(d[0], d[1], d[2]); }
  void AddPoint(Double_t x, Double_t y, Double_t z, Double_t t) {
    fLast[0] = x; fLast[1] = y; fLast[2] = z; fLast[3] = t;
    fNpoints++;
  }
  const TObjArray* GetNodeHistory() const { return &fNodeHistory; }
  const char* GetLastNodeName() const { return fLastNode.GetName(); }
  // src/ARay.cxx:42-63: first node-history entry whose name starts with `name` (entry n belongs to point n + 1).
  // Without a recorded history the list holds the last node only.
  TObject* FindNodeStartWith(const char* name) const {
    Int_t i = FindNodeNumberStartWith(name);
    return i < 0 ? nullptr : fNodeHistory.At(i);
  }
  Int_t FindNodeNumberStartWith(const char* name) const {
    for (Int_t i = 0; i <= fNodeHistory.GetLast(); i++) {
      TObject* node = fNodeHistory.At(i);
      if (node && strncmp(node->GetName(), name, strlen(name)) == 0) return i;
    }
    return -1;
  }
  void SetLineWidth(Int_t) {}
  void SetLineColor(Int_t) {}
  TPolyLine3D* MakePolyLine3D() const;
  // used by the tracer to write results back
  void SetTraced(const Double_t last[4], const Double_t dir[3], Int_t status, Int_t npoints, const char* last_node) {
    memcpy(fLast, last, sizeof(fLast));
    fDirection.SetXYZ(dir[0], dir[1], dir[2]);
    fStatus = status;
    fNpoints = npoints;
    fNodeHistory.Clear();
    fHist.clear();
    fNodeObjs.clear();
    if (last_node) {
      fLastNode.SetName(last_node);
      fNodeHistory.Add(&fLastNode);
    }
  }
  // recorded polyline of the last trace: npts points (x, y, z, t each) and the node entered with points 1..npts-1
  void SetHistory(const Double_t* pts, Int_t npts, const int32_t* nodes, const std::vector<std::string>* names) {
    fHist.assign(pts, pts + 4 * (size_t)npts);
    fNodeHistory.Clear();
    fNodeObjs.clear();
    fNodeObjs.resize(npts > 1 ? npts - 1 : 0);
    for (Int_t k = 1; k < npts; k++) {
      int32_t id = nodes[k];
      if (names && id >= 0 && id < (int32_t)names->size()) {
        fNodeObjs[k - 1].SetName((*names)[id].c_str());
        fNodeHistory.Add(&fNodeObjs[k - 1]);
      } else fNodeHistory.Add(nullptr);  // left the world: the reference adds a null node
    }
  }
};

// display stubs (drawing is out of scope)
class TPolyLine3D : public TObject {
 public:
  void SetLineColor(Int_t) {}
  void SetLineWidth(Int_t) {}
};
inline TPolyLine3D* ARay::MakePolyLine3D() const { return new TPolyLine3D; }
// display stub of the OpenGL viewer (tutorials/AshraOptics.C:194-196)
class TGLViewer : public TObject {
 public:
  enum ECameraType { kCameraPerspXOZ, kCameraPerspYOZ, kCameraPerspXOY, kCameraOrthoXOY, kCameraOrthoXOZ, kCameraOrthoZOY };
  void SetPerspectiveCamera(ECameraType, Double_t, Double_t, Double_t*, Double_t, Double_t) {}
  void SetPerspectiveCamera(ECameraType, Double_t, Double_t, int, Double_t, Double_t) {}
};
// TNtuple of doubles kept in memory (tutorials/AshraOptics.C:143,179): rows of the variables named in the constructor
class TNtuple : public TNamed {
  Int_t fNvar = 0;
  std::vector<Float_t> fRows;

 public:
  TNtuple(const char* name, const char* title, const char* varlist) : TNamed(name, title) {
    fNvar = 1;
    for (const char* c = varlist; c && *c; c++) fNvar += *c == ':';
  }
  Int_t Fill(Float_t x0, Float_t x1 = 0, Float_t x2 = 0, Float_t x3 = 0, Float_t x4 = 0, Float_t x5 = 0) {
    const Float_t v[6] = {x0, x1, x2, x3, x4, x5};
    for (Int_t i = 0; i < fNvar && i < 6; i++) fRows.push_back(v[i]);
    return fNvar * (Int_t)sizeof(Float_t);
  }
  Long64_t GetEntries() const { return fNvar ? (Long64_t)fRows.size() / fNvar : 0; }
  Int_t GetNvar() const { return fNvar; }
  const Float_t* GetRow(Long64_t i) const { return fRows.data() + i * fNvar; }
};
class TCanvas : public TNamed {
 public:
  TCanvas(const char* n = "", const char* t = "", Int_t = 0, Int_t = 0) : TNamed(n, t) {}
  TGLViewer* GetViewer3D(const char* = "") {
    static TGLViewer viewer;
    return &viewer;
  }
  void Divide(Int_t, Int_t, Double_t = 0, Double_t = 0) {}
  TCanvas* cd(Int_t = 0) { return this; }
  void SetGridx() {}
  void SetGridy() {}
  void SetLogx() {}
  void SetLogy() {}
  void SetLogz() {}
  void Update() {}
  TH1* DrawFrame(Double_t, Double_t, Double_t, Double_t, const char* = "") {  // display stub: an empty frame histogram
    static TH1D frame("frame", "", 1, 0, 1);
    return &frame;
  }
};
struct TStyle {
  void SetOptStat(Int_t = 1) {}
  void SetOptFit(Int_t = 1) {}
  void SetPalette(Int_t = 0) {}
};
inline TStyle*& gStyleRef() { static TStyle* p = new TStyle; return p; }
#define gStyle (gStyleRef())
class TLegend : public TObject {
 public:
  TLegend(Double_t, Double_t, Double_t, Double_t) {}
  void SetFillStyle(Int_t) {}
  void SetTextFont(Int_t) {}
  void SetTextSize(Double_t) {}
  void SetBorderSize(Int_t) {}
  void AddEntry(TObject*, const char*, const char* = "") {}
};
inline TCanvas*& gPadRef() { static TCanvas* p = new TCanvas; return p; }
#define gPad (gPadRef())

// reference include/ARayArray.h:22-47, src/ARayArray.cxx:21-73.  SoA-backed: one table of all rays in
// insertion order; the six status buckets are lazily materialised stable views.
class ARayArray : public TObject {
 public:
  struct Table {
    std::vector<Double_t> x0, y0, z0, t0, x, y, z, t, dx, dy, dz, lambda;
    std::vector<int32_t> status, npoints, last_node;
    std::vector<ARay*> obj;  // lazily created ARay views (owned)
    // optional polyline record, packed: ray i owns points hoff[i] .. hoff[i+1] of hpts (x,y,z,t per point) and hnode
    // (hoff is empty while no trace has recorded anything)
    std::vector<int64_t> hoff;
    std::vector<Double_t> hpts;
    std::vector<int32_t> hnode;
    size_t size() const { return x.size(); }
    bool HasHistory() const { return !hoff.empty(); }
    void EnsureHistoryOffsets() {  // rays appended since the last trace own empty records
      if (hoff.empty()) hoff.push_back(0);
      while (hoff.size() < size() + 1) hoff.push_back(hoff.back());
    }
    int32_t HistCount(size_t i) const { return hoff.size() > i + 1 ? (int32_t)(hoff[i + 1] - hoff[i]) : 0; }
  };

 private:
  Table fT;
  TObjArray fBucket[6];
  Bool_t fViewsValid = kFALSE;
  std::shared_ptr<std::vector<std::string>> fNodeNames;

  void Materialise() {
    if (fViewsValid) return;
    for (auto& b : fBucket) b.Clear();
    for (size_t i = 0; i < fT.size(); i++) {
      Double_t last[4] = {fT.x[i], fT.y[i], fT.z[i], fT.t[i]}, dir[3] = {fT.dx[i], fT.dy[i], fT.dz[i]};
      if (!fT.obj[i]) fT.obj[i] = new ARay((Int_t)i, fT.lambda[i], fT.x0[i], fT.y0[i], fT.z0[i], fT.t0[i], dir[0], dir[1], dir[2]);
      const char* nn = nullptr;
      if (fNodeNames && fT.last_node[i] >= 0 && fT.last_node[i] < (int32_t)fNodeNames->size()) nn = (*fNodeNames)[fT.last_node[i]].c_str();
      fT.obj[i]->SetTraced(last, dir, fT.status[i], fT.npoints[i], nn);
      if (fT.HistCount(i) > 0) fT.obj[i]->SetHistory(&fT.hpts[4 * (size_t)fT.hoff[i]], fT.HistCount(i), &fT.hnode[(size_t)fT.hoff[i]], fNodeNames.get());
      fBucket[fT.status[i]].Add(fT.obj[i]);
    }
    fViewsValid = kTRUE;
  }

 public:
  ARayArray() {}
  ~ARayArray() override {
    for (auto* o : fT.obj) delete o;
  }
  // appends a ray (takes ownership of `ray`, as the reference's owning TObjArrays do)
  virtual void Add(ARay* ray) {
    if (!ray) return;
    Double_t p[4], d[3];
    ray->GetLastPoint(p);
    ray->GetDirection(d);
    const Double_t* f = ray->GetFirstPoint();
    AddRaw(f[0], f[1], f[2], f[3], d[0], d[1], d[2], ray->GetLambda());
    size_t i = fT.size() - 1;
    fT.x[i] = p[0]; fT.y[i] = p[1]; fT.z[i] = p[2]; fT.t[i] = p[3];
    fT.status[i] = ray->GetStatus();
    fT.npoints[i] = ray->GetNpoints();
    fT.obj[i] = ray;
  }
  // SoA append of a running ray; direction is normalised like ARay's constructor (src/ARay.cxx:28-36)
  void AddRaw(Double_t x, Double_t y, Double_t z, Double_t t, Double_t dx, Double_t dy, Double_t dz, Double_t lambda) {
    Double_t mag = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (mag > 0) { dx /= mag; dy /= mag; dz /= mag; }
    fT.x0.push_back(x); fT.y0.push_back(y); fT.z0.push_back(z); fT.t0.push_back(t);
    fT.x.push_back(x); fT.y.push_back(y); fT.z.push_back(z); fT.t.push_back(t);
    fT.dx.push_back(dx); fT.dy.push_back(dy); fT.dz.push_back(dz); fT.lambda.push_back(lambda);
    fT.status.push_back(RBG_RUN); fT.npoints.push_back(1); fT.last_node.push_back(-1);
    fT.obj.push_back(nullptr);
    if (fT.HasHistory()) fT.EnsureHistoryOffsets();
    fViewsValid = kFALSE;
  }
  void Reserve(size_t n) {
    for (auto* v : {&fT.x0, &fT.y0, &fT.z0, &fT.t0, &fT.x, &fT.y, &fT.z, &fT.t, &fT.dx, &fT.dy, &fT.dz, &fT.lambda}) v->reserve(n);
    fT.status.reserve(n); fT.npoints.reserve(n); fT.last_node.reserve(n); fT.obj.reserve(n);
  }
  // appends all rays of `array` bucket by bucket in this order: absorbed, exited, focused, running,
  // stopped, suspended (src/ARayArray.cxx:59-73); `array` is emptied
  virtual void Merge(ARayArray* array) {
    if (!array) return;
    Table& o = array->fT;
    static const int order[6] = {RBG_ABSORB, RBG_EXIT, RBG_FOCUSED, RBG_RUN, RBG_STOP, RBG_SUSPEND};
    const bool hist = fT.HasHistory() || o.HasHistory();
    if (hist) { fT.EnsureHistoryOffsets(); o.EnsureHistoryOffsets(); }
    for (int s : order)
      for (size_t i = 0; i < o.size(); i++) {
        if (o.status[i] != s) continue;
        if (hist) {
          fT.hpts.insert(fT.hpts.end(), o.hpts.begin() + 4 * o.hoff[i], o.hpts.begin() + 4 * o.hoff[i + 1]);
          fT.hnode.insert(fT.hnode.end(), o.hnode.begin() + o.hoff[i], o.hnode.begin() + o.hoff[i + 1]);
          fT.hoff.push_back(fT.hoff.back() + (o.hoff[i + 1] - o.hoff[i]));
        }
        fT.x0.push_back(o.x0[i]); fT.y0.push_back(o.y0[i]); fT.z0.push_back(o.z0[i]); fT.t0.push_back(o.t0[i]);
        fT.x.push_back(o.x[i]); fT.y.push_back(o.y[i]); fT.z.push_back(o.z[i]); fT.t.push_back(o.t[i]);
        fT.dx.push_back(o.dx[i]); fT.dy.push_back(o.dy[i]); fT.dz.push_back(o.dz[i]); fT.lambda.push_back(o.lambda[i]);
        fT.status.push_back(o.status[i]); fT.npoints.push_back(o.npoints[i]); fT.last_node.push_back(o.last_node[i]);
        fT.obj.push_back(o.obj[i]);
      }
    if (!fNodeNames) fNodeNames = array->fNodeNames;
    o = Table();
    array->fViewsValid = kFALSE;
    fViewsValid = kFALSE;
  }
  TObjArray* GetAbsorbed() { Materialise(); return &fBucket[RBG_ABSORB]; }
  TObjArray* GetExited() { Materialise(); return &fBucket[RBG_EXIT]; }
  TObjArray* GetFocused() { Materialise(); return &fBucket[RBG_FOCUSED]; }
  TObjArray* GetRunning() { Materialise(); return &fBucket[RBG_RUN]; }
  TObjArray* GetStopped() { Materialise(); return &fBucket[RBG_STOP]; }
  TObjArray* GetSuspended() { Materialise(); return &fBucket[RBG_SUSPEND]; }
  // SoA access for large arrays (extension; avoids materialising ARay objects)
  Table& GetTable() { fViewsValid = kFALSE; return fT; }
  const Table& GetTable() const { return fT; }
  Long64_t GetN() const { return (Long64_t)fT.size(); }
  Long64_t Count(Int_t status) const { return std::count(fT.status.begin(), fT.status.end(), status); }
  void SetNodeNames(std::shared_ptr<std::vector<std::string>> n) { fNodeNames = n; fViewsValid = kFALSE; }
};

// reference include/ARayShooter.h:30-54, src/ARayShooter.cxx:122-460 (host generators; the device
// generators are rbg_shoot).  Random kinds use gRandom like the reference.
class ARayShooter : public TObject {
  static void Dir(TGeoRotation* rot, TVector3* v, Double_t* nd) {
    Double_t dir[3] = {0, 0, 1};
    if (v) v->GetXYZ(dir);
    if (rot) rot->LocalToMaster(dir, nd);
    else memcpy(nd, dir, sizeof(dir));
  }
  static void Place(TGeoRotation* rot, TGeoTranslation* tr, Double_t* x) {
    Double_t tmp[3];
    if (rot) { rot->LocalToMaster(x, tmp); memcpy(x, tmp, sizeof(tmp)); }
    if (tr) { tr->LocalToMaster(x, tmp); memcpy(x, tmp, sizeof(tmp)); }
  }

 public:
  static ARayArray* Circle(Double_t lambda, Double_t rmax, Int_t nr, Int_t nphi, TGeoRotation* rot = 0, TGeoTranslation* tr = 0, TVector3* v = 0) {
    ARayArray* array = new ARayArray;
    if (0 > rmax || nr < 1 || nphi < 1) return array;
    Double_t nd[3], p[3] = {0, 0, 0};
    Dir(rot, v, nd);
    if (tr) { Double_t q[3]; tr->LocalToMaster(p, q); memcpy(p, q, sizeof(q)); }
    array->AddRaw(p[0], p[1], p[2], 0, nd[0], nd[1], nd[2], lambda);
    for (Int_t i = 0; i < nr; i++) {
      Double_t r = rmax * (i + 1) / nr;
      for (Int_t j = 0; j < nphi * (i + 1); j++) {
        Double_t phi = 2 * TMath::Pi() / nphi / (i + 1) * j;
        Double_t x[3] = {r * std::cos(phi), r * std::sin(phi), 0};
        Place(rot, tr, x);
        array->AddRaw(x[0], x[1], x[2], 0, nd[0], nd[1], nd[2], lambda);
      }
    }
    return array;
  }
  static ARayArray* RandomCircle(Double_t lambda, Double_t rmax, Int_t n, TGeoRotation* rot = 0, TGeoTranslation* tr = 0, TVector3* v = 0) {
    ARayArray* array = new ARayArray;
    if (0 > rmax) return array;
    Double_t nd[3];
    Dir(rot, v, nd);
    array->Reserve(n);
    for (Int_t i = 0; i < n; i++) {
      Double_t rx, ry;
      do {
        rx = gRandom->Uniform(-rmax, rmax);
        ry = gRandom->Uniform(-rmax, rmax);
      } while (std::sqrt(rx * rx + ry * ry) > rmax);
      Double_t x[3] = {rx, ry, 0};
      Place(rot, tr, x);
      array->AddRaw(x[0], x[1], x[2], 0, nd[0], nd[1], nd[2], lambda);
    }
    return array;
  }
  static ARayArray* RandomCone(Double_t lambda, Double_t r, Double_t d, Int_t n, TGeoRotation* rot = 0, TGeoTranslation* tr = 0) {
    ARayArray* array = new ARayArray;
    for (Int_t i = 0; i < n; i++) {
      Double_t x = gRandom->Uniform(-r, r), y = gRandom->Uniform(-r, r);
      if (x * x + y * y > r * r) { i--; continue; }
      Double_t goal[3] = {x, y, d}, start[3] = {0, 0, 0};
      Place(rot, tr, goal);
      Place(nullptr, tr, start);
      array->AddRaw(start[0], start[1], start[2], 0, goal[0] - start[0], goal[1] - start[1], goal[2] - start[2], lambda);
    }
    return array;
  }
  static ARayArray* RandomRectangle(Double_t lambda, Double_t dx, Double_t dy, Int_t n, TGeoRotation* rot = 0, TGeoTranslation* tr = 0, TVector3* v = 0) {
    ARayArray* array = new ARayArray;
    if (dx < 0 || dy < 0 || n < 1) return array;
    Double_t nd[3];
    Dir(rot, v, nd);
    array->Reserve(n);
    for (Int_t i = 0; i < n; i++) {
      Double_t rx = gRandom->Uniform(-dx / 2., dx / 2.), ry = gRandom->Uniform(-dy / 2., dy / 2.);
      Double_t x[3] = {rx, ry, 0};
      Place(rot, tr, x);
      array->AddRaw(x[0], x[1], x[2], 0, nd[0], nd[1], nd[2], lambda);
    }
    return array;
  }
  static ARayArray* RandomSphere(Double_t lambda, Int_t n, TGeoTranslation* tr = 0) {
    ARayArray* array = new ARayArray;
    for (Int_t i = 0; i < n; i++) {
      Double_t dir[3], p[3] = {0, 0, 0};
      gRandom->Sphere(dir[0], dir[1], dir[2], 1);
      Place(nullptr, tr, p);
      array->AddRaw(p[0], p[1], p[2], 0, dir[0], dir[1], dir[2], lambda);
    }
    return array;
  }
  static ARayArray* RandomSphericalCone(Double_t lambda, Int_t n, Double_t theta, TGeoRotation* rot = 0, TGeoTranslation* tr = 0) {
    ARayArray* array = new ARayArray;
    for (Int_t i = 0; i < n; i++) {
      Double_t ran = gRandom->Uniform(std::cos(theta * TMath::DegToRad()), 1), theta_ = TMath::ACos(ran), phi = gRandom->Uniform(0, TMath::TwoPi());
      Double_t dir[3] = {std::sin(theta_) * std::cos(phi), std::sin(theta_) * std::sin(phi), std::cos(theta_)}, nd[3], p[3] = {0, 0, 0};
      if (rot) rot->LocalToMaster(dir, nd);
      else memcpy(nd, dir, sizeof(dir));
      Place(nullptr, tr, p);
      array->AddRaw(p[0], p[1], p[2], 0, nd[0], nd[1], nd[2], lambda);
    }
    return array;
  }
  static ARayArray* RandomSquare(Double_t lambda, Double_t d, Int_t n, TGeoRotation* rot = 0, TGeoTranslation* tr = 0, TVector3* v = 0) {
    return RandomRectangle(lambda, d, d, n, rot, tr, v);
  }
  static ARayArray* Rectangle(Double_t lambda, Double_t dx, Double_t dy, Int_t nx, Int_t ny, TGeoRotation* rot = 0, TGeoTranslation* tr = 0, TVector3* v = 0) {
    ARayArray* array = new ARayArray;
    if (dx < 0 || dy < 0 || nx < 1 || ny < 1) return array;
    Double_t nd[3];
    Dir(rot, v, nd);
    Double_t deltax = nx == 1 ? dx / 2 : dx / (nx - 1), deltay = ny == 1 ? dy / 2 : dy / (ny - 1);
    array->Reserve(size_t(nx) * ny);
    for (Int_t i = 0; i < nx; i++)
      for (Int_t j = 0; j < ny; j++) {
        Double_t x[3] = {i * deltax - dx / 2, j * deltay - dy / 2, 0};
        Place(rot, tr, x);
        array->AddRaw(x[0], x[1], x[2], 0, nd[0], nd[1], nd[2], lambda);
      }
    return array;
  }
  static ARayArray* Square(Double_t lambda, Double_t d, Int_t n, TGeoRotation* rot = 0, TGeoTranslation* tr = 0, TVector3* v = 0) {
    return Rectangle(lambda, d, d, n, n, rot, tr, v);
  }
};

// ============================================================================ scene export (flat tables)
struct ASceneExport {
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
  rbg_scene_desc desc;

  std::map<const TGeoShape*, int> shape_id;
  std::map<const TGeoMatrix*, int> matrix_id;
  std::map<const TGeoVolume*, int> volume_id;
  std::map<const TGraph*, int> graph_id;
  std::map<const TH2*, int> th2_id;
  std::map<const ARefractiveIndex*, int> index_id;
  std::map<const AMultilayer*, int> multilayer_id;
  std::map<const TGraph2D*, int> graph2d_id;

  int AddMatrix(const TGeoMatrix* m) {
    if (!m || m->IsIdentity()) return -1;
    auto it = matrix_id.find(m);
    if (it != matrix_id.end()) return it->second;
    rbg_matrix r;
    memcpy(r.rot, m->GetRotationMatrix(), sizeof(r.rot));
    memcpy(r.tr, m->GetTranslation(), sizeof(r.tr));
    matrices.push_back(r);
    return matrix_id[m] = (int)matrices.size() - 1;
  }
  int AddShape(const TGeoShape* s) {
    if (!s) throw std::runtime_error("volume without shape");
    auto it = shape_id.find(s);
    if (it != shape_id.end()) return it->second;
    rbg_shape r;
    r.left = r.right = r.lmat = r.rmat = -1;
    r.ipar = (int)dpar.size();
    auto P = [&](std::initializer_list<double> v) { dpar.insert(dpar.end(), v); };
    switch (s->Kind()) {
      case TGeoShape::kBBox: {
        auto* b = static_cast<const TGeoBBox*>(s);
        r.type = RBG_SHAPE_BBOX;
        P({b->GetDX(), b->GetDY(), b->GetDZ(), b->GetOrigin()[0], b->GetOrigin()[1], b->GetOrigin()[2]});
        break;
      }
      case TGeoShape::kTube: {
        auto* t = static_cast<const TGeoTube*>(s);
        r.type = RBG_SHAPE_TUBE;
        P({t->GetRmin(), t->GetRmax(), t->GetDz()});
        break;
      }
      case TGeoShape::kSphere: {
        auto* t = static_cast<const TGeoSphere*>(s);
        r.type = RBG_SHAPE_SPHERE;
        P({t->GetRmin(), t->GetRmax(), t->GetTheta1(), t->GetTheta2(), t->GetPhi1(), t->GetPhi2()});
        break;
      }
      case TGeoShape::kParaboloid: {
        auto* t = static_cast<const TGeoParaboloid*>(s);
        r.type = RBG_SHAPE_PARABOLOID;
        P({t->GetRlo(), t->GetRhi(), t->GetDz()});
        break;
      }
      case TGeoShape::kPgon: {
        auto* t = static_cast<const TGeoPgon*>(s);
        r.type = RBG_SHAPE_PGON;
        P({t->GetPhi1(), t->GetDphi(), (double)t->GetNedges(), (double)t->GetNz()});
        for (Int_t i = 0; i < t->GetNz(); i++) P({t->GetZ(i), t->GetRmin(i), t->GetRmax(i)});
        break;
      }
      case TGeoShape::kPcon: {
        auto* t = static_cast<const TGeoPcon*>(s);
        r.type = RBG_SHAPE_PCON;
        P({t->GetPhi1(), t->GetDphi(), (double)t->GetNz()});
        for (Int_t i = 0; i < t->GetNz(); i++) P({t->GetZ(i), t->GetRmin(i), t->GetRmax(i)});
        break;
      }
      case TGeoShape::kAsphere: {
        auto* t = static_cast<const AGeoAsphericDisk*>(s);
        r.type = RBG_SHAPE_ASPHERE;
        P({t->GetZ1(), t->GetZ2(), t->GetCurve1(), t->GetCurve2(), t->GetKappa1(), t->GetKappa2(), t->GetRmin(), t->GetRmax(),
           (double)t->GetNPol1(), (double)t->GetNPol2(), t->GetOrigin()[2], t->GetDZ()});
        for (Int_t i = 0; i < t->GetNPol1(); i++) dpar.push_back(t->GetK1()[i]);
        for (Int_t i = 0; i < t->GetNPol2(); i++) dpar.push_back(t->GetK2()[i]);
        break;
      }
      case TGeoShape::kWinston2D: {
        auto* t = static_cast<const AGeoWinstonCone2D*>(s);
        r.type = RBG_SHAPE_WINSTON2D;
        P({t->GetR1(), t->GetR2(), t->GetDY()});
        break;
      }
      case TGeoShape::kWinstonPoly: {
        auto* t = static_cast<const AGeoWinstonConePoly*>(s);
        r.type = RBG_SHAPE_WINSTONPOLY;
        P({t->GetR1(), t->GetR2(), (double)t->GetPolyN()});
        break;
      }
      case TGeoShape::kArb8: {
        auto* t = static_cast<const TGeoArb8*>(s);
        r.type = RBG_SHAPE_ARB8;
        P({t->GetDz()});
        for (int i = 0; i < 16; i++) dpar.push_back(t->GetVertices()[i]);
        break;
      }
      case TGeoShape::kXtru: {
        auto* t = static_cast<const TGeoXtru*>(s);
        r.type = RBG_SHAPE_XTRU;
        P({(double)t->GetNvert(), (double)t->GetNz()});
        for (Int_t i = 0; i < t->GetNvert(); i++) P({t->GetX(i), t->GetY(i)});
        for (Int_t i = 0; i < t->GetNz(); i++) P({t->GetZ(i), t->GetXOffset(i), t->GetYOffset(i), t->GetScale(i)});
        break;
      }
      case TGeoShape::kComposite: {
        auto* node = static_cast<const TGeoCompositeShape*>(s)->GetBoolNode();
        r.type = node->op == TGeoBoolNode::kUnion ? RBG_SHAPE_UNION : (node->op == TGeoBoolNode::kIntersection ? RBG_SHAPE_INTERSECTION : RBG_SHAPE_SUBTRACTION);
        r.left = AddShape(node->left);
        r.right = AddShape(node->right);
        r.lmat = AddMatrix(node->lmat);
        r.rmat = AddMatrix(node->rmat);
        r.ipar = (int)dpar.size();
        break;
      }
    }
    r.npar = (int)dpar.size() - r.ipar;
    shapes.push_back(r);
    return shape_id[s] = (int)shapes.size() - 1;
  }
  int AddGraph(const TGraph* g) {
    if (!g || g->GetN() == 0) return -1;
    auto it = graph_id.find(g);
    if (it != graph_id.end()) return it->second;
    std::vector<std::pair<double, double>> pts;
    for (Int_t i = 0; i < g->GetN(); i++) pts.emplace_back(g->GetX()[i], g->GetY()[i]);
    std::stable_sort(pts.begin(), pts.end(), [](auto& a, auto& b) { return a.first < b.first; });
    rbg_graph r = {(int)gx.size(), (int)pts.size()};
    for (auto& p : pts) { gx.push_back(p.first); gy.push_back(p.second); }
    graphs.push_back(r);
    return graph_id[g] = (int)graphs.size() - 1;
  }
  int AddTH2(const TH2* h) {
    if (!h) return -1;
    auto it = th2_id.find(h);
    if (it != th2_id.end()) return it->second;
    rbg_th2 r;
    r.first = (int)th2v.size();
    r.nx = h->GetNbinsX(); r.ny = h->GetNbinsY(); r.pad = 0;
    r.xmin = h->GetXaxis()->GetXmin(); r.xmax = h->GetXaxis()->GetXmax();
    r.ymin = h->GetYaxis()->GetXmin(); r.ymax = h->GetYaxis()->GetXmax();
    for (Int_t j = 1; j <= r.ny; j++)
      for (Int_t i = 1; i <= r.nx; i++) th2v.push_back(h->GetBinContent(i, j));
    th2.push_back(r);
    return th2_id[h] = (int)th2.size() - 1;
  }
  int AddIndex(const ARefractiveIndex* x) {
    if (!x) return -1;
    auto it = index_id.find(x);
    if (it != index_id.end()) return it->second;
    rbg_index r;
    memset(&r, 0, sizeof(r));
    r.kind = x->Kind();
    r.mix_a = r.mix_b = -1;
    r.ngraph = AddGraph(x->GetRefractiveIndexGraph().get());
    r.kgraph = AddGraph(x->GetExtinctionCoefficientGraph().get());
    if (x->Par()) memcpy(r.par, x->Par(), sizeof(r.par));
    if (r.kind == RBG_INDEX_MIXED) {
      auto* mx = static_cast<const AMixedRefractiveIndex*>(x);
      r.mix_a = AddIndex(mx->GetA().get());
      r.mix_b = AddIndex(mx->GetB().get());
      r.frac_a = mx->GetFractionA();
      r.frac_b = mx->GetFractionB();
    }
    indices.push_back(r);
    return index_id[x] = (int)indices.size() - 1;
  }
  int AddGraph2D(const TGraph2D* g) {  // baked to its Delaunay triangle list (TGraph2D::Interpolate)
    if (!g) return -1;
    auto it = graph2d_id.find(g);
    if (it != graph2d_id.end()) return it->second;
    const std::vector<Int_t>& t = g->GetTriangles();
    rbg_graph2d r;
    r.first_tri = (int32_t)(tri.size() / 3);
    r.ntri = (int32_t)(t.size() / 3);
    int32_t base = (int32_t)g2x.size();
    for (Int_t v : t) tri.push_back(base + v);
    g2x.insert(g2x.end(), g->GetX(), g->GetX() + g->GetN());
    g2y.insert(g2y.end(), g->GetY(), g->GetY() + g->GetN());
    g2z.insert(g2z.end(), g->GetZ(), g->GetZ() + g->GetN());
    graph2ds.push_back(r);
    return graph2d_id[g] = (int)graph2ds.size() - 1;
  }
  int AddMultilayer(const AMultilayer* m) {
    if (!m) return -1;
    auto it = multilayer_id.find(m);
    if (it != multilayer_id.end()) return it->second;
    rbg_multilayer r;
    r.n = (int)m->GetNLayers();
    std::vector<rbg_layer> tmp;
    for (int i = 0; i < r.n; i++) {
      rbg_layer l;
      l.index = AddIndex(m->GetIndexList()[i].get());
      l.incoherent = (i == 0 || i == r.n - 1 || !m->GetCoherentList()[i]) ? 1 : 0;
      l.thickness = m->GetThicknessList()[i];
      tmp.push_back(l);
    }
    r.first = (int)layers.size();
    layers.insert(layers.end(), tmp.begin(), tmp.end());
    r.table_r = AddTH2(m->GetPrecalculatedReflectanceMixed().get());
    r.table_t = AddTH2(m->GetPrecalculatedTransmittanceMixed().get());
    if (r.table_r < 0 || r.table_t < 0) r.table_r = r.table_t = -1;
    multilayers.push_back(r);
    return multilayer_id[m] = (int)multilayers.size() - 1;
  }
  void CollectVolumes(const TGeoVolume* v, std::vector<const TGeoVolume*>& order) {
    if (volume_id.count(v)) return;
    volume_id[v] = (int)order.size();
    order.push_back(v);
    for (Int_t i = 0; i < v->GetNdaughters(); i++) CollectVolumes(v->GetNode(i)->GetVolume(), order);
  }
  void Finish(int top) {
    memset(&desc, 0, sizeof(desc));
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
  void BuildFromTop(const TGeoVolume* topv) {
    std::vector<const TGeoVolume*> order;
    CollectVolumes(topv, order);
    volumes.resize(order.size());
    for (size_t i = 0; i < order.size(); i++) {
      const TGeoVolume* v = order[i];
      rbg_volume& r = volumes[i];
      memset(&r, 0, sizeof(r));
      auto* oc = dynamic_cast<const AOpticalComponent*>(v);
      r.type = oc ? oc->OpticalType() : RBG_OTHER;
      r.shape = AddShape(v->GetShape());
      r.index = r.mirror = r.focal = -1;
      r.name = (int)names.size();
      const char* nm = v->GetName();
      names.insert(names.end(), nm, nm + strlen(nm) + 1);
      if (auto* l = dynamic_cast<const ALens*>(v)) r.index = AddIndex(l->GetIndex().get());
      if (auto* m = dynamic_cast<const AMirror*>(v)) {
        rbg_mirror mm;
        mm.constant = m->GetConstantReflectance();
        mm.graph1d = AddGraph(m->GetReflectance1D().get());
        mm.th2 = AddTH2(m->GetReflectanceTH2().get());
        mm.graph2d = AddGraph2D(m->GetReflectance2D().get()); mm.pad = 0;
        mirrors.push_back(mm);
        r.mirror = (int)mirrors.size() - 1;
      }
      if (auto* f = dynamic_cast<const AFocalSurface*>(v)) {
        if (f->GetQELambdaGraph() || f->GetQEAngleGraph()) {
          rbg_focal ff = {AddGraph(f->GetQELambdaGraph()), AddGraph(f->GetQEAngleGraph())};
          focals.push_back(ff);
          r.focal = (int)focals.size() - 1;
        }
      }
      r.first_node = (int)nodes.size();
      r.nnodes = v->GetNdaughters();
      for (Int_t k = 0; k < v->GetNdaughters(); k++) {
        TGeoNode* n = v->GetNode(k);
        rbg_node nn = {volume_id[n->GetVolume()], AddMatrix(n->GetMatrix()), n->GetNumber(), n->IsOverlapping() ? 1 : 0};
        nodes.push_back(nn);
      }
    }
    for (size_t i = 0; i < order.size(); i++) {  // borders after all volume ids are known
      rbg_volume& r = volumes[i];
      r.first_border = (int)borders.size();
      auto* oc = dynamic_cast<const AOpticalComponent*>(order[i]);
      if (oc)
        for (auto* b : oc->GetBorders()) {
          rbg_border bb;
          AOpticalComponent* c2 = b->GetComponent2();
          bb.vol2 = !c2 ? -1 : (volume_id.count(c2) ? volume_id[c2] : -2);
          bb.multilayer = AddMultilayer(b->GetMultilayer().get());
          bb.lambertian = b->IsLambertian() ? 1 : 0;
          bb.pad = 0;
          bb.sigma = b->GetGaussianRoughness();
          borders.push_back(bb);
        }
      r.nborders = (int)borders.size() - r.first_border;
    }
    Finish(0);
  }
};

inline void AMultilayer::DeviceTMM(Int_t n, const Double_t* th, const Double_t* lam, Double_t* R, Double_t* T) const {
  ASceneExport ex;
  int id = ex.AddMultilayer(this);
  ex.Finish(-1);
  rbg_scene* sc = nullptr;
  if (rbg_scene_create(&ex.desc, 0, &sc) != RBG_OK) throw std::runtime_error(std::string("AMultilayer: ") + rbg_last_error());
  int rc = rbg_tmm_host(sc, id, n, th, lam, R, T);
  rbg_scene_destroy(sc);
  if (rc != RBG_OK) throw std::runtime_error(std::string("AMultilayer: ") + rbg_last_error());
}

inline void AMultilayer::DeviceTMMGeneral(Int_t mode, Int_t pol, Bool_t reverse, Int_t n, const std::complex<Double_t>* th, const Double_t* lam, Double_t* R,
                                          Double_t* T) const {
  ASceneExport ex;
  int id = ex.AddMultilayer(this);
  ex.Finish(-1);
  std::vector<Double_t> re(n), im(n);
  for (Int_t i = 0; i < n; i++) { re[i] = th[i].real(); im[i] = th[i].imag(); }
  rbg_scene* sc = nullptr;
  if (rbg_scene_create(&ex.desc, 0, &sc) != RBG_OK) throw std::runtime_error(std::string("AMultilayer: ") + rbg_last_error());
  int rc = rbg_tmm_general_host(sc, id, mode, pol, reverse ? 1 : 0, n, re.data(), im.data(), lam, R, T);
  rbg_scene_destroy(sc);
  if (rc != RBG_OK) throw std::runtime_error(std::string("AMultilayer: ") + rbg_last_error());
}

// ============================================================================ AOpticsManager
// reference include/AOpticsManager.h:38-105, src/AOpticsManager.cxx:27-49,304-332,523-594
class AOpticsManager : public TGeoManager {
 public:

 private:
  Int_t fLimit = 100;
  Bool_t fDisableFresnelReflection = kFALSE;
  UInt_t fQuirks = RBG_QUIRKS_DEFAULT;
  ULong64_t fSeed = 20180601ULL;
  ULong64_t fRayCounter = 0;  // global ray index so that successive calls use fresh random streams
  Int_t fDevice = 0;
  Int_t fHistoryDepth = -1;  // < 0: automatic (see SetHistoryDepth)
  std::vector<Double_t> fHistBuf;     // receive buffers of rbg_trace_history, reused across calls
  std::vector<int32_t> fHistNodeBuf;
  rbg_scene* fScene = nullptr;
  rbg_multi* fMulti = nullptr;             // scene replicas on several GPUs (GetNumberOfGPUs() > 1)
  std::string fSceneKey;
  std::shared_ptr<ASceneExport> fExport;   // flattened scene of the geometry epoch fExportEpoch
  unsigned long long fExportEpoch = 0;
  std::shared_ptr<std::vector<std::string>> fNodeNames;

  static std::string Key(const ASceneExport& e) {
    std::string k;
    auto app = [&](const void* p, size_t n) { if (n) k.append((const char*)p, n); };
    app(e.shapes.data(), e.shapes.size() * sizeof(rbg_shape)); app(e.dpar.data(), e.dpar.size() * 8);
    app(e.matrices.data(), e.matrices.size() * sizeof(rbg_matrix)); app(e.nodes.data(), e.nodes.size() * sizeof(rbg_node));
    app(e.volumes.data(), e.volumes.size() * sizeof(rbg_volume)); app(e.borders.data(), e.borders.size() * sizeof(rbg_border));
    app(e.graphs.data(), e.graphs.size() * sizeof(rbg_graph)); app(e.gx.data(), e.gx.size() * 8); app(e.gy.data(), e.gy.size() * 8);
    app(e.th2.data(), e.th2.size() * sizeof(rbg_th2)); app(e.th2v.data(), e.th2v.size() * 8);
    app(e.indices.data(), e.indices.size() * sizeof(rbg_index)); app(e.mirrors.data(), e.mirrors.size() * sizeof(rbg_mirror));
    app(e.focals.data(), e.focals.size() * sizeof(rbg_focal)); app(e.multilayers.data(), e.multilayers.size() * sizeof(rbg_multilayer));
    app(e.layers.data(), e.layers.size() * sizeof(rbg_layer)); app(e.names.data(), e.names.size());
    return k;
  }

 public:
  enum { kLens, kObs, kMirror, kFocus, kOpt, kOther, kNull };
  AOpticsManager() {}
  AOpticsManager(const char* name, const char* title) : TGeoManager(name, title) {}
  ~AOpticsManager() override {
    if (fScene) rbg_scene_destroy(fScene);
    if (fMulti) rbg_multi_destroy(fMulti);
  }
  static Double_t km() { return 1e3 * m(); }
  static Double_t m() { return 1e2 * cm(); }
  static Double_t cm() { return 1; }
  static Double_t mm() { return 1e-3 * m(); }
  static Double_t um() { return 1e-6 * m(); }
  static Double_t nm() { return 1e-9 * m(); }
  static Double_t inch() { return 2.54 * cm(); }
  static Double_t s() { return 1.; }
  static Double_t ms() { return 1e-3 * s(); }
  static Double_t us() { return 1e-6 * s(); }
  static Double_t ns() { return 1e-9 * s(); }
  static Double_t deg() { return TMath::DegToRad(); }
  static Double_t rad() { return 1.; }

  void DisableFresnelReflection(Bool_t disable) { fDisableFresnelReflection = disable; }
  void SetLimit(Int_t n) { if (n > 0) fLimit = n; }
  Int_t GetLimit() const { return fLimit; }
  static Bool_t IsLens(const TGeoNode* n) { return Type(n) == RBG_LENS; }
  static Bool_t IsMirror(const TGeoNode* n) { return Type(n) == RBG_MIRROR; }
  static Bool_t IsFocalSurface(const TGeoNode* n) { return Type(n) == RBG_FOCUS; }
  static Bool_t IsObscuration(const TGeoNode* n) { return Type(n) == RBG_OBS; }
  static Bool_t IsOpticalComponent(const TGeoNode* n) { return Type(n) == RBG_OPT; }
  static Int_t Type(const TGeoNode* n) {
    if (!n) return RBG_NULL;
    auto* oc = dynamic_cast<const AOpticalComponent*>(n->GetVolume());
    return oc ? oc->OpticalType() : RBG_OTHER;
  }
  // extensions (not in the reference): RNG seed, quirk switches, device, test hook, flat export
  void SetSeed(ULong64_t seed) { fSeed = seed; fRayCounter = 0; }
  void SetQuirks(UInt_t q) { fQuirks = q; }
  void SetDevice(Int_t d) { fDevice = d; }
  // polyline record (ARay::GetPoint(i), node history): points kept per ray.  < 0 = automatic: min(limit, 16) points
  // for batches below 262144 rays (the reference always keeps every point; large batches keep first/last only
  // unless a depth is requested here), 0 = never.
  void SetHistoryDepth(Int_t n) { fHistoryDepth = n; }
  Int_t GetHistoryDepth() const { return fHistoryDepth; }
  // The flattened scene (export + device tables) is kept across calls until a mutator of a shape, matrix, volume, table or
  // optical property ran anywhere (RbGeomEpoch, RootCompat.h).  InvalidateScene() forces a re-export, for changes made behind
  // the classes' backs (writes through raw pointers); RB_NO_SCENE_CACHE=1 re-exports on every call like round 1 did.
  void InvalidateScene() { fExport.reset(); }
  // SetMaxThreads(n) + SetMultiThread(true) is how a ROBAST macro asks for n workers (src/AOpticsManager.cxx:529-568): here the
  // workers are GPUs — min(n, visible devices) of them, rays in contiguous chunks, geometry replicated (rbg_multi_trace).
  Int_t GetNumberOfGPUs() const {
    Int_t g = IsMultiThread() ? std::min<Int_t>(GetMaxThreads(), rbg_device_count()) : 1;
    return g < 1 ? 1 : g;
  }
  std::shared_ptr<ASceneExport> ExportScene() const {
    if (!fTopVolume) throw std::runtime_error("AOpticsManager: no top volume");
    auto e = std::make_shared<ASceneExport>();
    e->BuildFromTop(fTopVolume);
    return e;
  }

  // the hot path: SoA batch through the C ABI
  void TraceNonSequential(ARayArray& array) {
    ARayArray::Table& T = array.GetTable();
    std::vector<size_t> run;
    for (size_t i = 0; i < T.size(); i++)
      if (T.status[i] == RBG_RUN) run.push_back(i);
    if (run.empty()) return;
    static const bool no_cache = getenv("RB_NO_SCENE_CACHE") != nullptr;
    const unsigned long long epoch = RbGeomEpoch().load(std::memory_order_relaxed);
    const bool fresh = !fExport || epoch != fExportEpoch || no_cache;
    if (fresh) {
      fExport = ExportScene();
      fExportEpoch = epoch;
    }
    const std::shared_ptr<ASceneExport>& ex = fExport;
    size_t n = run.size();
    bool contiguous = run.back() - run.front() + 1 == n;
    std::vector<Double_t> buf;
    std::vector<int32_t> ibuf;
    Double_t* col[8];
    int32_t* icol[3];
    size_t o = run.front();
    if (contiguous) {
      Double_t* c[8] = {&T.x[o], &T.y[o], &T.z[o], &T.t[o], &T.dx[o], &T.dy[o], &T.dz[o], &T.lambda[o]};
      memcpy(col, c, sizeof(c));
      icol[0] = &T.status[o]; icol[1] = &T.last_node[o]; icol[2] = &T.npoints[o];
    } else {
      buf.resize(8 * n);
      ibuf.resize(3 * n);
      for (int k = 0; k < 8; k++) col[k] = buf.data() + k * n;
      for (int k = 0; k < 3; k++) icol[k] = ibuf.data() + k * n;
      for (size_t j = 0; j < n; j++) {
        size_t i = run[j];
        col[0][j] = T.x[i]; col[1][j] = T.y[i]; col[2][j] = T.z[i]; col[3][j] = T.t[i];
        col[4][j] = T.dx[i]; col[5][j] = T.dy[i]; col[6][j] = T.dz[i]; col[7][j] = T.lambda[i];
      }
    }
    rbg_rays r;
    memset(&r, 0, sizeof(r));
    r.n = (int64_t)n;
    r.on_device = 0;
    r.x = col[0]; r.y = col[1]; r.z = col[2]; r.t = col[3]; r.dx = col[4]; r.dy = col[5]; r.dz = col[6]; r.lambda = col[7];
    r.ox = col[0]; r.oy = col[1]; r.oz = col[2]; r.ot = col[3]; r.odx = col[4]; r.ody = col[5]; r.odz = col[6];
    r.status = icol[0]; r.last_node = icol[1]; r.npoints = icol[2];
    rbg_trace_opts opts;
    memset(&opts, 0, sizeof(opts));
    opts.limit = fLimitForCall > 0 ? fLimitForCall : fLimit;
    opts.disable_fresnel = fDisableFresnelReflection;
    opts.quirks = fQuirks;
    opts.seed = fSeed;
    opts.ray_id_offset = fRayCounter;
    fRayCounter += n;
    int rc;
    {  // the CUDA library is the only trace path: no CPU fallback, no pluggable tracer
      std::string key = (fresh || !fScene) ? Key(*ex) : fSceneKey;  // untouched geometry: same tables, nothing to compare
      if (!fScene || key != fSceneKey) {
        if (fScene) rbg_scene_destroy(fScene);
        fScene = nullptr;
        if (fMulti) rbg_multi_destroy(fMulti);
        fMulti = nullptr;
        if (rbg_scene_create(&ex->desc, fDevice, &fScene) != RBG_OK)
          throw std::runtime_error(std::string("AOpticsManager::TraceNonSequential: ") + rbg_last_error());
        fSceneKey = key;
        fNodeNames = std::make_shared<std::vector<std::string>>();
        for (int i = 0; i < rbg_scene_num_nodes(fScene); i++) fNodeNames->push_back(rbg_scene_node_name(fScene, i));
      }
      int32_t depth = fHistoryDepth >= 0 ? fHistoryDepth : (n < 262144 ? std::min<Int_t>(fLimit, 16) : 0);
      // receive buffers in the layout of rbg_history (point k of ray j at k*n + j); kept across calls, never cleared: only
      // the entries k < npoints of each ray are read
      rbg_history hist;
      memset(&hist, 0, sizeof(hist));
      if (depth > 0) {
        if (fHistBuf.size() < (size_t)4 * depth * n) fHistBuf.resize((size_t)4 * depth * n);
        if (fHistNodeBuf.size() < (size_t)depth * n) fHistNodeBuf.resize((size_t)depth * n);
        hist.max_points = depth;
        hist.hx = fHistBuf.data(); hist.hy = fHistBuf.data() + (size_t)depth * n; hist.hz = fHistBuf.data() + (size_t)2 * depth * n;
        hist.ht = fHistBuf.data() + (size_t)3 * depth * n;
        hist.hnode = fHistNodeBuf.data();
      }
      const Int_t ngpu = GetNumberOfGPUs();
      if (ngpu > 1 && depth == 0 && n >= (size_t)ngpu * 262144) {  // fan out over the GPUs of the box
        if (fMulti && rbg_multi_num_devices(fMulti) != ngpu) { rbg_multi_destroy(fMulti); fMulti = nullptr; }
        if (!fMulti && rbg_multi_create(&ex->desc, ngpu, nullptr, &fMulti) != RBG_OK)
          throw std::runtime_error(std::string("AOpticsManager::TraceNonSequential: ") + rbg_last_error());
        rc = rbg_multi_trace(fMulti, &opts, &r);
      } else rc = rbg_trace_history(fScene, &opts, &r, depth > 0 ? &hist : nullptr, nullptr);
      if (rc != RBG_OK) throw std::runtime_error(std::string("AOpticsManager::TraceNonSequential: ") + rbg_last_error());
      if (depth > 0 || T.HasHistory()) {  // rebuild the packed per-ray records: traced rays get their new polyline
        T.EnsureHistoryOffsets();
        std::vector<int64_t> traced(T.size(), -1);
        for (size_t j = 0; j < n; j++) traced[run[j]] = (int64_t)j;
        std::vector<int64_t> noff(T.size() + 1, 0);
        for (size_t i = 0; i < T.size(); i++) {
          int64_t c = traced[i] >= 0 ? (depth > 0 ? std::min<int32_t>(icol[2][traced[i]], depth) : 0) : T.hoff[i + 1] - T.hoff[i];
          noff[i + 1] = noff[i] + c;
        }
        std::vector<Double_t> np((size_t)4 * noff.back());
        std::vector<int32_t> nn((size_t)noff.back());
        for (size_t i = 0; i < T.size(); i++) {
          int64_t c = noff[i + 1] - noff[i], dst = noff[i];
          if (traced[i] >= 0) {
            size_t j = (size_t)traced[i];
            for (int64_t k = 0; k < c; k++) {
              size_t src = (size_t)k * n + j;
              np[4 * (dst + k)] = hist.hx[src]; np[4 * (dst + k) + 1] = hist.hy[src]; np[4 * (dst + k) + 2] = hist.hz[src]; np[4 * (dst + k) + 3] = hist.ht[src];
              nn[dst + k] = hist.hnode[src];
            }
          } else if (c > 0) {
            memcpy(&np[4 * dst], &T.hpts[4 * T.hoff[i]], (size_t)c * 4 * sizeof(Double_t));
            memcpy(&nn[dst], &T.hnode[T.hoff[i]], (size_t)c * sizeof(int32_t));
          }
        }
        T.hoff.swap(noff); T.hpts.swap(np); T.hnode.swap(nn);
      }
    }
    if (!contiguous)
      for (size_t j = 0; j < n; j++) {
        size_t i = run[j];
        T.x[i] = col[0][j]; T.y[i] = col[1][j]; T.z[i] = col[2][j]; T.t[i] = col[3][j];
        T.dx[i] = col[4][j]; T.dy[i] = col[5][j]; T.dz[i] = col[6][j];
        T.status[i] = icol[0][j]; T.last_node[i] = icol[1][j]; T.npoints[i] = icol[2][j];
      }
    if (fNodeNames) array.SetNodeNames(fNodeNames);
  }
  void TraceNonSequential(ARayArray* array) { TraceNonSequential(*array); }
  void TraceNonSequential(ARay& ray) {
    if (!ray.IsRunning()) return;
    ARayArray tmp;
    Double_t p[4], d[3];
    ray.GetLastPoint(p);
    ray.GetDirection(d);
    tmp.AddRaw(p[0], p[1], p[2], p[3], d[0], d[1], d[2], ray.GetLambda());
    TraceWithHeldPoints(tmp, ray.GetNpoints());
    const ARayArray::Table& T = tmp.GetTable();
    Double_t last[4] = {T.x[0], T.y[0], T.z[0], T.t[0]}, dir[3] = {T.dx[0], T.dy[0], T.dz[0]};
    const char* nn = (fNodeNames && T.last_node[0] >= 0) ? (*fNodeNames)[T.last_node[0]].c_str() : nullptr;
    const bool fresh = ray.GetNpoints() == 1;
    ray.SetTraced(last, dir, T.status[0], ray.GetNpoints() + T.npoints[0] - 1, nn);
    if (fresh && T.HistCount(0) > 0) ray.SetHistory(&T.hpts[0], T.HistCount(0), &T.hnode[0], fNodeNames.get());
  }
  void TraceNonSequential(ARay* ray) { TraceNonSequential(*ray); }
  // a TObjArray of ARay (src/AOpticsManager.cxx:335): the running rays go through the batch path in ONE call (one per number of
  // points the rays already hold: the fLimit test counts those, see TraceWithHeldPoints)
  void TraceNonSequential(TObjArray* array) {
    std::map<Int_t, std::vector<ARay*>> groups;
    for (Int_t i = 0; i <= array->GetLast(); i++) {
      auto* r = dynamic_cast<ARay*>(array->At(i));
      if (r && r->IsRunning()) groups[r->GetNpoints()].push_back(r);
    }
    for (auto& grp : groups) {
      std::vector<ARay*>& rays = grp.second;
      ARayArray tmp;
      for (ARay* r : rays) {
        Double_t p[4], d[3];
        r->GetLastPoint(p);
        r->GetDirection(d);
        tmp.AddRaw(p[0], p[1], p[2], p[3], d[0], d[1], d[2], r->GetLambda());
      }
      TraceWithHeldPoints(tmp, grp.first);
      const ARayArray::Table& T = tmp.GetTable();
      for (size_t j = 0; j < rays.size(); j++) {
        ARay& ray = *rays[j];
        Double_t last[4] = {T.x[j], T.y[j], T.z[j], T.t[j]}, dir[3] = {T.dx[j], T.dy[j], T.dz[j]};
        const char* nn = (fNodeNames && T.last_node[j] >= 0) ? (*fNodeNames)[T.last_node[j]].c_str() : nullptr;
        const bool was_fresh = ray.GetNpoints() == 1;
        ray.SetTraced(last, dir, T.status[j], ray.GetNpoints() + T.npoints[j] - 1, nn);
        if (was_fresh && T.HistCount(j) > 0) ray.SetHistory(&T.hpts[4 * T.hoff[j]], T.HistCount(j), &T.hnode[T.hoff[j]], fNodeNames.get());
      }
    }
  }

 private:
  // The reference suspends a ray when ray->GetNpoints() >= fLimit (src/AOpticsManager.cxx:515-517), counting the points the ARay
  // held before this call.  The device counts from 1, so rays that already hold `held` points are traced with the limit lowered by
  // held - 1 (at least 2: a step is always taken before the test).
  Int_t fLimitForCall = 0;
  void TraceWithHeldPoints(ARayArray& tmp, Int_t held) {
    fLimitForCall = held > 1 ? std::max<Int_t>(2, fLimit - (held - 1)) : 0;
    try {
      TraceNonSequential(tmp);
    } catch (...) {
      fLimitForCall = 0;
      throw;
    }
    fLimitForCall = 0;
  }
};

#endif  // ROBAST_ROBAST_H
