// parabolic_demo.cpp — C++ use of the ROBAST-mirror API on the GPU tracer (same calls as a ROBAST macro):
// a parabolic mirror + focal plane, on-axis and 1 deg off-axis beams, spot statistics from the focused bucket.
//   g++ -std=c++17 -Iinclude/robast examples/parabolic_demo.cpp -Lrobast_b200 -lrobast_b200 -Wl,-rpath,$PWD/robast_b200 -o parabolic_demo
#include <cstdio>

#include "Robast.h"

static const Double_t cm = AOpticsManager::cm(), um = AOpticsManager::um(), nm = AOpticsManager::nm(), m = AOpticsManager::m();

int main() {
  const Double_t radius = 1.5 * m, focal = 3 * m, sag = radius * radius / 4. / focal;
  AOpticsManager* manager = new AOpticsManager("manager", "parabolic demo");
  AOpticalComponent* world = new AOpticalComponent("world", new TGeoBBox("worldbox", 10 * m, 10 * m, 10 * m));
  manager->SetTopVolume(world);
  new TGeoParaboloid("para", 0, radius, sag / 2.);
  (new TGeoTranslation("tr1", 0, 0, sag / 2.))->RegisterYourself();
  (new TGeoTranslation("tr2", 0, 0, sag / 2. - 1 * um))->RegisterYourself();
  world->AddNode(new AMirror("mirror", new TGeoCompositeShape("shell", "para:tr2 - para:tr1")), 1);
  world->AddNode(new AFocalSurface("focal", new TGeoTube("focal_tube", 0, 20 * cm, 10 * um)), 1, new TGeoTranslation("ftr", 0, 0, focal + 10 * um));
  world->AddNode(new AObscuration("obs", new TGeoTube("obs_tube", 0, 20 * cm + 10 * um, 10 * um)), 1, new TGeoTranslation("otr", 0, 0, focal + 30 * um));
  manager->CloseGeometry();
  for (int i = 0; i < 2; i++) {
    Double_t rad = i * 1.0 * TMath::DegToRad();
    TGeoTranslation raytr("raytr", -focal * 2 * TMath::Sin(rad), 0, focal * 2 * TMath::Cos(rad));
    TVector3 dir;
    dir.SetMagThetaPhi(1, TMath::Pi() - rad, 0);
    ARayArray* array = ARayShooter::Square(400 * nm, 5 * m, 301, 0, &raytr, &dir);
    manager->TraceNonSequential(*array);
    TObjArray* focused = array->GetFocused();
    TH2D spot("spot", "", 400, -20, 20, 400, -20, 20);
    for (Int_t j = 0; j <= focused->GetLast(); j++) {
      Double_t p[4];
      ((ARay*)(*focused)[j])->GetLastPoint(p);
      spot.Fill(p[0], p[1]);
    }
    printf("theta=%.1f focused=%d stopped=%d exited=%d mean_x=%.9f rms_x=%.9f rms_y=%.9f\n", i * 1.0, focused->GetLast() + 1, array->GetStopped()->GetLast() + 1,
           array->GetExited()->GetLast() + 1, spot.GetMean(1), spot.GetRMS(1), spot.GetRMS(2));
    delete array;
  }
  return 0;
}
