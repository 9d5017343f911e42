"""The five BASELINE.json configurations, built through the ROBAST-mirror API.

Each builder returns (manager, keepalive) where `manager` is an AOpticsManager holding the geometry
and `keepalive` keeps Python references to shared optical data.  `beam(cfg, ...)` returns the
rbg_shoot_desc parameters of the configuration's synthetic beam.  Geometry parameters follow the
reference tutorials (cited per function, paths under /root/reference/tutorials).
"""
import math
import os

from . import _robast as R

cm, mm, um, nm, m, inch = 1.0, 0.1, 1e-4, 1e-7, 100.0, 2.54
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


# ----------------------------------------------------------------------------- 1: SimpleParabolicTelescope.C:15-83
def simple_parabolic():
    mirror_rad, focal_length, focal_rad = 1.5 * m, 3 * m, 20 * cm
    sag = mirror_rad * mirror_rad / 4. / focal_length
    mgr = R.AOpticsManager("manager", "SimpleParabolicTelescope")
    world = R.AOpticalComponent("world", R.TGeoBBox("worldbox", 10 * m, 10 * m, 10 * m))
    mgr.SetTopVolume(world)
    R.TGeoParaboloid("mirror_para", 0, mirror_rad, sag / 2.)
    R.TGeoTranslation("mirror_tr1", 0, 0, sag / 2.).RegisterYourself()
    R.TGeoTranslation("mirror_tr2", 0, 0, sag / 2. - 1 * um).RegisterYourself()
    comp = R.TGeoCompositeShape("mirror_comp", "mirror_para:mirror_tr2 - mirror_para:mirror_tr1")
    world.AddNode(R.AMirror("mirror", comp), 1)
    world.AddNode(R.AFocalSurface("focal", R.TGeoTube("focal_tube", 0, focal_rad, 10 * um)), 1,
                  R.TGeoTranslation("focal_tr", 0, 0, focal_length + 10 * um))
    world.AddNode(R.AObscuration("obs1", R.TGeoTube("obs_tube1", 0, focal_rad + 10 * um, 10 * um)), 1,
                  R.TGeoTranslation("obs_tr1", 0, 0, focal_length + 30 * um))
    world.AddNode(R.AObscuration("obs2", R.TGeoTube("obs_tube2", focal_rad, focal_rad + 10 * um, 10 * um)), 1,
                  R.TGeoTranslation("obs_tr2", 0, 0, focal_length + 10 * um))
    mgr.CloseGeometry()
    return mgr, []


# ----------------------------------------------------------------------------- 2: DaviesCotton.C:14-182
_DC_X = [0] * 8 + [1.5] * 10 + [-1.5] * 10 + [3] * 9 + [-3] * 9 + [4.5] * 8 + [-4.5] * 8 + [6] * 7 + [-6] * 7 + [7.5] * 6 + [-7.5] * 6
_DC_Y = ([2, 4, 6, 8, -2, -4, -6, -8] + [1, 3, 5, 7, 9, -1, -3, -5, -7, -9] * 2 + [0, 2, 4, 6, 8, -2, -4, -6, -8] * 2 +
         [1, 3, 5, 7, -1, -3, -5, -7] * 2 + [0, 2, 4, 6, -2, -4, -6] * 2 + [1, 3, 5, -1, -3, -5] * 2)


def davies_cotton():
    kF = 16 * m
    mirror_r, mirror_d, mirror_t = kF * 2, 1.2 * m, 0.1 * mm
    camera_d, box_d, box_h = 2.2 * m, 2.5 * m, 1 * m
    mgr = R.AOpticsManager("manager", "Davies-Cotton System")
    mgr.DisableFresnelReflection(True)
    world = R.AOpticalComponent("world", R.TGeoBBox("boxWorld", 20 * m, 20 * m, 20 * m))
    mgr.SetTopVolume(world)
    # mirrors (:52-128)
    cut = R.TGeoPgon("mirCut", 0., 360., 6, 2)
    cut.DefineSection(0, -100 * mm, 0, mirror_d / 2.)
    cut.DefineSection(1, 100 * mm, 0, mirror_d / 2.)
    theta = math.degrees(math.asin(mirror_d / math.sqrt(3) / mirror_r))
    R.TGeoSphere("mirSphere", mirror_r, mirror_r + mirror_t, 180. - theta, 180.)
    R.TGeoTranslation("transZ", 0, 0, mirror_r).RegisterYourself()
    mirror = R.AMirror("mirror", R.TGeoCompositeShape("mirComposite", "mirSphere:transZ*mirCut"))
    dx, dy = mirror_d / math.sqrt(3), mirror_d / 2.
    assert len(_DC_X) == 88 and len(_DC_Y) == 88
    for i in range(88):
        x, y = _DC_X[i] * dx, _DC_Y[i] * dy
        r2 = x * x + y * y
        z = kF - math.sqrt(kF * kF - r2)
        trans = R.TGeoTranslation("mirTrans%d" % i, x, y, z)
        phi = math.degrees(math.atan2(y, x))
        rot = R.TGeoRotation("mirRot%d" % i, -phi + 90., 0, 0)
        th = math.degrees(math.atan2(math.sqrt(r2), 2 * kF - z))
        rot.MultiplyBy(R.TGeoRotation("", phi - 90., th, 0), False)
        world.AddNode(mirror, i + 1, R.TGeoCombiTrans(trans, rot))
    # camera (:130-157)
    world.AddNode(R.AFocalSurface("focalPlane", R.TGeoTube("tubeCamera", 0, camera_d / 2., 1 * mm)), 1, R.TGeoTranslation(0, 0, kF + 1 * mm))
    t = 10 * cm
    R.TGeoBBox("boxCamera", box_d / 2., box_d / 2., box_h / 2.)
    R.TGeoBBox("boxCamera2", box_d / 2. - t, box_d / 2. - t, box_h / 2. - t)
    R.TGeoTranslation("transZ1", 0, 0, kF + box_h / 2.).RegisterYourself()
    R.TGeoTranslation("transZ2", 0, 0, kF + box_h / 2. - t - 1 * mm).RegisterYourself()
    world.AddNode(R.AObscuration("cameraBox", R.TGeoCompositeShape("boxComposite", "boxCamera:transZ1-boxCamera2:transZ2")), 1)
    # masts (:159-182)
    for i in range(4):
        x1, y1, z1 = 5 * m, 5 * m, 0.
        x2, y2, z2 = box_d / 2. + 10 * cm, box_d / 2. + 10 * cm, kF
        c, s = math.cos(math.pi / 2 * i), math.sin(math.pi / 2 * i)
        v1 = R.TVector3(c * x1 - s * y1, s * x1 + c * y1, z1)
        v2 = R.TVector3(c * x2 - s * y2, s * x2 + c * y2, z2)
        tube, combi = R.MakePointToPointTube("mast%d" % i, v1, v2, 10 * cm)
        world.AddNode(R.AObscuration("obsMast%d" % i, tube), 1, combi)
    mgr.CloseGeometry()
    return mgr, []


# ----------------------------------------------------------------------------- 3: SchwarzschildCouder.C:20-90
def schwarzschild_couder():
    mgr = R.AOpticsManager("manager", "SC")
    world = R.AOpticalComponent("world", R.TGeoBBox("worldbox", 30 * m, 30 * m, 30 * m))
    mgr.SetTopVolume(world)
    top = R.AOpticalComponent("top", R.TGeoBBox("topbox", 30 * m, 30 * m, 30 * m))
    kDp, kDpinner, kFp = 9.40 * m, 4.68 * m, 16.915 * m
    kZs, kDs, kFs = 9.980 * m, 6.61 * m, -3.553 * m
    kZf, kFf = 7.631 * m, -1.481 * m
    kZi = [kFp ** -1 * 0.25, kFp ** -3 * -0.189377, kFp ** -5 * -0.604706, kFp ** -7 * -4.21374, kFp ** -9 * 21.8275, kFp ** -11 * -425.160]
    kVi = [kFs ** -1 * 0.25, kFs ** -3 * 0.013625, kFs ** -5 * -0.010453, kFs ** -7 * 0.014241, kFs ** -9 * -0.012213, kFs ** -11 * 0.005184]
    kYi = [kFf ** -1 * 0.25]
    primary = R.AGeoAsphericDisk("primaryV", -1 * um, 0, 0 * m, 0, kDp / 2., kDpinner / 2.)
    primary.SetPolynomials(6, kZi, 6, kZi)
    top.AddNode(R.AMirror("primaryMirror", primary), 1)
    secondary = R.AGeoAsphericDisk("secondaryV", kZs, 0, kZs + 1 * um, 0, kDs / 2., 0 * m)
    secondary.SetPolynomials(6, kVi, 6, kVi)
    top.AddNode(R.AMirror("secondaryMirror", secondary), 1)
    focal = R.AGeoAsphericDisk("focalV", kZf - 1 * mm, 0, kZf, 0, 10 * cm * 7, 0.)
    focal.SetPolynomials(1, kYi, 1, kYi)
    top.AddNodeOverlap(R.AFocalSurface("focalPlane", focal), 1)
    obs = R.AGeoAsphericDisk("focalObsV", focal.CalcF1(10 * cm * 7) - 1 * cm, 0, focal.CalcF1(10 * cm * 7), 0, 10 * cm * 7.3, 0.)
    top.AddNodeOverlap(R.AObscuration("focalObs", obs), 1)
    world.AddNode(top, 1)
    mgr.CloseGeometry()
    return mgr, []


# ----------------------------------------------------------------------------- 4: SchmidtCassegrain.C:34-111
def schmidt_cassegrain(disable_fresnel=False):
    offset = 10 * cm
    mgr = R.AOpticsManager("Zemax", "Zemax")
    top = R.AOpticalComponent("top", R.TGeoBBox("box", 10 * m, 10 * m, 10 * m))
    mgr.SetTopVolume(top)
    disk = R.AGeoAsphericDisk("disk", 0. * inch + offset, 0 / inch, 0.65 * inch + offset, -8.721454939626E-005 / inch, 12 * inch, 0. * inch)
    coeff = [0, 3.68090959E-7 / inch ** 3, 2.73643352E-11 / inch ** 5, 3.20036892E-14 / inch ** 7]
    disk.SetPolynomials(0, [], 4, coeff)
    catalog = R.AGlassCatalog(os.path.join(DATA, "nbk7.agf"))
    bk7 = catalog.GetRefractiveIndex("N-BK7")
    lens = R.ALens("lens", disk)
    lens.SetRefractiveIndex(bk7)
    top.AddNode(lens, 1)
    top.AddNode(R.AObscuration("aperture", R.TGeoTube("tube", 12 * inch, 18 * inch, disk.GetDZ())), 1, R.TGeoTranslation(0, 0, disk.GetOrigin()[2]))
    R.TGeoTube("tube2", 0 * inch, 4.5 * inch, 0.01 * mm)
    box = R.TGeoBBox("box2", 1.210290505556E1 / 2. * inch, 1 * inch, 0.01 * mm)
    bdx = box.GetDX()
    for name, ang in (("tr1", 90), ("tr2", 210), ("tr3", 330)):
        R.TGeoCombiTrans(name, bdx * math.cos(math.radians(ang)), bdx * math.sin(math.radians(ang)), 0, R.TGeoRotation("", ang, 0, 0)).RegisterYourself()
    spider = R.AObscuration("obs", R.TGeoCompositeShape("comp", "box2:tr1 + box2:tr2 + box2:tr3 + tube2"))
    top.AddNode(spider, 1, R.TGeoTranslation(0, 0, (0.65 + 40.0) * inch + offset))
    disk2 = R.AGeoAsphericDisk("disk2", (0.65 + 40. + 32.) * inch + offset, -1.049567394559E-002 / inch, (0.65 + 40. + 32. + 0.1) * inch + offset,
                               -1.049567394559E-002 / inch, 12.183 * inch, 4 * inch)
    disk2.SetConicConstants(0.077235, 0.077235)
    top.AddNode(R.AMirror("primary", disk2), 1)
    disk3 = R.AGeoAsphericDisk("disk3", (0.65 + 40. + 32. - 30.86635 - 0.1) * inch + offset, -2.01270013787E-002 / inch,
                               (0.65 + 40. + 32. - 30.86635) * inch + offset, -2.01270013787E-002 / inch, 4.322385947053 * inch, 0)
    top.AddNode(R.AMirror("secondary", disk3), 1)
    origin1 = [0, 0, (0.65 + 40. + 32. - 30.86635 + 50.6706488) * inch + offset + 5 * um]
    top.AddNode(R.AFocalSurface("screen1", R.TGeoBBox("box1", 5 * inch, 5 * inch, 5 * um, origin1)), 1)
    mgr.CloseGeometry()
    mgr.DisableFresnelReflection(disable_fresnel)
    return mgr, [bk7, catalog]


# ----------------------------------------------------------------------------- 5: HexWinstonCone.C:30-73 (mode 0), hex-packed array
def hex_winston_cone(rings=2, coating="multilayer", precalc=False):
    """rings=0 -> 1 cell, 2 -> 19 cells, 10 -> 331 cells.  coating: 'multilayer' (air/SiO2 25.4 nm/Al as in
    unittest_robast.py:256-262), 'ideal' (R=1)."""
    rin, rout = 20 * mm, 10 * mm
    mgr = R.AOpticsManager("manager", "HexWinstonCone")
    world = R.AOpticalComponent("world", R.TGeoBBox("worldbox", 30 * m, 30 * m, 30 * m))
    mgr.SetTopVolume(world)
    rot30 = R.TGeoRotation("rot30", 30, 0, 0)
    rot30.RegisterYourself()
    cone = R.AGeoWinstonCone2D("coneV", rin, rout, rin * 1.733)
    dz = cone.GetDZ()
    pgon = R.TGeoPgon("pgon", 0, 360, 6, 4)
    pgon.DefineSection(0, -dz * 0.999, 0, rout * 1.1)
    pgon.DefineSection(1, -dz * 0.5, 0, rin * 0.9)
    pgon.DefineSection(2, -dz * 0., 0, rin * 0.99)
    pgon.DefineSection(3, dz * 0.999, 0, rin * 1.001)
    R.AGeoWinstonConePoly("hexV", rin, rout, 6)
    cone_mirror = R.AMirror("coneMirror", R.TGeoCompositeShape("coneComp1", "pgon:rot30 - hexV"))
    pgon_pmt = R.TGeoPgon("pgonPMT", 0, 360, 6, 2)
    pgon_pmt.DefineSection(0, -dz - 0.01 * mm, 0, rout * 1.01)
    pgon_pmt.DefineSection(1, -dz, 0, rout * 1.01)
    pmt = R.AFocalSurface("pmt", pgon_pmt)
    keep = []
    if coating == "multilayer":
        air = R.ARefractiveIndex(1., 0.)
        sio2 = R.AFilmetrixDotCom(os.path.join(DATA, "SiO2.nk.txt"))
        al = R.AFilmetrixDotCom(os.path.join(DATA, "Al.nk.txt"))
        layer = R.AMultilayer(air, al)
        layer.InsertLayer(sio2, 25.4 * nm)
        if precalc:
            layer.PreCalculateCoherentTMM(801, 199.5 * nm, 1000.5 * nm, 90, math.radians(-0.5), math.radians(89.5))
        border = R.ABorderSurfaceCondition(world, cone_mirror)
        border.SetMultilayer(layer)
        keep += [air, sio2, al, layer, border]
    # hex packing: with rot30 the flat sides of the outer pgon (apothem rin*1.001) face azimuths 0, 60, 120 ... deg
    pitch = 2 * rin * 1.001 * 1.001
    copy = 0
    for q in range(-rings, rings + 1):
        for r in range(max(-rings, -q - rings), min(rings, -q + rings) + 1):
            # axial hex coordinates; neighbours along the apothem directions (0 and 60 deg)
            x = pitch * (q + r * math.cos(math.radians(60)))
            y = pitch * (r * math.sin(math.radians(60)))
            copy += 1
            world.AddNode(cone_mirror, copy, R.TGeoTranslation(x, y, 0))
            world.AddNode(pmt, copy, R.TGeoCombiTrans(x, y, 0, rot30))
    mgr.CloseGeometry()
    return mgr, keep


# ----------------------------------------------------------------------------- HexOkumuraCone.C:34-83 (mode 1) and a round variant
def okumura_cone(kind="pgon", nz=100):
    """single Bezier-profile light guide: kind='pgon' = hex-hex Okumura cone (AGeoBezierPgon, HexOkumuraCone.C:63-68),
    kind='pcon' = its round counterpart (AGeoBezierPcon); ideal mirror, flat PMT just below the exit aperture."""
    rin, rout = 20 * mm, 10 * mm
    mgr = R.AOpticsManager("manager", "HexOkumuraCone")
    world = R.AOpticalComponent("world", R.TGeoBBox("worldbox", 10 * cm, 10 * cm, 10 * cm))
    mgr.SetTopVolume(world)
    rot30 = R.TGeoRotation("rot30", 30, 0, 0)
    rot30.RegisterYourself()
    dz = R.AGeoWinstonConePoly("hexWin", rin, rout, 6).GetDZ()
    if kind == "pgon":
        outer = R.TGeoPgon("pgon", 0, 360, 6, 4)
        inner = R.AGeoBezierPgon("hexBez", 0, 360, 6, nz, rin, rout, dz)
        expr = "pgon:rot30 - hexBez:rot30"
        pmt_shape = R.TGeoPgon("pgonPMT", 0, 360, 6, 2)
    else:
        outer = R.TGeoPcon("pgon", 0, 360, 4)
        inner = R.AGeoBezierPcon("hexBez", 0, 360, nz, rin, rout, dz)
        expr = "pgon - hexBez"
        pmt_shape = R.TGeoPcon("pgonPMT", 0, 360, 2)
    outer.DefineSection(0, -dz * 0.9999, 0, rout * 1.1)
    outer.DefineSection(1, -dz * 0.5, 0, rin * 0.9)
    outer.DefineSection(2, -dz * 0., 0, rin * 1.001)
    outer.DefineSection(3, dz * 0.9999, 0, rin * 1.001)
    inner.SetControlPoints(0.39, 0.18, 0.87, 0.36)
    cone_mirror = R.AMirror("coneMirror", R.TGeoCompositeShape("coneComp1", expr))
    world.AddNode(cone_mirror, 1)
    pmt_shape.DefineSection(0, -dz - 0.01 * mm, 0, rout * 1.01)
    pmt_shape.DefineSection(1, -dz, 0, rout * 1.01)
    pmt = R.AFocalSurface("pmt", pmt_shape)
    world.AddNode(pmt, 1, rot30)
    mgr.CloseGeometry()
    return mgr, [outer, inner, pmt_shape]


# ----------------------------------------------------------------------------- beams
def beam(cfg, theta_deg=0.0, n_side=None):
    """rbg_shoot_desc parameters (dict) of the configuration's synthetic beam."""
    th = math.radians(theta_deg)
    d = dict(kind=0, nx=1, ny=1, dx=0., dy=0., lambda_min=400 * nm, lambda_max=400 * nm, rot=[1, 0, 0, 0, 1, 0, 0, 0, 1], tr=[0, 0, 0], dir=[0, 0, 1], seed=20180601)
    # SetMagThetaPhi(1, pi - theta, 0)
    tilt = [math.sin(math.pi - th), 0.0, math.cos(math.pi - th)]
    if cfg == 1:
        n = n_side or 1000
        d.update(nx=n, ny=n, dx=5 * m, dy=5 * m, tr=[-6 * m * math.sin(th), 0, 6 * m * math.cos(th)], dir=tilt)
    elif cfg == 2:
        n = n_side or 3334
        kF = 16 * m
        d.update(nx=n, ny=n, dx=14 * m, dy=14 * m, tr=[-2 * kF * math.sin(th), 0, 2 * kF * math.cos(th)], dir=tilt)
    elif cfg == 3:
        n = n_side or 10000
        kZs = 9.980 * m
        d.update(nx=n, ny=n, dx=20 * m, dy=20 * m, tr=[-1.2 * kZs * math.sin(th), 0, 1.2 * kZs * math.cos(th)], dir=tilt)
    elif cfg == 4:
        # RandomCircle(lambda, 12.5 inch, N, rot) with rot = SetAngles(0, theta, 0); lambda uniform in [300, 700] nm
        c, s = math.cos(th), math.sin(th)
        d.update(kind=2, dx=12.5 * inch, lambda_min=300 * nm, lambda_max=700 * nm, rot=[1, 0, 0, 0, c, -s, 0, s, c], dir=[0, 0, 1])
    elif cfg == 5:
        # RandomSquare(400 nm, side, N, rayrot, raytr) with rayrot = (90, 180+deg, 0), HexWinstonCone.C:82-93
        side = n_side or (100 * mm)
        ph, t2 = math.radians(90), math.radians(180 + theta_deg)
        sp, cp, st, ct = math.sin(ph), math.cos(ph), math.sin(t2), math.cos(t2)
        rot = [cp, -ct * sp, st * sp, sp, ct * cp, -st * cp, 0, st, ct]
        d.update(kind=1, dx=side, dy=side, rot=rot, tr=[100 * mm * math.sin(th), 0, 100 * mm * math.cos(th)], dir=[0, 0, 1], seed=20110306)
    else:
        raise ValueError("cfg must be 1..5")
    return d


BUILDERS = {1: simple_parabolic, 2: davies_cotton, 3: schwarzschild_couder, 4: schmidt_cassegrain, 5: hex_winston_cone}
NAMES = {1: "SimpleParabolicTelescope", 2: "DaviesCotton", 3: "SchwarzschildCouder", 4: "SchmidtCassegrain", 5: "HexWinstonCone"}
