"""Small scenes of the reference's unit tests (tutorials/unittest_robast.py) built with the mirror API."""
import robast_b200 as ROOT

cm, mm, um, nm, m = 1.0, 0.1, 1e-4, 1e-7, 100.0


def make_the_world():  # unittest_robast.py:46-54
    manager = ROOT.AOpticsManager("manager", "manager")
    world = ROOT.AOpticalComponent("world", ROOT.TGeoBBox("worldbox", 20 * m, 20 * m, 20 * m))
    manager.SetTopVolume(world)
    return manager


def snell_slab(idx=1.5):  # unittest_robast.py:428-447: focal box nested inside a lens slab
    manager = make_the_world()
    manager.DisableFresnelReflection(True)
    lens = ROOT.ALens("lens", ROOT.TGeoBBox("lensbox", 0.5 * m, 0.5 * m, 1 * mm))
    refidx = ROOT.ARefractiveIndex(idx)
    lens.SetRefractiveIndex(refidx)
    manager.GetTopVolume().AddNode(lens, 1)
    focal = ROOT.AFocalSurface("focal", ROOT.TGeoBBox("focalbox", 0.5 * m, 0.5 * m, 0.1 * mm))
    lens.AddNode(focal, 1)
    manager.CloseGeometry()
    return manager, [refidx]


def sphere_shell_mirror():  # unittest_robast.py:390-400
    manager = make_the_world()
    manager.SetLimit(1000)
    mirror = ROOT.AMirror("mirror", ROOT.TGeoSphere("mirrorsphere", 0.1 * m, 0.2 * m))
    manager.GetTopVolume().AddNode(mirror, 1)
    manager.CloseGeometry()
    return manager, []


def lens_box(refidx, half=0.5 * m):  # unittest_robast.py:67-79, 122-143
    manager = make_the_world()
    lens = ROOT.ALens("lens", ROOT.TGeoBBox("lensbox", half, half, half))
    lens.SetRefractiveIndex(refidx)
    manager.GetTopVolume().AddNode(lens, 1)
    manager.CloseGeometry()
    return manager, lens


def mirror_box_with_border(multilayer=None, sigma=0.0, lambertian=False, reflectance=None):
    """a 45-degree incidence geometry: mirror slab in the x-y plane (unittest_robast.py:186-289, 333-388)"""
    manager = make_the_world()
    mirror = ROOT.AMirror("mirror", ROOT.TGeoBBox("mirrorbox", 0.5 * m, 0.5 * m, 0.5 * m))
    if reflectance is not None:
        mirror.SetReflectance(reflectance)
    world = manager.GetTopVolume()
    world.AddNode(mirror, 1)
    keep = []
    if multilayer is not None or sigma or lambertian:
        # the border belongs to (world -> mirror)
        from robast_b200 import _robast
        border = ROOT.ABorderSurfaceCondition(world, mirror)
        if multilayer is not None:
            border.SetMultilayer(multilayer)
        if sigma:
            border.SetGaussianRoughness(sigma)
        if lambertian:
            border.EnableLambertian(True)
        keep.append(border)
    manager.CloseGeometry()
    return manager, mirror, keep


def focal_box_with_qe(qe_lambda=None, qe_angle=None):  # unittest_robast.py:470-522
    manager = make_the_world()
    focal = ROOT.AFocalSurface("focal", ROOT.TGeoBBox("focalbox", 0.5 * m, 0.5 * m, 1 * mm))
    if qe_lambda is not None:
        focal.SetQuantumEfficiency(qe_lambda)
    if qe_angle is not None:
        focal.SetQuantumEfficiencyAngle(qe_angle)
    manager.GetTopVolume().AddNode(focal, 1)
    manager.CloseGeometry()
    return manager, focal


def hollow_poly(kind, phi1=0., dphi=360., material="mirror"):
    """one general TGeoPcon ('pcon') or TGeoPgon ('pgon') in the world: hollow sections (rmin > 0, like the Haube of
    tutorials/AshraOptics.C:381-384), a radius step, and optionally an azimuthal range"""
    manager = make_the_world()
    if kind == "pcon":
        shape = ROOT.TGeoPcon("poly", phi1, dphi, 5)
    else:
        shape = ROOT.TGeoPgon("poly", phi1, dphi, 5, 5)
    for i, (z, rmin, rmax) in enumerate(((-12., 3., 6.), (-4., 5., 9.), (-4., 5., 11.), (3., 2., 11.), (10., 0., 4.))):
        shape.DefineSection(i, z, rmin, rmax)
    keep = [shape]
    if material == "mirror":
        comp = ROOT.AMirror("polymirror", shape)
    else:  # glass: refraction, Fresnel reflection and total internal reflection at every face
        comp = ROOT.ALens("polylens", shape)
        idx = ROOT.ARefractiveIndex(1.5)
        comp.SetRefractiveIndex(idx)
        keep.append(idx)
    manager.GetTopVolume().AddNode(comp, 1, ROOT.TGeoCombiTrans(1., -2., 3., ROOT.TGeoRotation("polyrot", 20., 35., 10.)))
    manager.CloseGeometry()
    manager.SetLimit(30)
    return manager, keep + [comp]


# outline of the extruded aluminium profile of tutorials/AshraOptics.C:782-789 (12 vertices, concave), in cm
_ASHRA_X = [-5.0, -0.4, -0.4, -5.0, -5.0, 5.0, 5.0, 0.4, 0.4, 5.0, 5.0, -5.0]
_ASHRA_Y = [5.0, 4.2, -4.2, -4.2, -5.0, -5.0, -4.2, -4.2, 4.2, 4.2, 5.0, 5.0]


def arb8_xtru_shape(kind):
    """TGeoArb8 / TGeoXtru solids of the kinds tutorials/AshraOptics.C builds (all lengths in cm, enlarged to tens of cm)"""
    if kind == "arb8_prism":  # trapezoidal prism, equal faces (AGeoUtil::MakeArb8FromPoints output, src/AGeoUtil.cxx:47-82)
        v = [-6., -4., -5., 5., 7., 4., 6., -5.]
        return ROOT.TGeoArb8("arb", 8., v + v)
    if kind == "arb8_twisted":  # AshraOptics.C:403-411 (pip_arb1): two coinciding vertices, non-parallel lower and upper edges
        s = ROOT.TGeoArb8("arb", 10.)
        for i, (x, y) in enumerate(((20., 8.0936), (20., -14.), (20., -14.), (4.66, -14.), (20., -6.7987), (20., -14.), (20., -14.), (15., -14.))):
            s.SetVertex(i, x, y)
        return s
    if kind == "arb8_pyramid":  # AshraOptics.C:264-276 (mir_cut2): the lower face shrunk to (almost) a point
        import math
        s = ROOT.TGeoArb8("arb", 12.)
        for j, r in enumerate((1e-7, 15.)):
            for i, a in enumerate((60., -8.6, 188.6, 120.)):
                s.SetVertex(i + 4 * j, r * math.cos(math.radians(a)), r * math.sin(math.radians(a)))
        return s
    if kind == "arb8_ccw":  # counter-clockwise input: ROOT re-orders it (TGeoArb8::ComputeTwist)
        v = [-6., -4., 6., -5., 7., 4., -5., 5.]
        return ROOT.TGeoArb8("arb", 8., v + [1.2 * c for c in v])
    if kind == "xtru_profile":  # AshraOptics.C:791-795 (30_xtru1)
        s = ROOT.TGeoXtru(2)
        s.SetName("xtru")
        s.DefinePolygon(_ASHRA_X, _ASHRA_Y)
        s.DefineSection(0, -9.)
        s.DefineSection(1, 9.)
        return s
    if kind == "xtru_scaled":  # three sections with offsets and scales, and an outline jump (two sections at one z)
        s = ROOT.TGeoXtru(4)
        s.SetName("xtru")
        s.DefinePolygon(_ASHRA_X, _ASHRA_Y)
        s.DefineSection(0, -9., 0., 0., 1.)
        s.DefineSection(1, -1., 1., -0.5, 0.7)
        s.DefineSection(2, -1., 1., -0.5, 1.2)
        s.DefineSection(3, 8., -1., 0.5, 0.9)
        return s
    raise ValueError(kind)


def arb8_xtru(kind, material="mirror", composite=False):
    """one TGeoArb8 / TGeoXtru (optionally cut by a sphere, like the mirror segments of AshraOptics.C:286-300) in the world"""
    manager = make_the_world()
    shape = arb8_xtru_shape(kind)
    keep = [shape]
    if composite:
        sph = ROOT.TGeoSphere("cutsph", 0., 13.)
        tr = ROOT.TGeoTranslation("cuttr", *((14., -8., 1.) if kind == "arb8_twisted" else (2., 1., -3.)))
        tr.RegisterYourself()
        shape = ROOT.TGeoCompositeShape("cutcomp", "cutsph:cuttr*%s" % shape.GetName())
        keep += [sph, tr, shape]
    if material == "mirror":
        comp = ROOT.AMirror("solidmirror", shape)
    else:
        comp = ROOT.ALens("solidlens", shape)
        idx = ROOT.ARefractiveIndex(1.5)
        comp.SetRefractiveIndex(idx)
        keep.append(idx)
    manager.GetTopVolume().AddNode(comp, 1, ROOT.TGeoCombiTrans(1., -2., 3., ROOT.TGeoRotation("solidrot", 20., 35., 10.)))
    manager.CloseGeometry()
    manager.SetLimit(30)
    return manager, keep + [comp]


def overlapping_frame(nested=False, holder=True, bars=True, stop_ring=True):
    """nodes placed with AddNodeOverlap ("MANY"), as tutorials/AshraOptics.C:1117-1120 uses them: a big AOpticalComponent box `comp`
    holding obscuring bars is laid over the whole optical system; a mirror, a glass plate and a focal plane are ordinary daughters
    of the same mother and share space with it; a ring-shaped stop is a second overlapping node.  nested=True puts a second
    overlapping holder with its own bar inside `comp`'s mother so that two overlapping nodes hold the same points."""
    manager = make_the_world()
    opt = ROOT.AOpticalComponent("opt", ROOT.TGeoBBox("optbox", 3 * m, 3 * m, 3 * m))
    keep = [opt]
    glass = ROOT.ALens("plate", ROOT.TGeoTube("platetube", 0., 40., 1.5))
    idx = ROOT.ARefractiveIndex(1.5)
    glass.SetRefractiveIndex(idx)
    opt.AddNode(glass, 1, ROOT.TGeoTranslation(0., 0., -20.))
    mirror = ROOT.AMirror("dish", ROOT.TGeoSphere("dishsph", 200., 201., 166., 180.))  # concave towards +z, focus near z = 0
    opt.AddNode(mirror, 1, ROOT.TGeoTranslation(0., 0., 100.))
    focal = ROOT.AFocalSurface("focal", ROOT.TGeoTube("focaltube", 0., 6., 0.05))
    opt.AddNode(focal, 1, ROOT.TGeoTranslation(0., 0., 1.))
    comp = ROOT.AOpticalComponent("comp", ROOT.TGeoBBox("compbox", 1.9 * m, 1.9 * m, 1.9 * m))
    for i, (x, y, z) in enumerate(((25., 0., -40.), (-25., 5., 30.), (0., -30., 60.)) if bars else ()):
        bar = ROOT.AObscuration("bar%d" % i, ROOT.TGeoBBox("barbox%d" % i, 2., 60., 2.))
        comp.AddNode(bar, 1, ROOT.TGeoCombiTrans(x, y, z, ROOT.TGeoRotation("barrot%d" % i, 30. * i, 0., 0.)))
        keep.append(bar)
    if holder:
        opt.AddNodeOverlap(comp, 1, ROOT.TGeoCombiTrans(5., -3., 8., ROOT.TGeoRotation("comprot", 10., 5., 0.)))
    stop = ROOT.AObscuration("stop", ROOT.TGeoTube("stoptube", 45., 150., 0.5))
    if stop_ring:
        opt.AddNodeOverlap(stop, 1, ROOT.TGeoTranslation(0., 0., -30.))
    keep += [glass, idx, mirror, focal, comp, stop]
    if nested:
        comp2 = ROOT.AOpticalComponent("comp2", ROOT.TGeoBBox("comp2box", 80., 80., 40.))
        bar = ROOT.AObscuration("bar9", ROOT.TGeoBBox("barbox9", 50., 1.5, 1.5))
        comp2.AddNode(bar, 1, ROOT.TGeoTranslation(0., 12., 5.))
        opt.AddNodeOverlap(comp2, 1, ROOT.TGeoTranslation(0., 0., 40.))
        keep += [comp2, bar]
    manager.GetTopVolume().AddNode(opt, 1, ROOT.TGeoRotation("optrot", 0., 0., 0.))
    manager.CloseGeometry()
    manager.SetLimit(20)
    return manager, keep


# ----------------------------------------------------------------------------- round-2 parity branches
def winston2d(mode, material="mirror"):
    """AGeoWinstonCone2D traced for real (src/AGeoWinstonCone2D.cxx:120-428).
    mode 'solid': one 2-D cone as a solid body (rotated, shifted) — DistFromOutside, ComputeNormal on all five kinds of face,
                  and, as a glass body, DistFromInside + Contains;
    mode 'hex3' : tutorials/HexWinstonCone.C:54-61 (mode 1) — pgon:rot30 - coneV*(coneV:rot60)*(coneV:rot120) with the PMT below."""
    rin, rout = 20 * mm, 10 * mm
    manager = ROOT.AOpticsManager("manager", "Winston2D")
    world = ROOT.AOpticalComponent("world", ROOT.TGeoBBox("worldbox", 30 * cm, 30 * cm, 30 * cm))
    manager.SetTopVolume(world)
    cone = ROOT.AGeoWinstonCone2D("coneV", rin, rout, rin * 1.733)
    dz = cone.GetDZ()
    keep = [cone]
    if mode == "solid":
        if material == "mirror":
            comp = ROOT.AMirror("coneSolid", cone)
        else:
            comp = ROOT.ALens("coneSolid", cone)
            idx = ROOT.ARefractiveIndex(1.5)
            comp.SetRefractiveIndex(idx)
            keep.append(idx)
        world.AddNode(comp, 1, ROOT.TGeoCombiTrans(0.3, -0.2, 0.5, ROOT.TGeoRotation("conerot", 25., 15., 40.)))
        keep.append(comp)
        manager.SetLimit(40)
    else:
        rots = []
        for name, ang in (("rot30", 30), ("rot60", 60), ("rot120", 120)):
            r = ROOT.TGeoRotation(name, ang, 0, 0)
            r.RegisterYourself()
            rots.append(r)
        pgon = ROOT.TGeoPgon("pgon", 0, 360, 6, 4)
        pgon.DefineSection(0, -dz * 0.999, 0, rout * 1.1)
        pgon.DefineSection(1, -dz * 0.5, 0, rin * 0.9)
        pgon.DefineSection(2, -dz * 0., 0, rin * 0.99)
        pgon.DefineSection(3, dz * 0.999, 0, rin * 1.001)
        c1 = ROOT.TGeoCompositeShape("coneComp1", "coneV*(coneV:rot60)*(coneV:rot120)")
        c2 = ROOT.TGeoCompositeShape("coneComp2", "pgon:rot30 - coneComp1")
        mirror = ROOT.AMirror("coneMirror", c2)
        world.AddNode(mirror, 1)
        pmt_shape = ROOT.TGeoPgon("pgonPMT", 0, 360, 6, 2)
        pmt_shape.DefineSection(0, -dz - 0.01 * mm, 0, rout * 1.01)
        pmt_shape.DefineSection(1, -dz, 0, rout * 1.01)
        pmt = ROOT.AFocalSurface("pmt", pmt_shape)
        world.AddNode(pmt, 1, rots[0])
        keep += rots + [pgon, c1, c2, mirror, pmt_shape, pmt]
    manager.CloseGeometry()
    return manager, keep


def th2_mirror():
    """AMirror with a TH2 reflectance R(lambda, angle) (src/AMirror.cxx:39-60; priority TGraph2D > TH2 > TGraph > constant):
    0 at (300 nm, 0), rising along both axes, so that the bilinear interpolation between bin centres and the edge half-bins
    are all visited by a beam spread in wavelength and incidence angle."""
    import math
    h = ROOT.TH2D("refl", "refl", 8, 300 * nm, 500 * nm, 6, 0., math.pi / 2)
    for i in range(1, 9):
        for j in range(1, 7):
            h.SetBinContent(i, j, min(1.0, 0.08 * i + 0.07 * j))
    manager, mirror, keep = mirror_box_with_border(reflectance=h)
    return manager, mirror, keep + [h]


def dispersive_lens(index, half=0.5 * m):
    """a glass cube of the given ARefractiveIndex (Schott / Cauchy / mixed formulas: src/ASchottFormula.cxx:43-55,
    src/ACauchyFormula.cxx:40-46, include/AMixedRefractiveIndex.h:36-45) in the world of the unit tests"""
    manager, lens = lens_box(index, half)
    manager.SetLimit(30)
    return manager, lens
