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
