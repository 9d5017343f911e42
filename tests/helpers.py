"""Test helpers: load the CPU oracle (checker) and the host emulation of the device code, build ray
batches, and run a trace through a chosen backend on the same flat scene description."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT_DIR, "oracle", "_build", "liboracle.so")
EMUL_SO = os.path.join(ROOT_DIR, "tests", "_build", "libemul.so")


def _build(cmd, out, srcs):
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(cmd, cwd=ROOT_DIR)


def load_oracle():
    srcs = [os.path.join(ROOT_DIR, "oracle", "oracle.cpp"), os.path.join(ROOT_DIR, "include", "robast_b200.h")]
    _build(["make", "-s", "-C", "oracle"], ORACLE_SO, srcs)
    import robast_b200 as R
    lib = C.CDLL(ORACLE_SO)
    lib.orc_trace.restype = C.c_int
    lib.orc_trace.argtypes = [C.c_void_p, C.POINTER(R.rbg_trace_opts), C.POINTER(R.rbg_rays), C.c_int]
    lib.orc_trace_history.restype = C.c_int
    lib.orc_trace_history.argtypes = [C.c_void_p, C.POINTER(R.rbg_trace_opts), C.POINTER(R.rbg_rays), C.POINTER(R.rbg_history), C.c_int]
    lib.orc_tmm_general.restype = C.c_int
    lib.orc_tmm_general.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.orc_tmm.restype = C.c_int
    lib.orc_tmm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for name in ("orc_index_n", "orc_index_k", "orc_index_abslen", "orc_graph_eval"):
        f = getattr(lib, name)
        f.restype = C.c_double
        f.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.orc_th2_interp.restype = C.c_double
    lib.orc_th2_interp.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    lib.orc_graph2d_interp.restype = C.c_double
    lib.orc_graph2d_interp.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
    lib.orc_containment_radius.restype = C.c_int
    lib.orc_containment_radius.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_double, C.c_void_p]
    lib.orc_uniform.restype = C.c_double
    lib.orc_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    lib.orc_shoot.restype = C.c_int
    lib.orc_shoot.argtypes = [C.POINTER(R.rbg_shoot_desc), C.c_int64, C.c_int64] + [C.c_void_p] * 8
    lib.orc_shoot_bunches.restype = C.c_int
    lib.orc_shoot_bunches.argtypes = [C.POINTER(R.rbg_bunches), C.c_int64, C.c_int64] + [C.c_void_p] * 8
    lib.orc_set_voxels.restype = C.c_int
    lib.orc_set_voxels.argtypes = [C.c_int]
    lib.orc_shape_contains.restype = C.c_int
    lib.orc_shape_contains.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.orc_shape_dist.restype = C.c_double
    lib.orc_shape_dist.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.orc_shape_normal.restype = C.c_int
    lib.orc_shape_normal.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def load_emul():
    srcs = [os.path.join(ROOT_DIR, "tests", "emul", "emul.cpp")] + [os.path.join(ROOT_DIR, "robast_b200", "csrc", f)
                                                                    for f in ("rb_device.cuh", "rb_build.h", "rb_scene.h")]
    _build(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wl,-Bsymbolic", "-o", EMUL_SO, "tests/emul/emul.cpp"], EMUL_SO, srcs)
    import robast_b200 as R
    lib = C.CDLL(EMUL_SO)
    lib.emul_trace.restype = C.c_int
    lib.emul_trace.argtypes = [C.c_void_p, C.POINTER(R.rbg_trace_opts), C.POINTER(R.rbg_rays), C.c_int]
    lib.emul_trace_history.restype = C.c_int
    lib.emul_trace_history.argtypes = [C.c_void_p, C.POINTER(R.rbg_trace_opts), C.POINTER(R.rbg_rays), C.POINTER(R.rbg_history), C.c_int]
    lib.emul_div.restype = C.c_double
    lib.emul_div.argtypes = [C.c_double, C.c_double]
    lib.emul_tmm.restype = C.c_int
    lib.emul_tmm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return lib


def shoot_desc(params):
    import robast_b200 as R
    d = R.rbg_shoot_desc()
    for k in ("kind", "nx", "ny", "dx", "dy", "lambda_min", "lambda_max", "seed"):
        setattr(d, k, params[k])
    for i in range(9):
        d.rot[i] = params["rot"][i]
    for i in range(3):
        d.tr[i] = params["tr"][i]
        d.dir[i] = params["dir"][i]
    return d


class Rays:
    """Host SoA batch: inputs (n,8) columns x,y,z,t,dx,dy,dz,lambda and separate outputs."""

    def __init__(self, inp):
        inp = np.ascontiguousarray(np.asarray(inp, dtype=np.float64).T)  # (8, n)
        self.inp = inp
        self.n = inp.shape[1]
        self.out = np.zeros((7, self.n))
        self.iout = np.zeros((3, self.n), dtype=np.int32)

    def struct(self):
        import robast_b200 as R
        r = R.rbg_rays()
        r.n = self.n
        r.on_device = 0
        names = ["x", "y", "z", "t", "dx", "dy", "dz", "lambda_"]
        for i, k in enumerate(names):
            setattr(r, k, self.inp[i].ctypes.data)
        for i, k in enumerate(["ox", "oy", "oz", "ot", "odx", "ody", "odz"]):
            setattr(r, k, self.out[i].ctypes.data)
        for i, k in enumerate(["status", "last_node", "npoints"]):
            setattr(r, k, self.iout[i].ctypes.data)
        return r

    pos = property(lambda s: s.out[0:3].T)
    time = property(lambda s: s.out[3])
    dirs = property(lambda s: s.out[4:7].T)
    status = property(lambda s: s.iout[0])
    last_node = property(lambda s: s.iout[1])
    npoints = property(lambda s: s.iout[2])


class History:
    """polyline record in the layout of rbg_history: pts[a, k, i] = coordinate a (x,y,z,t) of point k of ray i"""

    def __init__(self, n, depth):
        self.n, self.depth = n, depth
        self.pts = np.full((4, depth, n), np.nan)
        self.node = np.full((depth, n), -99, dtype=np.int32)

    def struct(self):
        import robast_b200 as R
        h = R.rbg_history()
        h.max_points = self.depth
        for a, k in enumerate(["hx", "hy", "hz", "ht"]):
            setattr(h, k, self.pts[a].ctypes.data)
        h.hnode = self.node.ctypes.data
        return h


def compare_history(ha, hb, npoints, tol_pos=1e-7, tol_time=1e-7 / 2.99792458e10):
    """recorded points k < min(npoints, depth) of every ray agree (positions, times, node ids)"""
    bad = 0
    for k in range(ha.depth):
        m = npoints > k
        if not m.any():
            break
        dp = np.linalg.norm(ha.pts[:3, k, m] - hb.pts[:3, k, m], axis=0)
        dt = np.abs(ha.pts[3, k, m] - hb.pts[3, k, m])
        bad += int(((dp > tol_pos) | (dt > tol_time + 1e-12 * np.abs(ha.pts[3, k, m])) | (ha.node[k, m] != hb.node[k, m])).sum())
    return bad


def trace_history_with(fn, export, rays, o, depth, nthreads=1):
    h = History(rays.n, depth)
    r, hs = rays.struct(), h.struct()
    rc = fn(export.desc_ptr(), C.byref(o), C.byref(r), C.byref(hs), nthreads)
    assert rc == 0, "trace backend returned %d" % rc
    return h


def trace_history_gpu(export, rays, o, depth, device=0):
    import robast_b200 as R
    hnd = C.c_void_p()
    R.check(R.rbg_scene_create(export.desc_ptr(), device, C.byref(hnd)))
    h = History(rays.n, depth)
    try:
        r, hs = rays.struct(), h.struct()
        R.check(R.rbg_trace_history(hnd, C.byref(o), C.byref(r), C.byref(hs), None))
    finally:
        R.rbg_scene_destroy(hnd)
    return h


def psf_histogram(x, y, nx, xmin, xmax, ny, ymin, ymax):
    """TH2::Fill of the points: bins[i + nx*j] (double) and the in-range statistics {sum w, x, y, x^2, y^2}"""
    m = (x >= xmin) & (x < xmax) & (y >= ymin) & (y < ymax)
    xs, ys = x[m], y[m]
    bx = np.minimum((nx * (xs - xmin) / (xmax - xmin)).astype(np.int64), nx - 1)
    by = np.minimum((ny * (ys - ymin) / (ymax - ymin)).astype(np.int64), ny - 1)
    bins = np.bincount(bx + nx * by, minlength=nx * ny).astype(np.float64)
    stats = np.array([m.sum(), xs.sum(), ys.sum(), (xs ** 2).sum(), (ys ** 2).sum()], dtype=np.float64)
    return bins, stats


def oracle_containment(oracle, bins, stats, nx, xmin, xmax, ny, ymin, ymax, fraction):
    out = np.zeros(3)
    assert oracle.orc_containment_radius(bins.ctypes.data, nx, xmin, xmax, ny, ymin, ymax, stats.ctypes.data, fraction, out.ctypes.data) == 0
    return out


def make_rays(oracle, params, first, n):
    """generate a beam with the oracle's ARayShooter restatement (same Philox stream as rbg_shoot)"""
    a = np.zeros((8, n))
    d = shoot_desc(params)
    rc = oracle.orc_shoot(C.byref(d), first, n, *[a[i].ctypes.data for i in range(8)])
    assert rc == 0
    return Rays(a.T)


def opts(limit=100, disable_fresnel=0, quirks=3, steps_per_launch=0, seed=1234, ray_id_offset=0):
    import robast_b200 as R
    o = R.rbg_trace_opts()
    o.limit, o.disable_fresnel, o.quirks, o.steps_per_launch, o.seed, o.ray_id_offset = limit, disable_fresnel, quirks, steps_per_launch, seed, ray_id_offset
    return o


def trace_with(fn, export, rays, o, nthreads=1):
    r = rays.struct()
    rc = fn(export.desc_ptr(), C.byref(o), C.byref(r), nthreads)
    assert rc == 0, "trace backend returned %d" % rc
    return rays


def trace_gpu(export, rays, o, device=0):
    """through the C ABI: rbg_scene_create + rbg_trace with host pointers"""
    import robast_b200 as R
    h = C.c_void_p()
    R.check(R.rbg_scene_create(export.desc_ptr(), device, C.byref(h)))
    try:
        r = rays.struct()
        R.check(R.rbg_trace(h, C.byref(o), C.byref(r), None))
    finally:
        R.rbg_scene_destroy(h)
    return rays


def compare(a, b, tol_pos=1e-7, tol_dir=1e-9, tol_time=1e-7 / 2.99792458e10):
    """per-ray parity report between two traced batches"""
    same_status = a.status == b.status
    dpos = np.linalg.norm(a.pos - b.pos, axis=1)
    cosang = np.clip(np.sum(a.dirs * b.dirs, axis=1), -1, 1)
    cross = np.linalg.norm(np.cross(a.dirs, b.dirs), axis=1)
    dang = np.arctan2(cross, cosang)
    dt = np.abs(a.time - b.time)
    ok = same_status & (dpos <= tol_pos) & (dang <= tol_dir) & (dt <= tol_time + 1e-12 * np.abs(a.time))
    return dict(n=a.n, status_mismatch=int((~same_status).sum()), max_dpos=float(dpos[same_status].max() if same_status.any() else 0),
                max_dang=float(dang[same_status].max() if same_status.any() else 0), max_dt=float(dt[same_status].max() if same_status.any() else 0),
                npoints_mismatch=int((a.npoints != b.npoints).sum()), node_mismatch=int((a.last_node != b.last_node).sum()), bad=int((~ok).sum()))


def make_bunches(n_bunch, seed, z=1000., telescope_z=250., refidx=1.00027, lam_min=300., lam_max=600.):
    """synthetic CORSIKA photon bunches of one telescope (the columns ACorsikaIACTFile keeps, src/ACorsikaIACTFile.cxx:86-97):
    positions over a 12 m dish, near-vertical directions, bunch sizes around 1 with fractional parts, a third of the bunches with
    undetermined wavelength (lambda = 0).  Returns (rbg_bunches, dict of the float32 arrays kept alive)."""
    import robast_b200 as R
    rng = np.random.default_rng(seed)
    a = {}
    a["x"] = (rng.random(n_bunch) * 1200. - 600.).astype(np.float32)
    a["y"] = (rng.random(n_bunch) * 1200. - 600.).astype(np.float32)
    a["time"] = (rng.random(n_bunch) * 50.).astype(np.float32)
    th, ph = np.radians(rng.random(n_bunch) * 3.), rng.random(n_bunch) * 2 * np.pi
    a["cx"] = (np.sin(th) * np.cos(ph)).astype(np.float32)
    a["cy"] = (np.sin(th) * np.sin(ph)).astype(np.float32)
    a["cz"] = (-np.cos(th)).astype(np.float32)
    lam = 300. + 300. * rng.random(n_bunch)
    lam[rng.random(n_bunch) < 0.33] = 0.
    a["lambda_"] = lam.astype(np.float32)
    ph_ = rng.random(n_bunch) * 3.2
    ph_[rng.random(n_bunch) < 0.1] = 0.
    ph_[::7] = 2.0
    a["photons"] = ph_.astype(np.float32)
    b = R.rbg_bunches()
    b.nbunches = n_bunch
    for k, v in a.items():
        setattr(b, k, v.ctypes.data_as(C.POINTER(C.c_float)))
    b.z, b.telescope_z, b.refractive_index, b.lambda_min_nm, b.lambda_max_nm, b.seed = z, telescope_z, refidx, lam_min, lam_max, 20110306
    return b, a
