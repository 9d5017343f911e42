"""The C-ABI library loads on a GPU-less box and exports every symbol include/robast_b200.h declares."""
import ctypes
import os
import re

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT_DIR, "include", "robast_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rbg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(R):
    syms = declared_symbols()
    assert "rbg_trace" in syms and "rbg_scene_create" in syms and len(syms) >= 14
    lib = ctypes.CDLL(R.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    assert sorted(R.ABI_SYMBOLS) == syms
    assert R.rbg_abi_version() == 1


def test_struct_layouts_match_header(R):
    # sizes the C compiler gives (checked against the pybind-built host layer through a round trip)
    assert ctypes.sizeof(R.rbg_trace_opts) == 32
    assert ctypes.sizeof(R.rbg_rays) == 16 + 18 * 8
    assert ctypes.sizeof(R.rbg_shoot_desc) == 16 + 4 * 8 + 15 * 8 + 8
    assert ctypes.sizeof(R.rbg_history) == 8 + 5 * 8


def test_no_gpu_fails_loudly(R):
    if R.rbg_device_count() > 0:
        return
    from robast_b200 import configs
    mgr, _ = configs.simple_parabolic()
    ex = mgr.ExportScene()
    h = ctypes.c_void_p()
    rc = R.rbg_scene_create(ex.desc_ptr(), 0, ctypes.byref(h))
    assert rc != 0 and b"no CPU fallback" in R.rbg_last_error()
    rays = R.ARayShooter.Square(400e-7, 100., 3)
    try:
        mgr.TraceNonSequential(rays)
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("TraceNonSequential must raise without a GPU")


def test_bad_arguments_are_reported_not_thrown(R):
    h = ctypes.c_void_p()
    assert R.rbg_scene_create(None, 0, ctypes.byref(h)) == -1
    assert b"null" in R.rbg_last_error()
    assert R.rbg_trace(None, None, None, None) == -1
    assert R.rbg_scene_destroy(None) == 0
