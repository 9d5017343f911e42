"""include/robast/RootExporter.h — the exporter for a host with real CERN ROOT and the unmodified reference classes — parses and
type-checks against stand-in declarations of the ROOT / ROBAST API it uses (tests/fake_root/).  ROOT itself is not in this image;
the same export logic, written against the ROOT-free mirror classes (ASceneExport in Robast.h), is what every GPU test runs."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_root_exporter_parses_against_standin_root_headers(tmp_path):
    src = tmp_path / "use_exporter.cpp"
    src.write_text('#include <algorithm>\n#include "robast/RootExporter.h"\n'
                   'int use(AOpticsManager* m, ARayArray* a) { robast_b200::GpuTracer g(m); g.TraceNonSequential(*a); return (int)sizeof(robast_b200::RootExporter); }\n')
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-DROBAST_HAVE_ROOT", "-I", os.path.join(ROOT, "tests", "fake_root"), "-I", os.path.join(ROOT, "include"), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_root_exporter_covers_every_abi_shape_and_index_kind():
    text = open(os.path.join(ROOT, "include", "robast", "RootExporter.h")).read()
    header = open(os.path.join(ROOT, "include", "robast_b200.h")).read()
    for name in re.findall(r"\b(RBG_SHAPE_[A-Z0-9]+)\s*=", header) + re.findall(r"\b(RBG_INDEX_[A-Z]+)\s*=", header):
        assert name in text, name + " is not exported"
    for cls in ("ALens", "AMirror", "AFocalSurface", "AObscuration", "ABorderSurfaceCondition", "AMultilayer"):
        assert cls in text


def test_header_is_inert_without_root(tmp_path):
    src = tmp_path / "no_root.cpp"
    src.write_text('#include "robast/RootExporter.h"\nint main() { return 0; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
