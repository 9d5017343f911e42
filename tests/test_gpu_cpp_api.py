"""The C++ mirror API end to end on the GPU: compile examples/parabolic_demo.cpp against include/robast and
librobast_b200.so, run it, and check the closed-form expectations of a parabolic mirror."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_macro_style_program_runs_on_gpu(tmp_path):
    exe = str(tmp_path / "parabolic_demo")
    lib = os.path.join(ROOT, "robast_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include", "robast"), os.path.join(ROOT, "examples", "parabolic_demo.cpp"),
                           "-L", lib, "-lrobast_b200", "-Wl,-rpath," + lib, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    rows = [dict(re.findall(r"(\w+)=([-\d.e]+)", ln)) for ln in out.strip().splitlines()]
    assert len(rows) == 2
    on, off = rows
    assert int(on["focused"]) + int(on["stopped"]) + int(on["exited"]) == 301 * 301
    assert int(on["focused"]) > 20000 and float(on["rms_x"]) < 2e-6 and abs(float(on["mean_x"])) < 1e-8  # on-axis: a point (2e-6 cm step-back quirk)
    assert 5.2 < float(off["mean_x"]) < 6.0 and float(off["rms_x"]) > 1e-3  # 1 deg off-axis: f*tan(theta) = 5.24 cm plus the outward comatic centroid shift
