"""Drop-in check of the C++ mirror headers: the reference's own tutorial macros compile UNMODIFIED against
include/robast/ (ROOT-free).  Needs the read-only reference checkout, so it only runs in the build container."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/tutorials"
MACROS = ["SimpleParabolicTelescope", "DaviesCotton", "SchwarzschildCouder", "HexWinstonCone", "SchmidtCassegrain",
          "HESS1", "MST", "HexOkumuraCone", "AbsLengthTest", "EdmundOptics", "multilayer", "multithread", "AshraOptics"]
# not covered: CORSIKA.C (ACorsikaIACTFile), Optimize.C / optimize_multilayer.C (MINUIT),
# SellmeierFit.C (TFile / TF1 fitting)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not available (GPU box)")
@pytest.mark.parametrize("macro", MACROS)
def test_reference_macro_compiles_against_mirror(tmp_path, macro):
    src = tmp_path / (macro + "_main.cpp")
    # cling macros rely on ROOT's implicit includes; SchmidtCassegrain.C also includes a file that is not in the repo
    (tmp_path / "oxon.C").write_text("// placeholder for the missing tutorials/oxon.C\n")
    src.write_text('#include "Robast.h"\n#include "%s/%s.C"\nint main() { return 0; }\n' % (REF, macro))
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include", "robast"), "-I", os.path.join(ROOT, "include", "robast", "compat"),
           "-I", str(tmp_path), str(src)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
