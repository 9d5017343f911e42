"""Builds the reference's tutorial macros, unmodified, against the mirror headers into tests/_build/tutorials/ (git-ignored).
Runs only where the reference checkout is mounted (the build container); the sources are compiled from where they lie and are
never copied into this repository.  The executables travel to the GPU box, where tests/test_gpu_tutorials.py runs them."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/tutorials"
OUT = os.path.join(ROOT, "tests", "_build", "tutorials")
MACROS = {"SimpleParabolicTelescope": "SimpleParabolicTelescope()", "DaviesCotton": "DaviesCotton()", "HESS1": "HESS1()", "MST": "MST()",
          "SchwarzschildCouder": "SchwarzschildCouder()", "AshraOptics": "AshraOptics()", "HexWinstonCone": "HexWinstonCone()",
          "HexOkumuraCone": "HexOkumuraCone()", "SchmidtCassegrain": "SchmidtCassegrain()", "AbsLengthTest": "AbsLengthTest()",
          "EdmundOptics": "EdmundOptics()", "multilayer": "multilayer()"}


def build(verbose=False):
    if not os.path.isdir(REF):
        return []
    os.makedirs(OUT, exist_ok=True)
    built = []
    lib = os.path.join(ROOT, "robast_b200")
    deps = [os.path.join(ROOT, "include", "robast", f) for f in ("Robast.h", "RootCompat.h")] + [os.path.join(lib, "librobast_b200.so")]
    with tempfile.TemporaryDirectory() as tmp:
        for macro, call in MACROS.items():
            exe = os.path.join(OUT, macro)
            srcs = deps + [os.path.join(REF, macro + ".C")]
            if os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in srcs):
                built.append(exe)
                continue
            main = os.path.join(tmp, macro + "_main.cpp")
            with open(os.path.join(tmp, "oxon.C"), "w") as f:  # SchmidtCassegrain.C includes a file that is not in the reference repo
                f.write("// placeholder for the missing tutorials/oxon.C\n")
            with open(main, "w") as f:
                f.write('#include "Robast.h"\n#include "%s/%s.C"\nint main() { %s; return 0; }\n' % (REF, macro, call))
            cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include", "robast"), "-I", os.path.join(ROOT, "include", "robast", "compat"), "-I", tmp, main,
                   "-L", lib, "-lrobast_b200", "-Wl,-rpath,$ORIGIN/../../../robast_b200", "-o", exe]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            built.append(exe)
    return built


if __name__ == "__main__":
    print("\n".join(build(verbose="-v" in sys.argv)))
