#!/bin/bash
# builds the pybind11 host-layer module in-tree (links against librobast_b200.so)
set -e
cd "$(dirname "$0")"
PYINC=$(python -c "import sysconfig;print(sysconfig.get_paths()['include'])")
PBINC=$(python -c "import pybind11;print(pybind11.get_include())")
SUF=$(python -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
g++ -O1 -std=c++17 -fPIC -shared -fvisibility=hidden -I$PYINC -I$PBINC robast_b200/csrc/pybind.cpp -Lrobast_b200 -lrobast_b200 -Wl,-rpath,'$ORIGIN' -o robast_b200/_robast$SUF
