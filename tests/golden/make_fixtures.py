#!/usr/bin/env python
"""Extracts the small material-data fixtures this repo needs from the read-only reference checkout.

Run once in the build container (where /root/reference exists); the outputs are committed so that
nothing reads /root/reference at test/bench time (it does not exist on the GPU box).

  robast_b200/data/nbk7.agf      N-BK7 + N-BK7HT blocks of misc/schottzemax-20180601.agf (Zemax AGF, Schott data)
  robast_b200/data/<X>.nk.txt    tutorials/{Al,Si,Si3N4,SiO2,TiO2}.txt  (filmetrics.com n,k tables)
"""
import os
import shutil

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "robast_b200", "data")


def main():
    os.makedirs(OUT, exist_ok=True)
    lines = open(os.path.join(REF, "misc", "schottzemax-20180601.agf"), encoding="latin-1").read().splitlines()
    keep, on = [], False
    for ln in lines:
        if ln.startswith("NM "):
            on = ln.split()[1] in ("N-BK7", "N-BK7HT", "N-SF6", "F2")
        if on:
            keep.append(ln.rstrip())
    with open(os.path.join(OUT, "nbk7.agf"), "w") as f:
        f.write("CC excerpt of schottzemax-20180601.agf (N-BK7, N-BK7HT, N-SF6, F2)\n" + "\n".join(keep) + "\n")
    for name in ("Al", "Si", "Si3N4", "SiO2", "TiO2"):
        shutil.copyfile(os.path.join(REF, "tutorials", name + ".txt"), os.path.join(OUT, name + ".nk.txt"))


if __name__ == "__main__":
    main()
