#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — makes the unmodified reference importable on the GPU box.

`/root/reference` only exists in the build container.  This recipe copies the reference's pure-Python package
(`/root/reference/src/pyinterpolate`, nothing is compiled — it has no native code) into `oracle/_ref/`, which is listed
in .gitignore (never committed) but NOT in .gpurunignore, so it travels with the snapshot like the built `.so` files.
`bench.py --impl reference` then times the real `ExperimentalVariogram` / `ordinary_kriging` on BASELINE config 1
next to the C port; the parity fixtures under tests/golden/ were generated from the same sources.
Run by `__graft_entry__.build()` whenever /root/reference is present."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/pyinterpolate"
DST = os.path.join(HERE, "_ref", "pyinterpolate")


def main():
    if not os.path.isdir(SRC):
        print("make_ref: /root/reference is not present here; nothing to do")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    n = sum(len(f) for _, _, f in os.walk(DST))
    print(f"make_ref: {n} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
