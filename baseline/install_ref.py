#!/usr/bin/env python
"""Stage the UNMODIFIED reference (simplify23/MRN) under baseline/_ref/ so that the comparison arm can run where
/root/reference does not exist (the GPU box).

    python baseline/install_ref.py [--src /root/reference]

The reference is plain Python without setup.py / pyproject, so `pip install --target baseline/_ref /root/reference`
has nothing to build (recorded in DESIGN.md §7); the files the stage-1 path imports are copied byte for byte instead.
baseline/_ref/ is git-ignored (never part of the repo's history) but NOT gpurun-ignored, so it travels to the box.
Nothing under mrn_b200/ imports it; only bench.py's reference arms (baseline/ref_arm.py) do.
"""
import argparse
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
KEEP = ("modules", "il_modules", "tools", "data", "config", "test.py", "tiny_train.py", "LICENSE")


def install(src="/root/reference"):
    if not os.path.isfile(os.path.join(src, "modules", "model.py")):
        return None
    os.makedirs(DST, exist_ok=True)
    digest = hashlib.sha256()
    for name in KEEP:
        s, d = os.path.join(src, name), os.path.join(DST, name)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(s):
            shutil.copy2(s, d)
    for root, _, files in sorted(os.walk(DST)):
        for f in sorted(files):
            if f.endswith(".py"):
                digest.update(open(os.path.join(root, f), "rb").read())
    with open(os.path.join(DST, "SOURCE.txt"), "w") as fh:
        fh.write("copied unmodified from %s\nsha256(.py files) %s\n" % (src, digest.hexdigest()))
    return DST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    print(install(ap.parse_args().src))
