#!/usr/bin/env python3
"""Recipe: compile the UNMODIFIED reference hot path into oracle/_ref/ (test infrastructure only).

TEST INFRASTRUCTURE -- nothing under oracle/ is imported by the product package
(`nodeembedding-to-communityembedding_b200/`).  Only tests/, __graft_entry__.smoke()/build() and bench.py's
cpu_baseline / --impl reference legs may touch it.

What it does
------------
The reference's SGD hot path is one Cython file, /root/reference/utils/training_sdg_inner.pyx (+ voidptr.h).
This script runs `cython` on that file WHERE IT LIES (no source is copied into the repo), directs the generated C
into oracle/_ref/build/, and compiles it with gcc into

    oracle/_ref/stock/utils/training_sdg_inner.<EXT_SUFFIX>   directives exactly as /root/reference/cython_utils.py:8
    oracle/_ref/tuned/utils/training_sdg_inner.<EXT_SUFFIX>   + legacy_implicit_noexcept=True, -O3 (Cython 3 otherwise
                                                              re-acquires the GIL after every BLAS call inside nogil;
                                                              SURVEY.md section 6) -- the headline CPU baseline.

oracle/_ref/ is git-ignored (it is a build product) but NOT gpurun-ignored, so the two .so files travel to the GPU box,
where /root/reference does not exist.  If /root/reference is absent the script is a no-op (prebuilt files are used).

The reference's own build system (cython_utils.py: distutils `setup(... cythonize ...)`) is NOT run: it writes into the
read-only source tree.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("COMEMB_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PYX = os.path.join(REF, "utils", "training_sdg_inner.pyx")

BASE_DIRECTIVES = ["boundscheck=False", "wraparound=False", "cdivision=True"]  # cython_utils.py:8
VARIANTS = {
    "stock": dict(directives=BASE_DIRECTIVES, cflags=["-O2"]),
    "tuned": dict(directives=BASE_DIRECTIVES + ["legacy_implicit_noexcept=True"],
                  cflags=["-O3", "-march=x86-64-v3"]),
}


def so_path(variant):
    return os.path.join(OUT, variant, "utils", "training_sdg_inner" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(variant, force=False):
    import numpy
    cfg = VARIANTS[variant]
    so = so_path(variant)
    if os.path.exists(so) and not force and os.path.getmtime(so) >= os.path.getmtime(PYX):
        return so
    bdir = os.path.join(OUT, "build", variant)
    os.makedirs(bdir, exist_ok=True)
    os.makedirs(os.path.dirname(so), exist_ok=True)
    c_file = os.path.join(bdir, "training_sdg_inner.c")
    cmd = [sys.executable, "-m", "cython", "-3", "-o", c_file]
    for d in cfg["directives"]:
        cmd += ["-X", d]
    cmd += [PYX]
    subprocess.run(cmd, check=True, cwd=bdir, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    gcc = ["gcc", "-shared", "-fPIC", "-fwrapv", "-fno-strict-aliasing", "-w"] + cfg["cflags"] + [
        "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(),
        "-I", os.path.join(REF, "utils"),  # voidptr.h, read in place
        "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
        c_file, "-o", so, "-lm"]
    subprocess.run(gcc, check=True)
    init = os.path.join(os.path.dirname(so), "__init__.py")
    if not os.path.exists(init):
        open(init, "w").close()
    return so


def main(force=False):
    if not os.path.exists(PYX):
        print("oracle/build_ref.py: %s not present; using prebuilt oracle/_ref if any" % PYX)
        return [so_path(v) for v in VARIANTS if os.path.exists(so_path(v))]
    return [build(v, force) for v in VARIANTS]


if __name__ == "__main__":
    for p in main(force="--force" in sys.argv):
        print(p)
