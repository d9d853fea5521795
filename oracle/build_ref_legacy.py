#!/usr/bin/env python3
"""Recipe: rebuild and compile the reference's LEGACY fused kernel `train_sg` (test infrastructure only).

The shipped /root/reference/utils/training_sdg_inner.c was generated (Cython 0.24.1) from an OLDER .pyx than the one
in the tree; that older source exports the fused pass `train_sg` (o3 gradient + SGNS pair per window pair) which the
north star calls "the fused o1+o2+o3 pass".  The .c cannot be compiled for CPython 3.12 (it includes longintrepr.h),
but Cython embeds the .pyx source in comments: every `/* "utils/training_sdg_inner.pyx":N` block shows ~5 source
lines around line N.  This script reads those blocks IN PLACE, reassembles the old .pyx line by line (345 of 430
lines), fills the uncovered lines -- all of them declaration boilerplate (argument-list continuations whose content
is fixed by the C prototypes at .c:1597 and .c:2520, two DEFs, three cdef lines, a docstring tail) -- and compiles
the result with Cython 3 into  oracle/_ref/legacy/utils/training_sdg_inner.<EXT_SUFFIX>  (git-ignored build product).

Nothing of the reference is copied into the repository; the reconstructed .pyx lives only under oracle/_ref/.
"""
import os
import re
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("COMEMB_REFERENCE", "/root/reference")
STALE_C = os.path.join(REF, "utils", "training_sdg_inner.c")
OUT = os.path.join(HERE, "_ref", "legacy")

SG_PARAMS = ["unsigned long long table_len,", "REAL_t *node_embedding,", "REAL_t *negative_embedding,",
             "const int size,", "const np.uint32_t word_index,", "const np.uint32_t word2_index,",
             "const REAL_t alpha,", "const REAL_t _lambda,", "REAL_t *work,", "unsigned long long next_random,"]
COM_PARAMS = ["REAL_t *inv_covariance_mat,", "REAL_t *pi,", "const int k,", "const REAL_t alpha,",
              "const REAL_t lambda2,", "const int size,", "const np.uint32_t word2_index,", "REAL_t * work,",
              "REAL_t * work1,"]


def fillers():
    f = {12: "    void* PyCObject_AsVoidPtr(object obj)"}
    for i, p in enumerate(SG_PARAMS + ["const int is_node_embedding", ") nogil"]):
        f[33 + i] = "    " + p
    for i, p in enumerate(COM_PARAMS):
        f[50 + i] = "        " + p
    f.update({74: "cdef fast_community_sdg_ptr fast_community_sdg", 75: "", 76: "DEF EXP_TABLE_SIZE = 1000",
              77: "DEF MAX_EXP = 6", 78: ""})
    for base in (92, 151):
        for i, p in enumerate(SG_PARAMS + ["const int is_node_embedding) nogil:"]):
            f[base + i] = "        " + p
    for base in (208, 259):
        for i, p in enumerate(COM_PARAMS):
            f[base + i] = "        " + p
    for base in (242, 293):
        f[base], f[base + 1], f[base + 2] = "                work, &size,", "                &ONEF,", \
            "                work1, &size"
    for n in (139, 140, 141, 142):
        f[n] = ""
    f.update({315: "    cdef REAL_t *work1_o3", 316: "    cdef REAL_t *work2_o3",
              332: "    cdef np.uint32_t indexes[MAX_SENTENCE_LEN]",
              399: "    into table EXP_TABLE.", 400: '    """', 401: "    global fast_context_sg_neg",
              402: "    global fast_community_sdg", 403: ""})
    return f


def reconstruct():
    src = open(STALE_C, encoding="utf-8", errors="replace").read()
    lines = {}
    for m in re.finditer(r'/\* "utils/training_sdg_inner\.pyx":(\d+)\n(.*?)\*/', src, re.S):
        n = int(m.group(1))
        body = [b[3:] if b.startswith(" * ") else (b[2:] if b.startswith(" *") else b)
                for b in m.group(2).split("\n")]
        mark = [i for i, b in enumerate(body) if "# <<<<<<<<<<<<<<" in b]
        if not mark:
            continue
        for i, b in enumerate(body[:-1]):
            lines.setdefault(n + (i - mark[0]), b.replace("             # <<<<<<<<<<<<<<", ""))
    fill = fillers()
    last = max(lines)
    out = []
    for n in range(1, last + 1):
        if n in lines:
            out.append(lines[n])
        elif n in fill:
            out.append(fill[n])
        else:
            raise RuntimeError("legacy pyx line %d neither embedded in the .c nor a known declaration line" % n)
    return "\n".join(out) + "\n"


def so_path():
    return os.path.join(OUT, "utils", "training_sdg_inner" + sysconfig.get_config_var("EXT_SUFFIX"))


def main(force=False):
    if not os.path.exists(STALE_C):
        print("oracle/build_ref_legacy.py: %s not present; using prebuilt oracle/_ref/legacy if any" % STALE_C)
        return so_path() if os.path.exists(so_path()) else None
    so = so_path()
    if os.path.exists(so) and not force:
        return so
    import numpy
    bdir = os.path.join(OUT, "build")
    os.makedirs(bdir, exist_ok=True)
    os.makedirs(os.path.dirname(so), exist_ok=True)
    pyx = os.path.join(bdir, "training_sdg_inner.pyx")
    with open(pyx, "w") as f:
        f.write(reconstruct())
    c_file = os.path.join(bdir, "training_sdg_inner.c")
    cmd = [sys.executable, "-m", "cython", "-3", "-o", c_file, "-I", os.path.join(REF, "utils")]
    for d in ("boundscheck=False", "wraparound=False", "cdivision=True", "legacy_implicit_noexcept=True"):
        cmd += ["-X", d]
    subprocess.run(cmd + [pyx], check=True, cwd=bdir, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run(["gcc", "-shared", "-fPIC", "-fwrapv", "-fno-strict-aliasing", "-w", "-O2",
                    "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(), "-I", os.path.join(REF, "utils"),
                    "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", c_file, "-o", so, "-lm"], check=True)
    open(os.path.join(os.path.dirname(so), "__init__.py"), "a").close()
    return so


if __name__ == "__main__":
    print(main(force="--force" in sys.argv))
