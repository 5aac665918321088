"""Build libjitr_b200.so (sm_100a) in-tree with nvcc; sources under jitr_b200/csrc.

    python -m jitr_b200._build [--force]

All translation units are compiled in parallel; objects are cached by content hash.
"""
from __future__ import annotations

import hashlib
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
# JITR_B200_VARIANT=<name> + JITR_B200_NVCC_FLAGS="-D..." build a tuning variant into build/variants/<name>/
VARIANT = os.environ.get("JITR_B200_VARIANT", "")
LIBDIR = os.path.join(ROOT, "build", "variants", VARIANT) if VARIANT else os.path.join(PKG, "_lib")
OBJDIR = os.path.join(ROOT, "build", "obj_" + VARIANT if VARIANT else "obj")
LIB = os.path.join(LIBDIR, "libjitr_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC,
] + os.environ.get("JITR_B200_NVCC_FLAGS", "").split()


def sweep_modes():
    txt = open(os.path.join(CSRC, "sweep_decl.cuh")).read()
    m = re.search(r"#define JITR_SWEEP_MODES\(X\)(.*)", txt)
    return [int(v) for v in re.findall(r"X\((-?\d+)\)", m.group(1))]


def _units():
    units = [(name, os.path.join(CSRC, name + ".cu"), []) for name in ("api", "probe", "elastic", "general", "packed_solve", "nonlocal_gen", "quasielastic", "uq")]
    for mode in sweep_modes():
        units.append((f"sweep_{mode}".replace("-", "g"), os.path.join(CSRC, "sweep_inst.cu"), [f"-DJITR_SWEEP_MODE={mode}"]))
    return units


def _deps(src, seen=None):
    """Transitive closure of the quoted #includes of `src` (searched in csrc/ and include/)."""
    seen = set() if seen is None else seen
    if src in seen:
        return seen
    seen.add(src)
    for inc in re.findall(r'#include "([^"]+)"', open(src).read()):
        for root in (CSRC, INCLUDE):
            cand = os.path.join(root, inc)
            if os.path.exists(cand):
                _deps(cand, seen)
    return seen


def _digest(src, extra):
    h = hashlib.sha256()
    for fn in sorted(_deps(src)):
        h.update(fn.encode())
        h.update(open(fn, "rb").read())
    h.update(" ".join(FLAGS + extra).encode())
    return h.hexdigest()


def _compile(unit):
    name, src, extra = unit
    obj = os.path.join(OBJDIR, name + ".o")
    stamp = obj + ".sha"
    dig = _digest(src, extra)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = [NVCC] + FLAGS + extra + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(dig)
    return obj, True


def build(force=False, verbose=True):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for fn in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, fn))
    units = _units()
    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        res = list(ex.map(_compile, units))
    objs = [o for o, _ in res]
    if any(c for _, c in res) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[jitr_b200] built {LIB} ({sum(c for _, c in res)} units recompiled)")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
