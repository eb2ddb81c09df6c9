"""Builds alphagpu_b200/libalphagpu.so with nvcc for sm_100a (in-tree, so the .so travels with the repo snapshot).

    python -m alphagpu_b200.build [--force] [-v]
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
SO = os.path.join(HERE, "libalphagpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

SOURCES = ["api.cu", "engine_connect4.cu", "engine_gobang.cu", "engine_hex.cu", "engine_reversi.cu", "nn_tc.cu", "nn_tc512.cu", "train.cu"]
# --fmad=false: the search / fp32-NN arithmetic is specified operation by operation (canonical fp32)
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
         "--fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math", "-Xcompiler", "-ffp-contract=off"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "alphagpu.h"),
                                                                os.path.join(HERE, "..", "include", "alphagpu_train.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    if _stale(obj, _deps()):
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr
    return obj, ""


def build(force: bool = False, verbose: bool = False) -> str:
    global OBJ, SO
    extra = os.environ.get("AGPU_EXTRA_NVCC", "")       # development: e.g. "-DAG_LANES_SHIFT=2" with AGPU_VARIANT=w2
    variant = os.environ.get("AGPU_VARIANT", "")
    if variant:
        OBJ = os.path.join(HERE, "_obj_" + variant)
        SO = os.path.join(HERE, f"libalphagpu_{variant}.so")
    if extra:
        FLAGS.extend(extra.split())
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    if not force and not _stale(SO, _deps()):
        return SO
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [o for o, _ in res]
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
