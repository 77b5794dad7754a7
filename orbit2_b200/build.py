"""Builds orbit2_b200/libo2b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import concurrent.futures as cf
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libo2b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--use_fast_math",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# --use_fast_math would change erff/expf accuracy; we keep IEEE division/sqrt and only take ftz/fmad
FLAGS = [f for f in FLAGS if f != "--use_fast_math"] + os.environ.get("O2_NVCC_DEFS", "").split()


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, obj, log):
    cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "o2b200.h")]
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    hdr_dig = _digest([d for d in deps if not d.endswith(".cu")])
    jobs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        for s in srcs:
            base = os.path.splitext(os.path.basename(s))[0]
            obj = os.path.join(OBJ, base + ".o")
            ostamp = obj + ".stamp"
            od = _digest([s]) + hdr_dig
            if not force and os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read() == od:
                continue
            jobs.append((ex.submit(_compile, s, obj, os.path.join(OBJ, base + ".log")), ostamp, od))
        for fut, ostamp, od in jobs:
            fut.result()
            with open(ostamp, "w") as f:
                f.write(od)
    objs = [os.path.join(OBJ, os.path.splitext(os.path.basename(s))[0] + ".o") for s in srcs]
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
           "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print("built", LIB, file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
