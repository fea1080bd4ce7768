#!/usr/bin/env python
"""Build libegaze.so (the C-ABI drop-in boundary, include/egaze.h) for sm_100a.

nvcc cross-compiles without a GPU.  Output lives IN-TREE next to the Python host mirror
(`egocentric-gaze-prediction_b200/egaze/libegaze.so`) so it travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "egaze")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(OUT_DIR, "libegaze.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-Xptxas", "-v" if os.environ.get("EGAZE_PTXAS_V") else "-O3"]
if os.environ.get("EGAZE_MBAR_SPIN"):   # experiment: mbarrier waits poll test_wait instead of try_wait + suspend hint
    FLAGS.append("-DEGAZE_MBAR_SPIN")
if os.environ.get("EGAZE_CONV_PROF"):   # per-role cycle counters in the conv kernel (tools/conv_prof.py); never the shipped build
    FLAGS.append("-DEGAZE_CONV_PROF")


def _sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _stamp(path):
    h = hashlib.sha1()
    for dep in [path] + [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith(".cuh")]:
        with open(dep, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS + ARCH).encode())
    return h.hexdigest()


def _compile(src):
    path = os.path.join(HERE, src)
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(path)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, False, ""
    cmd = [NVCC] + ARCH + FLAGS + ["-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return obj, True, res.stderr


def build(verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    rebuilt = any(r[1] for r in results)
    if verbose:
        for (_, did, log), src in zip(results, srcs):
            if did and log.strip():
                print("== %s\n%s" % (src, log))
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + [r[0] for r in results] + ["-lcudart", "-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
