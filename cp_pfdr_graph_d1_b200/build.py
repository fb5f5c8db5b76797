"""Builds lib/libpfdr_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--threads", "2"]


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src, *extra))


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "pfdr_b200.h"))
    cus = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs, objs = [], []
    for f in cus:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJDIR, f[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, obj, headers):
            jobs.append([NVCC, *FLAGS, "-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    out = os.path.join(LIBDIR, "libpfdr_b200.so")
    if jobs or not os.path.exists(out):
        run([NVCC, "-Wno-deprecated-gpu-targets", "-shared", "-o", out, *objs])
    # drop-in C++ shim: the reference's template symbols on top of the C ABI
    shim_srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.startswith("shim_") and f.endswith(".cpp"))
    shim = os.path.join(LIBDIR, "libpfdr_b200_shim.so")
    if shim_srcs and (force or any(_newer(s, shim, headers) for s in shim_srcs)):
        run(["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-o", shim, *shim_srcs,
             "-L" + LIBDIR, "-lpfdr_b200", "-Wl,-rpath,$ORIGIN"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
