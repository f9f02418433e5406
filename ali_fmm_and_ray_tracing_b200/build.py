"""Builds libalifmm.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libalifmm.so")
SOURCES = ["alifmm.cu"]
HEADERS = ["ali_core.cuh", "ali_seq.cuh", "ali_band.cuh", "ali_ray.cuh", "ali_glibcmath.cuh", "ali_glxmath.cuh", "ali_strip.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # the reference never fuses a*b+c; keep its roundings
    "--shared", "-Xcompiler", "-fPIC",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libalifmm.so cannot be built")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "alifmm.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compiles csrc/alifmm.cu into libalifmm.so next to this file."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    cmd += ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
