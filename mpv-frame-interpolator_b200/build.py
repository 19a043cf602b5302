"""In-tree builds: the CUDA C-ABI library (sm_100a) and the C host layer.

Nothing here is JIT: `nvcc` cross-compiles for sm_100a without a GPU, the resulting `.so`
files stay in the tree (git-ignored) and travel to the GPU box with the snapshot.
"""
import os
import pathlib
import shutil
import subprocess

PKG = pathlib.Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
HOSTC = PKG / "mpv" / "video" / "filter" / "HopperRender"

CUDA_LIB = CSRC / "libhopperrender_cuda.so"
OFC_LIB = HOSTC / "libhopperrender_ofc.so"

HOST_CC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",  # no implicit contraction: every fma of the warp arithmetic is written out (hr_warp.cuh mirrors what the NVIDIA OpenCL compiler emits for the reference kernel)
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target: pathlib.Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(pathlib.Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd, cwd=None):
    r = subprocess.run([str(c) for c in cmd], cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(map(str, cmd)), r.stdout, r.stderr))
    return r.stdout + r.stderr


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def build_cuda(force=False, verbose_ptxas=False):
    srcs = [CSRC / "hr_cuda.cu", *sorted(CSRC.glob("*.cuh")), ROOT / "include" / "hopperrender_cuda.h"]
    if not force and _newer(CUDA_LIB, srcs):
        return CUDA_LIB
    flags = list(NVCC_FLAGS)
    if verbose_ptxas:
        flags += ["-Xptxas", "-v"]
    out = _run([nvcc_path(), *flags, "-o", CUDA_LIB, CSRC / "hr_cuda.cu"])
    if verbose_ptxas:
        print(out)
    return CUDA_LIB


def build_host(force=False):
    """The reference-language (C) host: opticalFlowCalc.c forwarding to the C ABI, hrReplay.c (the filter's call
    order as a loop) and hrControl.c. (The mpv-runtime stand-in that runs the reference's own filter source on top of
    this layer is test infrastructure and is built by the test tree, not from here.)"""
    build_cuda()
    srcs = [HOSTC / "opticalFlowCalc.c", HOSTC / "hrReplay.c", HOSTC / "hrControl.c", HOSTC / "hrControl.h", HOSTC / "opticalFlowCalc.h", HOSTC / "config.h"]
    if force or not _newer(OFC_LIB, srcs):
        _run([HOST_CC, "-O2", "-std=c11", "-Wall", "-fPIC", "-shared", "-I", ROOT / "include", "-o", OFC_LIB,
              HOSTC / "opticalFlowCalc.c", HOSTC / "hrReplay.c", HOSTC / "hrControl.c", "-L", CSRC, "-lhopperrender_cuda", "-Wl,-rpath,$ORIGIN/../../../../csrc", "-lm"])
    return OFC_LIB


def build_all(force=False):
    build_cuda(force)
    if (HOSTC / "opticalFlowCalc.c").exists():
        build_host(force)
