"""Dump the PTX the NVIDIA OpenCL compiler generates for the reference's .cl kernels (GPU box only).

Purpose: see which float instructions (fma contraction, div.full / div.approx, cvt modes) the
reference's warp kernel really executes on the B200, so the CUDA path can mirror them. Output goes
to gpurun_out/ref_ptx/<kernel>.ptx (scratch; nothing from the reference is committed)."""
import ctypes as C
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_opencl

ok, why = ref_opencl.available()
print(ok, why)
if not ok:
    sys.exit(0)
ocl = ref_opencl._state["ocl"]
ker = C.CDLL(str(ref_opencl.REFDIR / "libhr_ref_kernels.so"))
ker.hr_ref_kernel_source.restype = C.c_char_p
ker.hr_ref_kernel_source.argtypes = [C.c_char_p]
ker.hr_ref_kernel_name.restype = C.c_char_p
ker.hr_ref_kernel_name.argtypes = [C.c_int]

plat = C.c_void_p()
n = C.c_uint()
assert ocl.clGetPlatformIDs(1, C.byref(plat), C.byref(n)) == 0
dev = C.c_void_p()
CL_DEVICE_TYPE_GPU = 1 << 2
ocl.clGetDeviceIDs.argtypes = [C.c_void_p, C.c_ulong, C.c_uint, C.c_void_p, C.c_void_p]
assert ocl.clGetDeviceIDs(plat, CL_DEVICE_TYPE_GPU, 1, C.byref(dev), C.byref(n)) == 0
err = C.c_int()
ocl.clCreateContext.restype = C.c_void_p
ocl.clCreateContext.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
ctx = ocl.clCreateContext(None, 1, C.byref(dev), None, None, C.byref(err))
assert err.value == 0
ocl.clCreateProgramWithSource.restype = C.c_void_p
ocl.clCreateProgramWithSource.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
ocl.clBuildProgram.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
ocl.clGetProgramInfo.argtypes = [C.c_void_p, C.c_uint, C.c_size_t, C.c_void_p, C.c_void_p]
out = ROOT / "gpurun_out" / "ref_ptx"
out.mkdir(parents=True, exist_ok=True)
for i in range(ker.hr_ref_kernel_count()):
    name = ker.hr_ref_kernel_name(i).decode()
    src = ker.hr_ref_kernel_source(name.encode())
    p = C.c_char_p(src)
    prog = ocl.clCreateProgramWithSource(ctx, 1, C.byref(p), None, C.byref(err))
    assert err.value == 0
    rc = ocl.clBuildProgram(prog, 1, C.byref(dev), b"", None, None)   # the reference passes no options (opticalFlowCalc.c:83)
    print(name, "build rc", rc)
    sz = C.c_size_t()
    CL_PROGRAM_BINARY_SIZES, CL_PROGRAM_BINARIES = 0x1165, 0x1166
    assert ocl.clGetProgramInfo(prog, CL_PROGRAM_BINARY_SIZES, C.sizeof(sz), C.byref(sz), None) == 0
    buf = C.create_string_buffer(sz.value)
    ptr = C.c_void_p(C.addressof(buf))
    assert ocl.clGetProgramInfo(prog, CL_PROGRAM_BINARIES, C.sizeof(ptr), C.byref(ptr), None) == 0
    (out / (name + ".ptx")).write_bytes(buf.raw)
    print(name, sz.value, "bytes")
