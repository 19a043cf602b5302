#!/usr/bin/env python
"""Recipe: build the REFERENCE ITSELF from the sources where they lie under /root/reference.

Outputs go only into oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun like any
other built .so). No reference source is copied into the repository.

  oracle/_ref/libhr_ref_ofc.so      video/filter/HopperRender/opticalFlowCalc.c compiled UNMODIFIED
                                    (gcc, against the 70-line oracle/cl_shim/CL/cl.h because the
                                    image has the OpenCL ICD loader but no OpenCL headers) and
                                    linked to the CUDA toolkit's libOpenCL.so.1. It exports the
                                    reference's own initOpticalFlowCalc / updateFrame /
                                    calculateOpticalFlow / warpFrames / downloadFrame / freeOFC.
  oracle/_ref/libhr_ref_kernels.so  the five reference .cl kernel sources embedded byte for byte as
                                    data (the reference compiles them at run time from
                                    $HOME/mpv-build/mpv/video/filter/HopperRender/Kernels,
                                    opticalFlowCalc.c:57,373); oracle/ref_opencl.py writes them to a
                                    scratch $HOME on the box before calling initOpticalFlowCalc.

The OpenCL *device* is the B200 through the driver's libnvidia-opencl.so.1 (present on the GPU
box, absent here), so this library can only RUN under gpurun; it builds here without a GPU.
The reference's own build system (meson, ffmpeg, libplacebo ...) is not used.
"""
import os
import pathlib
import subprocess
import sys
import tempfile

HERE = pathlib.Path(__file__).resolve().parent
REF = pathlib.Path("/root/reference/video/filter/HopperRender")
OUT = HERE / "_ref"
KERNELS = ["calcDeltaSumsKernel", "determineLowestLayerKernel", "adjustOffsetArrayKernel", "blurFlowKernel", "warpFrameKernel"]
CC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def find_loader():
    for d in ("/usr/local/cuda/targets/x86_64-linux/lib", "/usr/local/cuda-12.9/targets/x86_64-linux/lib", "/usr/local/cuda/lib64"):
        p = pathlib.Path(d) / "libOpenCL.so.1"
        if p.exists():
            return p
    return None


def main():
    if not REF.is_dir():
        print("build_ref: /root/reference not present (GPU box) — using the prebuilt oracle/_ref if any")
        return 0
    OUT.mkdir(exist_ok=True)
    loader = find_loader()
    if loader is None:
        print("build_ref: no libOpenCL.so.1 (ICD loader) in the CUDA toolkit — reference host not built")
        return 0
    # 1. the reference host, unmodified; stdlib/string are force-included because the file relies on
    #    <CL/cl.h> pulling them in
    cmd = [CC, "-O2", "-std=gnu11", "-w", "-fPIC", "-shared", "-I", str(HERE / "cl_shim"), "-I", str(REF),
           "-include", "stdlib.h", "-include", "string.h",
           "-o", str(OUT / "libhr_ref_ofc.so"), str(REF / "opticalFlowCalc.c"),
           "-L", str(loader.parent), "-l:libOpenCL.so.1", "-Wl,-rpath," + str(loader.parent), "-lm"]
    subprocess.run(cmd, check=True)
    # 2. the kernel sources as data in a shared object; the generated C lives in a temp dir only
    with tempfile.TemporaryDirectory(dir=str(OUT)) as td:
        c = pathlib.Path(td) / "k.c"
        with open(c, "w") as f:
            f.write("#include <string.h>\n")
            for k in KERNELS:
                data = (REF / "Kernels" / (k + ".cl")).read_bytes()
                f.write("static const unsigned char src_%s[] = {%s,0};\n" % (k, ",".join(str(b) for b in data)))
            f.write("const char *hr_ref_kernel_source(const char *name) {\n")
            for k in KERNELS:
                f.write('  if (!strcmp(name, "%s")) return (const char *)src_%s;\n' % (k, k))
            f.write("  return 0;\n}\n")
            f.write("int hr_ref_kernel_count(void) { return %d; }\n" % len(KERNELS))
            f.write("const char *hr_ref_kernel_name(int i) {\n  static const char *n[] = {%s};\n  return (i >= 0 && i < %d) ? n[i] : 0;\n}\n"
                    % (",".join('"%s"' % k for k in KERNELS), len(KERNELS)))
        subprocess.run([CC, "-O1", "-fPIC", "-shared", "-o", str(OUT / "libhr_ref_kernels.so"), str(c)], check=True)
    print("build_ref: built", OUT / "libhr_ref_ofc.so", "and", OUT / "libhr_ref_kernels.so")
    # 3. the reference FILTER (vf_HopperRender.c, unmodified) on top of this repository's optical-flow-calc
    #    layer and CUDA library: oracle/filter_host_sim.c supplies the slice of mpv's filter runtime it needs
    #    (oracle/mpv_shim). The radius auto-adjust is compiled out so that runs are reproducible.
    root = HERE.parent
    pkg_hr = root / "mpv-frame-interpolator_b200" / "mpv" / "video" / "filter" / "HopperRender"
    csrc = root / "mpv-frame-interpolator_b200" / "csrc"
    if (csrc / "libhopperrender_cuda.so").exists():
        # -include: this repository's opticalFlowCalc.h comes first and owns the include guard, so the
        # reference's OpenCL flavour of that header (same directory as the filter source) stays empty
        cmd = [CC, "-O2", "-std=gnu11", "-w", "-fPIC", "-shared", "-include", str(pkg_hr / "opticalFlowCalc.h"),
               "-I", str(pkg_hr), "-I", str(HERE / "mpv_shim"), "-I", str(root / "include"),
               "-o", str(OUT / "libhr_filter_sim.so"),
               str(REF / "vf_HopperRender.c"), str(HERE / "filter_host_sim.c"), str(pkg_hr / "opticalFlowCalc.c"),
               "-L", str(csrc), "-lhopperrender_cuda", "-Wl,-rpath,$ORIGIN/../../mpv-frame-interpolator_b200/csrc", "-lm", "-lpthread"]
        subprocess.run(cmd, check=True)
        print("build_ref: built", OUT / "libhr_filter_sim.so", "(reference filter + this repo's OFC layer)")
        # 4. the same for P010: the reference filter source is still compiled as it is; the two
        #    changes a maintainer makes in it for P010 — `ofc->pixelFormat = 1` before initOpticalFlowCalc and the
        #    stride in samples instead of bytes (vf_HopperRender.c:446; format negotiation :385, :668 is the harness's
        #    business here) — are expressed as a function-like macro seen by that translation unit only.
        with tempfile.TemporaryDirectory(dir=str(OUT)) as td:
            o1, o2, o3 = (str(pathlib.Path(td) / n) for n in ("vf.o", "sim.o", "ofc.o"))
            common = [CC, "-O2", "-std=gnu11", "-w", "-fPIC", "-c", "-include", str(pkg_hr / "opticalFlowCalc.h"),
                      "-I", str(pkg_hr), "-I", str(HERE / "mpv_shim"), "-I", str(root / "include")]
            subprocess.run(common + ["-include", str(HERE / "mpv_shim" / "hr_p010_patch.h"), "-o", o1, str(REF / "vf_HopperRender.c")], check=True)
            subprocess.run(common + ["-DHR_SIM_BPS=2", "-o", o2, str(HERE / "filter_host_sim.c")], check=True)
            subprocess.run(common + ["-o", o3, str(pkg_hr / "opticalFlowCalc.c")], check=True)
            subprocess.run([CC, "-shared", "-o", str(OUT / "libhr_filter_sim_p010.so"), o1, o2, o3,
                            "-L", str(csrc), "-lhopperrender_cuda", "-Wl,-rpath,$ORIGIN/../../mpv-frame-interpolator_b200/csrc", "-lm", "-lpthread"], check=True)
        print("build_ref: built", OUT / "libhr_filter_sim_p010.so", "(the same, P010)")
        # 5. the filter with this repository's patch series applied (patches/*.patch: P010 negotiation, the headless
        #    control channel, IMGFMT_CUDA in and out). The patches are applied to a scratch copy of the reference's
        #    source that lives only while this step runs; the harness gets the CUDA runtime for its device images.
        patches = sorted((root / "patches").glob("*.patch"))
        cudart = next((d for d in ("/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib") if (pathlib.Path(d) / "libcudart.so").exists()), None)
        if patches and cudart:
            with tempfile.TemporaryDirectory(dir=str(OUT)) as td:
                work = pathlib.Path(td) / "video" / "filter" / "HopperRender"
                work.mkdir(parents=True)
                (work / "vf_HopperRender.c").write_bytes((REF / "vf_HopperRender.c").read_bytes())
                for pt in patches:
                    subprocess.run(["patch", "-p1", "-s", "-d", td, "-i", str(pt)], check=True)
                cmd = [CC, "-O2", "-std=gnu11", "-w", "-fPIC", "-shared", "-DHR_SIM_CUDA", "-include", str(pkg_hr / "opticalFlowCalc.h"),
                       "-I", str(pkg_hr), "-I", str(HERE / "mpv_shim"), "-I", str(root / "include"),
                       "-o", str(OUT / "libhr_filter_sim_patched.so"),
                       str(work / "vf_HopperRender.c"), str(HERE / "filter_host_sim.c"), str(pkg_hr / "opticalFlowCalc.c"), str(pkg_hr / "hrControl.c"),
                       "-L", str(csrc), "-lhopperrender_cuda", "-Wl,-rpath,$ORIGIN/../../mpv-frame-interpolator_b200/csrc",
                       "-L", cudart, "-lcudart", "-Wl,-rpath," + cudart, "-lm", "-lpthread"]
                subprocess.run(cmd, check=True)
            print("build_ref: built", OUT / "libhr_filter_sim_patched.so", "(reference filter + patches/%s)" % ", ".join(p.name for p in patches))
    else:
        print("build_ref: libhopperrender_cuda.so not built yet — filter host sim skipped")
    return 0


if __name__ == "__main__":
    sys.exit(main())
