"""Run the REFERENCE ITSELF: its unmodified opticalFlowCalc.c (oracle/_ref/libhr_ref_ofc.so) driving
its unmodified .cl kernels through the NVIDIA OpenCL ICD on the B200. TEST INFRASTRUCTURE ONLY.

Only usable where an OpenCL device exists (the GPU box: /usr/lib/libnvidia-opencl.so.1); tests
that use it skip elsewhere. See oracle/build_ref.py for how oracle/_ref/ is produced.
"""
import ctypes as C
import glob
import os
import pathlib
import tempfile

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
REFDIR = HERE / "_ref"

_state = {}


def _find_icd():
    for pat in ("/usr/lib/libnvidia-opencl.so.1", "/usr/local/nvidia/lib/libnvidia-opencl.so.1",
                "/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so.1", "/usr/lib64/libnvidia-opencl.so.1"):
        if os.path.exists(pat):
            return pat
    g = glob.glob("/usr/**/libnvidia-opencl.so.1", recursive=True)
    return g[0] if g else None


class RefOFC(C.Structure):
    """struct OpticalFlowCalc exactly as declared by the reference (opticalFlowCalc.h:10-65)."""
    _fields_ = [
        ("isInitialized", C.c_bool), ("frameWidth", C.c_int), ("frameHeight", C.c_int), ("actualWidth", C.c_int),
        ("outputBlackLevel", C.c_float), ("outputWhiteLevel", C.c_float),
        ("opticalFlowResScalar", C.c_int), ("opticalFlowFrameWidth", C.c_int), ("opticalFlowFrameHeight", C.c_int),
        ("opticalFlowSearchRadius", C.c_int), ("ofcCalcTime", C.c_double), ("warpCalcTime", C.c_double),
        ("deltaScalar", C.c_int), ("neighborBiasScalar", C.c_int),
        ("clDeviceId", C.c_void_p), ("clContext", C.c_void_p),
        ("lowGrid16x16x2", C.c_size_t * 3), ("lowGrid16x16x1", C.c_size_t * 3), ("lowGrid8x8xL", C.c_size_t * 3),
        ("halfGrid16x16x1", C.c_size_t * 3), ("grid16x16x1", C.c_size_t * 3),
        ("threads16x16x1", C.c_size_t * 3), ("threads8x8x1", C.c_size_t * 3),
        ("queue", C.c_void_p), ("ofcStartedEvent", C.c_void_p), ("warpStartedEvent", C.c_void_p),
        ("offsetArray", C.c_void_p), ("blurredOffsetArray", C.c_void_p), ("summedDeltaValuesArray", C.c_void_p),
        ("lowestLayerArray", C.c_void_p), ("outputFrameArray", C.c_void_p), ("inputFrameArray", C.c_void_p * 2),
        ("calcDeltaSumsKernel", C.c_void_p), ("determineLowestLayerKernel", C.c_void_p),
        ("adjustOffsetArrayKernel", C.c_void_p), ("blurFlowKernel", C.c_void_p), ("warpFrameKernel", C.c_void_p),
        ("_slack", C.c_char * 256),
    ]


def available():
    """(ok, reason). Loads the libraries and checks that an OpenCL platform answers."""
    if "ok" in _state:
        return _state["ok"], _state["why"]

    def done(ok, why):
        _state["ok"], _state["why"] = ok, why
        return ok, why

    if not (REFDIR / "libhr_ref_ofc.so").exists() or not (REFDIR / "libhr_ref_kernels.so").exists():
        return done(False, "oracle/_ref not built (python oracle/build_ref.py where /root/reference exists)")
    icd = _find_icd()
    if not icd:
        return done(False, "no libnvidia-opencl.so.1 (no OpenCL device)")
    os.environ.setdefault("OCL_ICD_FILENAMES", icd)
    os.environ.setdefault("OCL_ICD_VENDORS", "/nonexistent")
    try:
        ocl = None
        for cand in ("/usr/local/cuda/targets/x86_64-linux/lib/libOpenCL.so.1", "/usr/local/cuda-12.9/targets/x86_64-linux/lib/libOpenCL.so.1", "libOpenCL.so.1"):
            try:
                ocl = C.CDLL(cand, mode=C.RTLD_GLOBAL)
                break
            except OSError:
                continue
        if ocl is None:
            return done(False, "libOpenCL.so.1 (ICD loader) not loadable")
        n = C.c_uint(0)
        rc = ocl.clGetPlatformIDs(0, None, C.byref(n))
        if rc != 0 or n.value == 0:
            return done(False, "clGetPlatformIDs rc=%d platforms=%d" % (rc, n.value))
        ofc = C.CDLL(str(REFDIR / "libhr_ref_ofc.so"))
        ker = C.CDLL(str(REFDIR / "libhr_ref_kernels.so"))
    except OSError as e:
        return done(False, "load failed: %s" % e)
    ker.hr_ref_kernel_source.restype = C.c_char_p
    ker.hr_ref_kernel_source.argtypes = [C.c_char_p]
    ker.hr_ref_kernel_name.restype = C.c_char_p
    ker.hr_ref_kernel_name.argtypes = [C.c_int]
    # the reference reads its kernels from $HOME/mpv-build/mpv/video/filter/HopperRender/Kernels
    home = tempfile.mkdtemp(prefix="hr_ref_home_")
    kd = pathlib.Path(home) / "mpv-build" / "mpv" / "video" / "filter" / "HopperRender" / "Kernels"
    kd.mkdir(parents=True)
    for i in range(ker.hr_ref_kernel_count()):
        name = ker.hr_ref_kernel_name(i).decode()
        (kd / (name + ".cl")).write_bytes(ker.hr_ref_kernel_source(name.encode()))
    _state["home"] = home
    for fn in ("initOpticalFlowCalc", "updateFrame", "downloadFrame", "calculateOpticalFlow", "warpFrames"):
        getattr(ofc, fn).restype = C.c_bool
    ofc.initOpticalFlowCalc.argtypes = [C.POINTER(RefOFC), C.c_int, C.c_int, C.c_int]
    ofc.updateFrame.argtypes = [C.POINTER(RefOFC), C.POINTER(C.c_void_p)]
    ofc.downloadFrame.argtypes = [C.POINTER(RefOFC), C.POINTER(C.c_void_p)]
    ofc.calculateOpticalFlow.argtypes = [C.POINTER(RefOFC)]
    ofc.warpFrames.argtypes = [C.POINTER(RefOFC), C.c_float, C.c_int]
    ofc.freeOFC.argtypes = [C.POINTER(RefOFC)]
    ofc.freeOFC.restype = None
    ocl.clEnqueueReadBuffer.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t, C.c_size_t, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
    ocl.clEnqueueWriteBuffer.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t, C.c_size_t, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
    ocl.clFinish.argtypes = [C.c_void_p]
    _state.update(ocl=ocl, ofc=ofc)
    return done(True, "NVIDIA OpenCL ICD %s" % icd)


class Reference:
    """The reference's six functions on the reference's own struct (NV12 only, like the reference)."""

    def __init__(self, frameHeight, frameWidth, actualWidth=None):
        ok, why = available()
        if not ok:
            raise RuntimeError(why)
        self.lib, self.ocl = _state["ofc"], _state["ocl"]
        self.s = RefOFC()
        old = os.environ.get("HOME")
        os.environ["HOME"] = _state["home"]
        try:
            failed = self.lib.initOpticalFlowCalc(C.byref(self.s), frameHeight, frameWidth, frameWidth if actualWidth is None else actualWidth)
        finally:
            if old is not None:
                os.environ["HOME"] = old
        if failed:
            raise RuntimeError("reference initOpticalFlowCalc failed")
        self.H, self.W = frameHeight, frameWidth
        self.lw, self.lh = self.s.opticalFlowFrameWidth, self.s.opticalFlowFrameHeight

    def close(self):
        if self.s.isInitialized:
            self.lib.freeOFC(C.byref(self.s))
            self.s.isInitialized = False

    def update_frame(self, y, uv):
        y = np.ascontiguousarray(y, np.uint8)
        uv = np.ascontiguousarray(uv, np.uint8)
        planes = (C.c_void_p * 2)(y.ctypes.data, uv.ctypes.data)
        assert not self.lib.updateFrame(C.byref(self.s), planes)

    def calc_flow(self, radius=5, deltaScalar=8, neighborBiasScalar=6):
        self.s.opticalFlowSearchRadius = radius
        self.s.deltaScalar = deltaScalar
        self.s.neighborBiasScalar = neighborBiasScalar
        assert not self.lib.calculateOpticalFlow(C.byref(self.s))
        return self.s.ofcCalcTime

    def warp(self, t, mode=2, black=0.0, white=255.0):
        self.s.outputBlackLevel = black
        self.s.outputWhiteLevel = white
        return bool(self.lib.warpFrames(C.byref(self.s), float(t), int(mode)))

    def download(self):
        y = np.zeros((self.H, self.W), np.uint8)
        uv = np.zeros((self.H // 2, self.W), np.uint8)
        planes = (C.c_void_p * 2)(y.ctypes.data, uv.ctypes.data)
        assert not self.lib.downloadFrame(C.byref(self.s), planes)
        return y, uv

    def _read(self, mem, arr):
        assert self.ocl.clEnqueueReadBuffer(self.s.queue, mem, 1, 0, arr.nbytes, arr.ctypes.data, 0, None, None) == 0
        return arr

    def get_offsets(self):
        raw = np.empty((2, self.lh, self.lw), np.int16)
        blurred = np.empty((2, self.lh, self.lw), np.int16)
        self._read(self.s.offsetArray, raw)
        self._read(self.s.blurredOffsetArray, blurred)
        return raw, blurred

    def set_blurred_offsets(self, b):
        b = np.ascontiguousarray(b, np.int16)
        assert self.ocl.clEnqueueWriteBuffer(self.s.queue, self.s.blurredOffsetArray, 1, 0, b.nbytes, b.ctypes.data, 0, None, None) == 0

    def get_last_layers(self):
        return self._read(self.s.lowestLayerArray, np.empty((self.lh, self.lw), np.uint8))

    def get_last_sums(self, radius):
        return self._read(self.s.summedDeltaValuesArray, np.empty((radius, self.lh, self.lw), np.uint32))
