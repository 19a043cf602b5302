"""ctypes wrapper of the CPU ORACLE (oracle/hr_oracle.c). TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs — never from the product package.
"""
import ctypes as C
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
LIB = HERE / "libhr_oracle.so"
# MUFU.RCP(i), i = 0..1023, read back from the B200 (tests/golden/make_rcp_table.py)
RCP_TABLE = HERE.parent / "tests" / "golden" / "mufu_rcp_table.npy"
ARITH_IEEE, ARITH_NVCL = 0, 1
_lib = None
_rcp = None


def build(force=False):
    srcs = [HERE / "hr_oracle.c", HERE / "hr_oracle.h"]
    if not force and LIB.exists() and all(s.stat().st_mtime <= LIB.stat().st_mtime for s in srcs):
        return LIB
    subprocess.run(["make", "-s", "-C", str(HERE), "libhr_oracle.so"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.hro_create.restype = C.c_void_p
        L.hro_create.argtypes = [C.c_int] * 4
        L.hro_destroy.argtypes = [C.c_void_p]
        for n in ("hro_low_width", "hro_low_height", "hro_res_scalar", "hro_num_steps"):
            getattr(L, n).argtypes = [C.c_void_p]
            getattr(L, n).restype = C.c_int
        L.hro_update_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hro_calc_flow.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.hro_warp.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_float, C.c_float]
        L.hro_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hro_get_offsets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hro_set_blurred_offsets.argtypes = [C.c_void_p, C.c_void_p]
        L.hro_get_step_layers.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hro_get_last_sums.argtypes = [C.c_void_p, C.c_void_p]
        L.hro_blur_flow.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.hro_blur_flow.restype = None
        L.hro_num_threads.restype = C.c_int
        L.hro_set_rcp_table.argtypes = [C.c_void_p, C.c_int]
        L.hro_set_rcp_table.restype = None
        L.hro_set_arith.argtypes = [C.c_void_p, C.c_int]
        _lib = L
        if RCP_TABLE.exists():
            global _rcp
            _rcp = np.ascontiguousarray(np.load(RCP_TABLE), np.float32)
            L.hro_set_rcp_table(C.c_void_p(_rcp.ctypes.data), int(_rcp.size))
    return _lib


def have_nvcl():
    lib()
    return _rcp is not None


def _p(a):
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


class Oracle:
    def __init__(self, frameHeight, frameWidth, actualWidth=None, pixfmt=0, arith=None):
        """arith: ARITH_IEEE, ARITH_NVCL, or None = NVCL when the MUFU.RCP table is available."""
        self.L = lib()
        aw = frameWidth if actualWidth is None else actualWidth
        self.o = self.L.hro_create(frameHeight, frameWidth, aw, pixfmt)
        if not self.o:
            raise RuntimeError("hro_create failed")
        self.H, self.W, self.aW, self.pixfmt = frameHeight, frameWidth, aw, pixfmt
        self.dtype = np.uint16 if pixfmt == 1 else np.uint8
        self.lw = self.L.hro_low_width(self.o)
        self.lh = self.L.hro_low_height(self.o)
        self.s = self.L.hro_res_scalar(self.o)
        self.steps = self.L.hro_num_steps(self.o)
        self.radius = 0
        self.arith = (ARITH_NVCL if _rcp is not None else ARITH_IEEE) if arith is None else arith
        if self.L.hro_set_arith(self.o, self.arith):
            raise RuntimeError("NVCL arithmetic needs tests/golden/mufu_rcp_table.npy")

    def close(self):
        if self.o:
            self.L.hro_destroy(self.o)
            self.o = None

    def __del__(self):
        self.close()

    def update_frame(self, y, uv):
        y = np.ascontiguousarray(y, self.dtype)
        uv = np.ascontiguousarray(uv, self.dtype)
        assert y.shape == (self.H, self.W) and uv.shape == (self.H // 2, self.W)
        assert self.L.hro_update_frame(self.o, _p(y), _p(uv)) == 0

    def calc_flow(self, radius=5, deltaScalar=8, neighborBiasScalar=6):
        self.radius = radius
        assert self.L.hro_calc_flow(self.o, radius, deltaScalar, neighborBiasScalar) == 0

    def warp(self, t, mode=2, black=0.0, white=255.0):
        return self.L.hro_warp(self.o, float(t), int(mode), float(black), float(white))

    def download(self):
        y = np.empty((self.H, self.W), self.dtype)
        uv = np.empty((self.H // 2, self.W), self.dtype)
        self.L.hro_download(self.o, _p(y), _p(uv))
        return y, uv

    def get_offsets(self):
        raw = np.empty((2, self.lh, self.lw), np.int16)
        blurred = np.empty((2, self.lh, self.lw), np.int16)
        self.L.hro_get_offsets(self.o, _p(raw), _p(blurred))
        return raw, blurred

    def set_blurred_offsets(self, b):
        b = np.ascontiguousarray(b, np.int16)
        assert b.shape == (2, self.lh, self.lw)
        self.L.hro_set_blurred_offsets(self.o, _p(b))

    def get_step_layers(self, step):
        out = np.empty((self.lh, self.lw), np.uint8)
        assert self.L.hro_get_step_layers(self.o, step, _p(out)) == 0
        return out

    def get_last_sums(self):
        out = np.empty((self.radius, self.lh, self.lw), np.uint32)
        assert self.L.hro_get_last_sums(self.o, _p(out)) == 0
        return out


def blur_flow(raw):
    raw = np.ascontiguousarray(raw, np.int16)
    out = np.empty_like(raw)
    lib().hro_blur_flow(_p(raw), _p(out), raw.shape[1], raw.shape[2])
    return out


def num_threads():
    return lib().hro_num_threads()
