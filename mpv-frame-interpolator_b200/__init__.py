"""hopperrender_b200 — B200-native HopperRender hot path (package dir `mpv-frame-interpolator_b200/`).

Python here is plumbing only: ctypes bindings of the C ABI (`include/hopperrender_cuda.h`)
and a mirror of the reference's optical-flow-calc interface
(`video/filter/HopperRender/opticalFlowCalc.h:77-124`) with the same names, argument meaning
and inverted-bool error convention, so tests read like calls from `vf_HopperRender.c`.

There is NO CPU fallback: if the CUDA library is missing or no GPU is present every compute
entry point raises / reports failure.
"""
import ctypes as C
import os
import pathlib

import numpy as np

from . import build as _build

PKG = pathlib.Path(__file__).resolve().parent
ROOT = PKG.parent

PIXFMT_NV12 = 0
PIXFMT_P010 = 1

# enum FrameOutput, video/filter/HopperRender/vf_HopperRender.c:21
WarpedFrame12, WarpedFrame21, BlendedFrame, HSVFlow, GreyFlow, SideBySide1, SideBySide2 = range(7)

# video/filter/HopperRender/config.h
MAX_CALC_RES = 270
MIN_SEARCH_RADIUS = 5
MAX_SEARCH_RADIUS = 16


class HrInfo(C.Structure):
    _fields_ = [
        ("abiVersion", C.c_int), ("device", C.c_int), ("frameHeight", C.c_int), ("frameWidth", C.c_int),
        ("actualWidth", C.c_int), ("pixfmt", C.c_int), ("resScalar", C.c_int), ("lowWidth", C.c_int),
        ("lowHeight", C.c_int), ("firstWindow", C.c_int), ("iterations", C.c_int), ("searchCtas", C.c_int),
        ("smCount", C.c_int), ("frameBytes", C.c_size_t), ("deviceBytes", C.c_size_t),
    ]


# name -> (restype, argtypes); also the list tests check against include/hopperrender_cuda.h
ABI = {
    "hr_abi_version": (C.c_int, []),
    "hr_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "hr_destroy": (C.c_int, [C.c_void_p]),
    "hr_get_info": (C.c_int, [C.c_void_p, C.POINTER(HrInfo)]),
    "hr_last_error": (C.c_char_p, [C.c_void_p]),
    "hr_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hr_synchronize": (C.c_int, [C.c_void_p]),
    "hr_update_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hr_update_frame_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "hr_set_pipeline": (C.c_int, [C.c_void_p, C.c_int]),
    "hr_pipeline_join": (C.c_int, [C.c_void_p]),
    "hr_step_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int,
                                 C.c_float, C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "hr_steps_device": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                  C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "hr_calc_flow": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "hr_warp": (C.c_int, [C.c_void_p, C.c_float, C.c_int, C.c_float, C.c_float]),
    "hr_warp_batch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "hr_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "hr_finish": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "hr_debug_host_transfer_bytes": (C.c_int, [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "hr_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "hr_host_free": (C.c_int, [C.c_void_p]),
    "hr_debug_host_pointer_kind": (C.c_int, [C.c_void_p]),
    "hr_get_output_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "hr_set_output_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hr_band_configure": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hr_band_set_max_radius": (C.c_int, [C.c_void_p, C.c_int]),
    "hr_band_get_halo": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)]),
    "hr_band_local_pointers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "hr_band_export_ipc": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hr_band_open_ipc": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "hr_band_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hr_band_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "hr_band_gather": (C.c_int, [C.c_void_p, C.c_int]),
    "hr_band_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "hr_get_offsets": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hr_set_blurred_offsets": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hr_blur_flow": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hr_set_trace": (C.c_int, [C.c_void_p, C.c_int]),
    "hr_get_step_layers": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "hr_debug_rcp_table": (C.c_int, [C.c_void_p, C.c_int]),
    "hr_debug_int_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "hr_debug_predict_pacing": (C.c_int, [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hr_debug_set_search_generation": (C.c_int, [C.c_void_p, C.c_int]),
    "hr_debug_last_search_generation": (C.c_int, [C.c_void_p]),
    "hr_debug_set_search_staged": (C.c_int, [C.c_void_p, C.c_int]),
    "hr_debug_last_search_staged": (C.c_int, [C.c_void_p]),
    "hr_set_timeline": (C.c_int, [C.c_void_p, C.c_int]),
    "hr_get_timeline": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hr_debug_peek_timeline": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hr_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "hr_get_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "hr_get_launch_count": (C.c_uint64, [C.c_void_p]),
}

_lib = None


def cuda_library_path() -> pathlib.Path:
    return _build.CUDA_LIB


def load_library(build_if_missing=True):
    """dlopen libhopperrender_cuda.so (in-tree). Raises if it cannot be built/loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = cuda_library_path()
    if os.environ.get("HR_CUDA_LIB"):          # kernel-variant experiments (tools/); never a fallback
        path = pathlib.Path(os.environ["HR_CUDA_LIB"])
    if not path.exists():
        if not build_if_missing:
            raise RuntimeError("libhopperrender_cuda.so is not built (run __graft_entry__.build())")
        _build.build_cuda()
    lib = C.CDLL(str(path))
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)  # AttributeError here = ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class HrError(RuntimeError):
    pass


def _ptr(a):
    """Host pointer of a numpy array / int address / torch tensor (data_ptr)."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("plane must be C-contiguous")
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError("unsupported buffer type %r" % type(a))


def debug_rcp_table(n=1024):
    """MUFU.RCP(i) for i in range(n), read back from the GPU (parity tap)."""
    out = np.zeros(n, np.float32)
    if load_library().hr_debug_rcp_table(_ptr(out), n):
        raise HrError(load_library().hr_last_error(None).decode())
    return out


def host_transfer_bytes():
    """(host->device, device->host) bytes moved by the interface calls in this process so far."""
    a, b = C.c_ulonglong(0), C.c_ulonglong(0)
    load_library().hr_debug_host_transfer_bytes(C.byref(a), C.byref(b))
    return a.value, b.value


def debug_int_peak(device=-1):
    """(packed-byte SADs per second over the whole GPU, VABSDIFF4 warp instructions per SM clock): the INT roofline
    denominator of the search, measured live (developer tap)."""
    a, b = C.c_double(0.0), C.c_double(0.0)
    if load_library().hr_debug_int_peak(int(device), C.byref(a), C.byref(b)):
        raise HrError(load_library().hr_last_error(None).decode())
    return a.value, b.value


def debug_predict_pacing(scalars):
    """The library's guess for every blending scalar of a call sequence, made from the scalars before it (developer tap,
    host only): list of (predicted float32, expects_new_source_frame, has_guess)."""
    n = len(scalars)
    ts = (C.c_float * max(1, n))(*[float(t) for t in scalars])
    pred, nf, have = (C.c_float * max(1, n))(), (C.c_int * max(1, n))(), (C.c_int * max(1, n))()
    if load_library().hr_debug_predict_pacing(ts, n, pred, nf, have):
        raise HrError("hr_debug_predict_pacing failed")
    return [(np.float32(pred[i]), bool(nf[i]), bool(have[i])) for i in range(n)]


class HrCuda:
    """Thin object wrapper over the C ABI (one HrContext)."""

    def __init__(self, frameHeight, frameWidth, actualWidth=None, pixfmt=PIXFMT_NV12, device=-1):
        self.lib = load_library()
        self.h = C.c_void_p()
        aw = frameWidth if actualWidth is None else actualWidth
        if self.lib.hr_create(C.byref(self.h), frameHeight, frameWidth, aw, pixfmt, device):
            raise HrError(self.lib.hr_last_error(None).decode())
        self.info = HrInfo()
        self.lib.hr_get_info(self.h, C.byref(self.info))
        self.pixfmt = pixfmt
        self.dtype = np.uint16 if pixfmt == PIXFMT_P010 else np.uint8

    def close(self):
        if self.h:
            self.lib.hr_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise HrError(self.lib.hr_last_error(self.h).decode())

    def last_error(self):
        return self.lib.hr_last_error(self.h).decode()

    def set_stream(self, stream_handle):
        self._chk(self.lib.hr_set_stream(self.h, C.c_void_p(stream_handle) if stream_handle else None))

    def synchronize(self):
        self._chk(self.lib.hr_synchronize(self.h))

    def update_frame(self, y, uv):
        self._chk(self.lib.hr_update_frame(self.h, _ptr(y), _ptr(uv)))

    def update_frame_device(self, dy, duv, borrow=False):
        self._chk(self.lib.hr_update_frame_device(self.h, _ptr(dy), _ptr(duv), 1 if borrow else 0))

    def set_pipeline(self, on=True):
        self._chk(self.lib.hr_set_pipeline(self.h, 1 if on else 0))

    def pipeline_join(self):
        self._chk(self.lib.hr_pipeline_join(self.h))

    def step_device(self, dy, duv, ts, outs, radius=MIN_SEARCH_RADIUS, deltaScalar=8, neighborBiasScalar=6, mode=BlendedFrame,
                    black=0.0, white=255.0, borrow=True):
        """One source frame of a device-resident stream: update + flow + len(ts) warps into outs[i] = (dY, dUV)."""
        n = len(ts)
        tarr = (C.c_float * max(n, 1))(*[float(t) for t in ts])
        oy = (C.c_void_p * max(n, 1))(*[_ptr(o[0]) for o in outs[:n]])
        ouv = (C.c_void_p * max(n, 1))(*[_ptr(o[1]) for o in outs[:n]])
        self._chk(self.lib.hr_step_device(self.h, _ptr(dy), _ptr(duv), 1 if borrow else 0, radius, deltaScalar, neighborBiasScalar, n, tarr,
                                          int(mode), float(black), float(white), oy, ouv))

    def steps_device(self, frames, ts_per_step, outs, radius=MIN_SEARCH_RADIUS, deltaScalar=8, neighborBiasScalar=6, mode=BlendedFrame,
                     black=0.0, white=255.0, borrow=True):
        """len(frames) consecutive source frames (dY, dUV) in one call; outs: one (dY, dUV) per warp, in order."""
        n = len(frames)
        fy = (C.c_void_p * max(n, 1))(*[_ptr(f[0]).value for f in frames])
        fuv = (C.c_void_p * max(n, 1))(*[_ptr(f[1]).value for f in frames])
        nw = (C.c_int * max(n, 1))(*[len(t) for t in ts_per_step])
        flat = [float(t) for step in ts_per_step for t in step]
        m = len(flat)
        tarr = (C.c_float * max(m, 1))(*flat)
        oy = (C.c_void_p * max(m, 1))(*[_ptr(o[0]).value for o in outs[:m]])
        ouv = (C.c_void_p * max(m, 1))(*[_ptr(o[1]).value for o in outs[:m]])
        self._chk(self.lib.hr_steps_device(self.h, n, fy, fuv, 1 if borrow else 0, radius, deltaScalar, neighborBiasScalar, nw, tarr, int(mode),
                                           float(black), float(white), oy, ouv))

    def calc_flow(self, radius=MIN_SEARCH_RADIUS, deltaScalar=8, neighborBiasScalar=6, blocking=True):
        sec = C.c_double(0.0)
        self._chk(self.lib.hr_calc_flow(self.h, radius, deltaScalar, neighborBiasScalar, C.byref(sec) if blocking else None))
        return sec.value

    def warp(self, t, mode=BlendedFrame, black=0.0, white=255.0):
        self._chk(self.lib.hr_warp(self.h, float(t), int(mode), float(black), float(white)))

    def warp_batch(self, ts, outs, mode=BlendedFrame, black=0.0, white=255.0):
        """len(ts) output frames of the current pair into outs[i] = (dY, dUV) (device planes), one launch per 8."""
        n = len(ts)
        tarr = (C.c_float * max(n, 1))(*[float(t) for t in ts])
        oy = (C.c_void_p * max(n, 1))(*[_ptr(o[0]).value for o in outs[:n]])
        ouv = (C.c_void_p * max(n, 1))(*[_ptr(o[1]).value for o in outs[:n]])
        self._chk(self.lib.hr_warp_batch(self.h, n, tarr, int(mode), float(black), float(white), oy, ouv))

    def download(self, y=None, uv=None):
        H, W = self.info.frameHeight, self.info.frameWidth
        if y is None:
            y = np.empty((H, W), self.dtype)
        if uv is None:
            uv = np.empty((H // 2, W), self.dtype)
        sec = C.c_double(0.0)
        self._chk(self.lib.hr_download(self.h, _ptr(y), _ptr(uv), C.byref(sec)))
        return y, uv, sec.value

    def finish(self):
        sec = C.c_double(0.0)
        self._chk(self.lib.hr_finish(self.h, C.byref(sec)))
        return sec.value

    def output_device_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._chk(self.lib.hr_get_output_device(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_output_device(self, dy, duv):
        self._chk(self.lib.hr_set_output_device(self.h, _ptr(dy), _ptr(duv)))

    def get_offsets(self):
        n = (2, self.info.lowHeight, self.info.lowWidth)
        raw, blurred = np.empty(n, np.int16), np.empty(n, np.int16)
        self._chk(self.lib.hr_get_offsets(self.h, _ptr(raw), _ptr(blurred)))
        return raw, blurred

    def set_blurred_offsets(self, blurred):
        b = np.ascontiguousarray(blurred, np.int16)
        assert b.shape == (2, self.info.lowHeight, self.info.lowWidth)
        self._chk(self.lib.hr_set_blurred_offsets(self.h, _ptr(b)))

    def blur_flow(self, raw):
        r = np.ascontiguousarray(raw, np.int16)
        assert r.shape == (2, self.info.lowHeight, self.info.lowWidth)
        out = np.empty_like(r)
        self._chk(self.lib.hr_blur_flow(self.h, _ptr(r), _ptr(out)))
        return out

    def set_trace(self, on=True):
        self._chk(self.lib.hr_set_trace(self.h, 1 if on else 0))

    def get_step_layers(self, step):
        out = np.empty((self.info.lowHeight, self.info.lowWidth), np.uint8)
        self._chk(self.lib.hr_get_step_layers(self.h, step, _ptr(out)))
        return out

    def set_search_generation(self, generation):
        """Developer knob: 0 = chosen per launch (default), 3 / 2 = csrc/hr_search3.cuh / hr_search2.cuh where they apply, 1 = csrc/hr_search.cuh always."""
        self._chk(self.lib.hr_debug_set_search_generation(self.h, int(generation)))

    def last_search_generation(self):
        return int(self.lib.hr_debug_last_search_generation(self.h))

    def set_search_staged(self, on=True):
        """Developer knob: the TMA-staged variant of generation 2 where it applies (default) or never."""
        self._chk(self.lib.hr_debug_set_search_staged(self.h, 1 if on else 0))

    def last_search_staged(self):
        return bool(self.lib.hr_debug_last_search_staged(self.h))

    def set_timeline(self, on=True):
        self._chk(self.lib.hr_set_timeline(self.h, 1 if on else 0))

    def get_timeline(self):
        out = np.zeros((self.info.searchCtas, 128), np.int64)
        self._chk(self.lib.hr_get_timeline(self.h, _ptr(out), self.info.searchCtas))
        return out

    def peek_timeline(self):
        out = np.zeros((self.info.searchCtas, 128), np.int64)
        self._chk(self.lib.hr_debug_peek_timeline(self.h, _ptr(out), self.info.searchCtas))
        return out

    def set_profiling(self, on=True):
        self._chk(self.lib.hr_set_profiling(self.h, 1 if on else 0))

    def kernel_times(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._chk(self.lib.hr_get_kernel_times(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"search": a.value, "warp": b.value, "pack": c.value}

    def launch_count(self):
        return int(self.lib.hr_get_launch_count(self.h))


    # ---- spatial bands (include/hopperrender_cuda.h, "spatial bands") ----
    def band_configure(self, rank, world, rows):
        r0 = (C.c_int * world)(*[a for a, _ in rows])
        r1 = (C.c_int * world)(*[b for _, b in rows])
        self._chk(self.lib.hr_band_configure(self.h, rank, world, r0, r1))
        self.band = (rank, world, list(rows))

    def band_set_max_radius(self, radius):
        self._chk(self.lib.hr_band_set_max_radius(self.h, int(radius)))

    def band_halo(self):
        """(first row, end row) of every frame this rank holds, bytes fetched from peers over NVLink so far."""
        lo, hi, n = C.c_int(), C.c_int(), C.c_ulonglong()
        self._chk(self.lib.hr_band_get_halo(self.h, C.byref(lo), C.byref(hi), C.byref(n)))
        return lo.value, hi.value, n.value

    def band_local_pointers(self):
        a, b, m = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._chk(self.lib.hr_band_local_pointers(self.h, C.byref(a), C.byref(b), C.byref(m)))
        return a.value, b.value, m.value

    def band_export_ipc(self):
        buf = C.create_string_buffer(3 * 64)
        self._chk(self.lib.hr_band_export_ipc(self.h, buf))
        return bytes(buf.raw)

    def band_open_ipc(self, handles):
        a, b, m = C.c_void_p(), C.c_void_p(), C.c_void_p()
        buf = C.create_string_buffer(handles, 3 * 64)
        self._chk(self.lib.hr_band_open_ipc(self.h, buf, C.byref(a), C.byref(b), C.byref(m)))
        return a.value, b.value, m.value

    def band_connect(self, peer_rank, ptrs, peer_device=-1):
        self._chk(self.lib.hr_band_connect(self.h, peer_rank, peer_device, C.c_void_p(ptrs[0]), C.c_void_p(ptrs[1]), C.c_void_p(ptrs[2])))

    def band_upload(self, y_band, uv_band, device=False):
        self._chk(self.lib.hr_band_upload(self.h, _ptr(y_band), _ptr(uv_band), 1 if device else 0))

    def band_gather(self, blocking=True):
        self._chk(self.lib.hr_band_gather(self.h, 1 if blocking else 0))

    def band_download(self, y_band, uv_band):
        sec = C.c_double(0.0)
        self._chk(self.lib.hr_band_download(self.h, _ptr(y_band), _ptr(uv_band), C.byref(sec)))
        return sec.value


class BandGroup:
    """One frame stream split into spatial bands over several contexts IN ONE PROCESS, one per entry of `devices`
    (different GPUs: the searches of the group wait for one another on the device, include/hopperrender_cuda.h).
    Presents the six calls of the optical-flow-calc interface for whole frames: updateFrame uploads every band to its
    own GPU and fetches the halos by P2P, every GPU searches its own lattice tiles (the flow ends up whole and
    bit-identical on all of them), warp + download work band by band and write straight into the caller's full-size
    planes. SURVEY.md §8e."""

    def __init__(self, frameHeight, frameWidth, actualWidth=None, pixfmt=PIXFMT_NV12, devices=(0, 1), max_radius=None):
        from . import sharding
        if len(set(devices)) != len(devices):
            raise ValueError("a band group needs one GPU per band (the searches wait for one another on the device)")
        self.H, self.W = frameHeight, frameWidth
        self.world = len(devices)
        self.ctx = [HrCuda(frameHeight, frameWidth, actualWidth, pixfmt, d) for d in devices]
        self.rows = sharding.band_rows(frameHeight, self.world, self.ctx[0].info.resScalar)
        for r, c in enumerate(self.ctx):
            c.band_configure(r, self.world, self.rows)
            if max_radius is not None:
                c.band_set_max_radius(max_radius)
        ptrs = [c.band_local_pointers() for c in self.ctx]
        for r, c in enumerate(self.ctx):
            for p in range(self.world):
                if p != r:
                    c.band_connect(p, ptrs[p], devices[p])
        self.devices = list(devices)
        self.dtype = self.ctx[0].dtype

    def close(self):
        for c in self.ctx:
            c.close()

    def _band_views(self, y, uv, r):
        r0, r1 = self.rows[r]
        return y[r0:r1], uv[r0 >> 1:r1 >> 1]

    def update_frame(self, y, uv, blocking=True):
        for r, c in enumerate(self.ctx):          # phase 1 on every rank first
            c.band_upload(*self._band_views(y, uv, r))
        for c in self.ctx:                        # then phase 2
            c.band_gather(blocking=False)
        if blocking:
            for c in self.ctx:
                c.synchronize()

    def calc_flow(self, radius=MIN_SEARCH_RADIUS, deltaScalar=8, neighborBiasScalar=6):
        for c in self.ctx:
            c.calc_flow(radius, deltaScalar, neighborBiasScalar, blocking=False)
        for c in self.ctx:
            c.synchronize()

    def warp(self, t, mode=BlendedFrame, black=0.0, white=255.0):
        for c in self.ctx:
            c.warp(t, mode, black, white)

    def download(self, y=None, uv=None):
        if y is None:
            y = np.empty((self.H, self.W), self.dtype)
        if uv is None:
            uv = np.empty((self.H // 2, self.W), self.dtype)
        for r, c in enumerate(self.ctx):
            c.band_download(*self._band_views(y, uv, r))
        return y, uv


def connect_bands_distributed(ctx, dist, rows, max_radius=None):
    """One process per GPU (torchrun): configure `ctx` as band `rank` and map every peer's frame slots and
    exchange arena through CUDA IPC handles exchanged once, at set-up, over the process group (control plane
    only; halos move by P2P copies, the searches store into one another's arenas)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx.band_configure(rank, world, rows)
    if max_radius is not None:
        ctx.band_set_max_radius(max_radius)
    handles = [None] * world
    dist.all_gather_object(handles, ctx.band_export_ipc())
    for p in range(world):
        if p != rank:
            ctx.band_connect(p, ctx.band_open_ipc(handles[p]), -1)
    dist.barrier()


class OpticalFlowCalc:
    """Mirror of `struct OpticalFlowCalc` (video/filter/HopperRender/opticalFlowCalc.h:10-65):
    the fields the filter reads and writes keep their names; the cl_* members are replaced by
    one opaque handle."""

    def __init__(self):
        self.isInitialized = False
        self.frameWidth = 0
        self.frameHeight = 0
        self.actualWidth = 0
        self.outputBlackLevel = 0.0
        self.outputWhiteLevel = 255.0
        self.opticalFlowResScalar = 0
        self.opticalFlowFrameWidth = 0
        self.opticalFlowFrameHeight = 0
        self.opticalFlowSearchRadius = MIN_SEARCH_RADIUS
        self.ofcCalcTime = 0.0
        self.warpCalcTime = 0.0
        self.deltaScalar = 8
        self.neighborBiasScalar = 6
        self.pixfmt = PIXFMT_NV12
        self.impl = None


# The six functions of opticalFlowCalc.h:77-124. Return value: False (0) = success,
# True (1) = failure — the reference's inverted bool (opticalFlowCalc.c:11-15).

def initOpticalFlowCalc(ofc: OpticalFlowCalc, frameHeight: int, frameWidth: int, actualWidth: int, pixfmt: int = PIXFMT_NV12, device: int = -1) -> bool:
    try:
        impl = HrCuda(frameHeight, frameWidth, actualWidth, pixfmt, device)
    except (HrError, OSError, RuntimeError) as e:
        print("[HopperRender] init failed: %s" % e)
        return True
    try:
        impl.set_pipeline(True)      # as the C host does (opticalFlowCalc.c): pack beside the search, same results
    except HrError as e:
        print("[HopperRender] init failed: %s" % e)
        impl.close()
        return True
    ofc.impl = impl
    ofc.frameWidth, ofc.frameHeight, ofc.actualWidth = frameWidth, frameHeight, actualWidth
    ofc.outputBlackLevel, ofc.outputWhiteLevel = 0.0, 255.0      # opticalFlowCalc.c:328-329
    ofc.opticalFlowSearchRadius = MIN_SEARCH_RADIUS              # :330
    ofc.opticalFlowResScalar = impl.info.resScalar
    ofc.opticalFlowFrameWidth = impl.info.lowWidth
    ofc.opticalFlowFrameHeight = impl.info.lowHeight
    ofc.ofcCalcTime = ofc.warpCalcTime = 0.0
    ofc.deltaScalar, ofc.neighborBiasScalar = 8, 6               # :339-340
    ofc.pixfmt = pixfmt
    ofc.isInitialized = True
    return False


def freeOFC(ofc: OpticalFlowCalc) -> None:
    if ofc.impl is not None:
        ofc.impl.close()
        ofc.impl = None
    ofc.isInitialized = False


def updateFrame(ofc: OpticalFlowCalc, inputPlanes) -> bool:
    if not ofc.isInitialized:
        return True
    return bool(ofc.impl.lib.hr_update_frame(ofc.impl.h, _ptr(inputPlanes[0]), _ptr(inputPlanes[1])))


def downloadFrame(ofc: OpticalFlowCalc, outputPlanes) -> bool:
    if not ofc.isInitialized:
        return True
    sec = C.c_double(0.0)
    rc = ofc.impl.lib.hr_download(ofc.impl.h, _ptr(outputPlanes[0]), _ptr(outputPlanes[1]), C.byref(sec))
    if not rc:
        ofc.warpCalcTime = sec.value
    return bool(rc)


def calculateOpticalFlow(ofc: OpticalFlowCalc) -> bool:
    if not ofc.isInitialized:
        return True
    sec = C.c_double(0.0)
    rc = ofc.impl.lib.hr_calc_flow(ofc.impl.h, ofc.opticalFlowSearchRadius, ofc.deltaScalar, ofc.neighborBiasScalar, C.byref(sec))
    if not rc:
        ofc.ofcCalcTime = sec.value
    return bool(rc)


def warpFrames(ofc: OpticalFlowCalc, blendingScalar: float, frameOutputMode: int) -> bool:
    if not ofc.isInitialized:
        return True
    return bool(ofc.impl.lib.hr_warp(ofc.impl.h, float(blendingScalar), int(frameOutputMode), float(ofc.outputBlackLevel), float(ofc.outputWhiteLevel)))


# ---- the C host layer itself (mpv/video/filter/HopperRender/opticalFlowCalc.{h,c} + hrReplay.c), for callers that
# ---- want the compiled drop-in rather than this module's mirror of it (bench.py's end-to-end leg)
class COpticalFlowCalc(C.Structure):
    """`struct OpticalFlowCalc` as laid out by mpv/video/filter/HopperRender/opticalFlowCalc.h."""
    _fields_ = [
        ("isInitialized", C.c_bool), ("frameWidth", C.c_int), ("frameHeight", C.c_int), ("actualWidth", C.c_int),
        ("outputBlackLevel", C.c_float), ("outputWhiteLevel", C.c_float),
        ("opticalFlowResScalar", C.c_int), ("opticalFlowFrameWidth", C.c_int), ("opticalFlowFrameHeight", C.c_int),
        ("opticalFlowSearchRadius", C.c_int), ("ofcCalcTime", C.c_double), ("warpCalcTime", C.c_double),
        ("deltaScalar", C.c_int), ("neighborBiasScalar", C.c_int),
        ("pixelFormat", C.c_int), ("cudaDevice", C.c_int), ("impl", C.c_void_p),
    ]


class HrControlState(C.Structure):
    """`HrControlState` of mpv/video/filter/HopperRender/hrControl.h."""
    _fields_ = [("interpolationActive", C.c_int), ("frameOutputMode", C.c_int), ("restartCounters", C.c_int), ("pinnedRadius", C.c_int),
                ("pendingLength", C.c_int), ("pending", C.c_char * 28)]


_ofc_lib = None


def load_ofc_library():
    """dlopen libhopperrender_ofc.so (the C host layer, linked against the CUDA library)."""
    global _ofc_lib
    if _ofc_lib is not None:
        return _ofc_lib
    load_library()
    path = _build.OFC_LIB
    if not path.exists():
        _build.build_host()
    lib = C.CDLL(str(path))
    P = C.POINTER(COpticalFlowCalc)
    PP = C.POINTER(C.c_void_p)
    for name, res, args in (
        ("initOpticalFlowCalc", C.c_bool, [P, C.c_int, C.c_int, C.c_int]),
        ("freeOFC", None, [P]),
        ("updateFrame", C.c_bool, [P, PP]),
        ("downloadFrame", C.c_bool, [P, PP]),
        ("calculateOpticalFlow", C.c_bool, [P]),
        ("warpFrames", C.c_bool, [P, C.c_float, C.c_int]),
        ("allocHostPlanes", C.c_void_p, [C.c_size_t]),
        ("freeHostPlanes", None, [C.c_void_p, C.c_void_p]),
        ("hrControlParse", C.c_int, [C.c_char_p]),
        ("hrControlApply", C.c_int, [P, C.POINTER(HrControlState), C.c_int]),
        ("hrControlPoll", C.c_int, [C.c_int, P, C.POINTER(HrControlState)]),
        ("hrControlStatus", C.c_int, [C.c_char_p, C.c_size_t, P, C.c_double, C.c_double, C.c_double, C.c_double]),
        ("hrReplayStream", C.c_longlong, [P, PP, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int, PP]),
    ):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _ofc_lib = lib
    return lib


def replay_stream_c(ofc, frames, first_frame, ts_per_step, mode, out_planes):
    """hrReplayStream: the filter's call sequence for len(ts_per_step) source frames as one C loop.
    frames: list of (y, uv) host planes used round-robin; out_planes: (y, uv). Returns outputs delivered."""
    lib = load_ofc_library()
    n = len(frames)
    planes = (C.c_void_p * (2 * n))(*[_ptr(p).value for f in frames for p in f])
    flat = [float(t) for step in ts_per_step for t in step]
    ts = (C.c_float * max(1, len(flat)))(*flat)
    starts, acc = [0], 0
    for step in ts_per_step:
        acc += len(step)
        starts.append(acc)
    st = (C.c_int * len(starts))(*starts)
    outs = (C.c_void_p * 2)(_ptr(out_planes[0]).value, _ptr(out_planes[1]).value)
    got = lib.hrReplayStream(C.byref(ofc), planes, n, first_frame, len(ts_per_step), ts, st, int(mode), outs)
    if got < 0:
        raise HrError("hrReplayStream: a call of the optical-flow-calc interface failed")
    return int(got)
