"""The drop-in claim, executed: the REFERENCE'S OWN FILTER (video/filter/HopperRender/vf_HopperRender.c,
compiled unmodified by oracle/build_ref.py) drives this repository's optical-flow-calc layer and CUDA
library through the six calls of opticalFlowCalc.h. oracle/filter_host_sim.c stands in for mpv's filter
runtime (single-slot pins, pass-through autoconvert, --vo-null-fps style display rate).

Checked: the number of frames the filter emits per source frame, their PTS, and
their pixels against direct C-ABI calls with the blend positions of the pacing replay."""
import ctypes as C
import pathlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent
SIM = ROOT / "oracle" / "_ref" / "libhr_filter_sim.so"


@pytest.fixture(scope="module")
def sim():
    if not SIM.exists():
        pytest.skip("oracle/_ref/libhr_filter_sim.so not built (python oracle/build_ref.py where /root/reference exists)")
    L = C.CDLL(str(SIM))
    L.hr_sim_create.restype = C.c_void_p
    L.hr_sim_create.argtypes = [C.c_int, C.c_double]
    L.hr_sim_push.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
    L.hr_sim_pop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.hr_sim_destroy.argtypes = [C.c_void_p]
    L.hr_sim_reset.argtypes = [C.c_void_p]
    L.hr_sim_command_speed.argtypes = [C.c_void_p, C.c_double]
    L.hr_sim_filter_name.restype = C.c_char_p
    return L


def _pop_all(L, s, w, h):
    outs = []
    while True:
        y = np.empty((h, w), np.uint8)
        uv = np.empty((h // 2, w), np.uint8)
        pts = C.c_double()
        if L.hr_sim_pop(s, C.c_void_p(y.ctypes.data), C.c_void_p(uv.ctypes.data), C.byref(pts), None):
            return outs
        outs.append((y, uv, pts.value))


@pytest.mark.parametrize("mode,display_fps", [(2, 60.0), (0, 60.0), (2, 144.0)])
def test_reference_filter_runs_on_the_cuda_path(hr, synth, sim, mode, display_fps):
    from hopperrender_b200 import pacing
    assert sim.hr_sim_filter_name() == b"HopperRender"
    w, h, fps = 1280, 720, 24.0
    clip = synth.MovingTextureClip(w, h)
    s = sim.hr_sim_create(mode, display_fps)
    assert s
    direct = hr.HrCuda(h, w, w)
    pacer = pacing.Pacer(fps, display_fps)
    total = 0
    for k in range(5):
        y, uv = clip.frame(k)
        n = sim.hr_sim_push(s, C.c_void_p(y.ctypes.data), C.c_void_p(uv.ctypes.data), w, h, k / fps, fps)
        assert n >= 0, "the filter marked itself failed"
        outs = _pop_all(sim, s, w, h)
        ts = pacer.next_source_frame()
        direct.update_frame(y, uv)
        if k == 0:
            # vf_HopperRender.c:490-501: the first source frame is delivered as it came
            assert len(outs) == 1 and np.array_equal(outs[0][0], y) and np.array_equal(outs[0][1], uv)
            assert outs[0][2] == 0.0
            continue
        assert len(outs) == len(ts), "frame %d: %d outputs, pacing rule gives %d" % (k, len(outs), len(ts))
        direct.calc_flow(5)                      # radius pinned (AUTO_SEARCH_RADIUS_ADJUST 0 in the harness build)
        for (oy, ouv, pts), t in zip(outs, ts):
            direct.warp(np.float32(t), mode)
            ey, euv, _ = direct.download()
            assert np.array_equal(oy, ey) and np.array_equal(ouv, euv), "frame %d t=%.3f differs from the direct calls" % (k, t)
            total += 1
        # PTS (vf_HopperRender.c:464-477 source frames, :389-390 intermediate frames): the source-frame PTS snaps to the incoming PTS for the first two frames, then advances by 1/display_fps
        assert abs(outs[0][2] - k / fps) < 0.05
        for a, b in zip(outs, outs[1:]):
            assert abs((b[2] - a[2]) - 1.0 / display_fps) < 1e-9
    assert total == sum(len(x) for x in pacing.schedule(5, fps, display_fps))
    sim.hr_sim_destroy(s)
    direct.close()


def test_reset_and_speed_command(hr, synth, sim):
    """Seek (reset, vf_HopperRender.c:562-567) restarts the two-frame warm-up; a speed change
    (:541-555) shortens the source frame time, so fewer frames are interpolated."""
    w, h = 1280, 720
    clip = synth.MovingTextureClip(w, h)
    s = sim.hr_sim_create(2, 60.0)
    counts = []
    for k in range(3):
        y, uv = clip.frame(k)
        counts.append(sim.hr_sim_push(s, C.c_void_p(y.ctypes.data), C.c_void_p(uv.ctypes.data), w, h, k / 24.0, 24.0))
        _pop_all(sim, s, w, h)
    assert counts == [1, 3, 2]
    sim.hr_sim_reset(s)
    y, uv = clip.frame(3)
    assert sim.hr_sim_push(s, C.c_void_p(y.ctypes.data), C.c_void_p(uv.ctypes.data), w, h, 3 / 24.0, 24.0) == 1   # warm-up again
    _pop_all(sim, s, w, h)
    sim.hr_sim_command_speed(s, 2.0)            # 48 source frames per second against a 60 Hz display
    y, uv = clip.frame(4)
    n = sim.hr_sim_push(s, C.c_void_p(y.ctypes.data), C.c_void_p(uv.ctypes.data), w, h, 4 / 24.0, 24.0)
    assert n in (1, 2)
    sim.hr_sim_destroy(s)


SIM_P010 = ROOT / "oracle" / "_ref" / "libhr_filter_sim_p010.so"


def test_reference_filter_p010(hr, oracle, synth):
    """the reference's filter source, compiled with the P010 change of INTEGRATION.md §1 stated as
    a macro (oracle/mpv_shim/hr_p010_patch.h), driving the P010 path: frame counts of the pacing rule, and every
    delivered frame against direct C-ABI calls and against the CPU oracle (+-4 LSB of the 10-bit value)."""
    from hopperrender_b200 import pacing
    if not SIM_P010.exists():
        pytest.skip("oracle/_ref/libhr_filter_sim_p010.so not built")
    L = C.CDLL(str(SIM_P010))
    L.hr_sim_create.restype = C.c_void_p
    L.hr_sim_create.argtypes = [C.c_int, C.c_double]
    L.hr_sim_push.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
    L.hr_sim_pop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.hr_sim_destroy.argtypes = [C.c_void_p]
    assert L.hr_sim_bytes_per_sample() == 2
    w, h, fps, disp = 1280, 720, 24.0, 60.0
    clip = synth.MovingTextureClip(w, h, pixfmt=1)
    s = L.hr_sim_create(2, disp)
    assert s
    direct = hr.HrCuda(h, w, w, 1)
    o = oracle.Oracle(h, w, w, 1)
    pacer = pacing.Pacer(fps, disp)
    checked = 0
    for k in range(4):
        y, uv = clip.frame(k)
        assert y.dtype == np.uint16
        n = L.hr_sim_push(s, C.c_void_p(y.ctypes.data), C.c_void_p(uv.ctypes.data), w, h, k / fps, fps)
        assert n >= 0, "the filter marked itself failed"
        outs = []
        while True:
            oy, ouv = np.empty((h, w), np.uint16), np.empty((h // 2, w), np.uint16)
            if L.hr_sim_pop(s, C.c_void_p(oy.ctypes.data), C.c_void_p(ouv.ctypes.data), None, None):
                break
            outs.append((oy, ouv))
        ts = pacer.next_source_frame()
        direct.update_frame(y, uv)
        o.update_frame(y, uv)
        if k == 0:
            assert len(outs) == 1 and np.array_equal(outs[0][0], y) and np.array_equal(outs[0][1], uv)
            continue
        assert len(outs) == len(ts)
        direct.calc_flow(5)
        o.calc_flow(5, 8, 6)
        for (oy, ouv), t in zip(outs, ts):
            direct.warp(np.float32(t), 2)
            ey, euv, _ = direct.download()
            assert np.array_equal(oy, ey) and np.array_equal(ouv, euv), "frame %d t=%.3f differs from the direct calls" % (k, t)
            assert o.warp(np.float32(t), 2, 0.0, 255.0) == 0
            ry, ruv = o.download()
            assert np.abs(oy.astype(np.int32) - ry.astype(np.int32)).max() <= 4 * 64
            assert np.abs(ouv.astype(np.int32) - ruv.astype(np.int32)).max() <= 4 * 64
            checked += 1
    assert checked == 8
    L.hr_sim_destroy(s)
    direct.close()
