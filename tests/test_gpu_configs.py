"""GPU parity at the BASELINE.json configurations that the base parity file does not reach:

* 8K (resolution scalar 4, opticalFlowCalc.c:331-336): a full 7680x4320 P010 pair and a narrow 256x4320 strip,
* lattices with more tiles than SMs (2560x1080, 3840x1080, 5120x2160: the search CTAs then own several tiles),
* 4K P010 in the HSV flow mode and the 24->144 pacing (6,1,6,6 warps per source frame, one of
  them at a blending scalar one ulp below 1.0), through the compiled C host layer,
* frames taller than 4320 lines are refused at creation.

Tolerance: flow bit-exact; pixels bit-identical to the oracle (which, for P010, implements the same definition —
DESIGN.md §4 — so nothing but 0 would catch a regression); HSV mode +-1 8-bit step (atan2f/fmodf differ between
libm and CUDA), i.e. +-256 on the 16-bit P010 sample.
"""
import ctypes as C

import numpy as np
import pytest

from test_gpu_parity import _assert_flow_equal, _diff_report, _run_pair, _warp_both

pytestmark = pytest.mark.gpu


def _check_modes(g, o, modes, ts, aw, is16, msgs, tag, black=0.0, white=255.0):
    for mode in modes:
        for t in ts:
            gy, guv, oy, ouv = _warp_both(g, o, t, mode, black, white)
            tol = (256 if is16 else 1) if mode == 3 else 0
            for nm, a, b in (("Y", gy, oy), ("UV", guv, ouv)):
                r = _diff_report("%s mode %d t=%r %s" % (tag, mode, float(t), nm), a[:, :aw], b[:, :aw], tol)
                if r:
                    msgs.append(r)


def _block_flow(lw, lh, seed, lo=-512, hi=393):
    rng = np.random.default_rng(seed)
    coarse = rng.integers(lo, hi, size=(2, (lh + 14) // 15, (lw + 15) // 16))
    flow = np.repeat(np.repeat(coarse, 15, axis=1), 16, axis=2)[:, :lh, :lw].astype(np.int16)
    flow[:, ::7, ::5] += rng.integers(-3, 4, size=flow[:, ::7, ::5].shape).astype(np.int16)
    return flow


def test_8k_p010_full_frame(hr, oracle, synth):
    """Config C4's geometry on one GPU: s = 4, 16x16-sample cells, pack_frame16_kernel<uint16_t, 4>."""
    w, h = 7680, 4320
    c = synth.MovingTextureClip(w, h, pixfmt=1)
    g, o = _run_pair(hr, oracle, c.frame(1), c.frame(2), h, w, w, 5, pixfmt=1)
    assert g.info.resScalar == 4 and (g.info.lowWidth, g.info.lowHeight) == (480, 270)
    _assert_flow_equal(g, o)
    assert np.abs(g.get_offsets()[0]).max() > 0
    msgs = []
    _check_modes(g, o, range(7), (0.4,), w, True, msgs, "8K P010")
    _check_modes(g, o, (2,), (0.0, np.float32(0.8), 1.0), w, True, msgs, "8K P010")
    _check_modes(g, o, (2, 5), (0.6,), w, True, msgs, "8K P010 16/219", 16.0, 219.0)
    flow = _block_flow(480, 270, 21)
    g.set_blurred_offsets(flow)
    o.set_blurred_offsets(flow)
    _check_modes(g, o, (0, 1, 2, 5, 6), (0.3,), w, True, msgs, "8K P010 frame-wide flow")
    assert not msgs, "\n".join(msgs)


@pytest.mark.parametrize("pixfmt", [0, 1])
@pytest.mark.parametrize("R", [5, 16])
def test_8k_narrow_strip(hr, oracle, synth, pixfmt, R):
    """s = 4 again, cheap: a 256x4320 strip (16x270 lattice) at both ends of the radius range, every mode."""
    w, h = 256, 4320
    c = synth.MovingTextureClip(w, h, pixfmt=pixfmt, velocity=(16, 48), fg_velocity=(-16, 32))
    g, o = _run_pair(hr, oracle, c.frame(1), c.frame(2), h, w, w, R, pixfmt=pixfmt)
    assert g.info.resScalar == 4
    _assert_flow_equal(g, o)
    msgs = []
    _check_modes(g, o, range(7), (0.0, 0.4, 1.0), w, pixfmt == 1, msgs, "strip")
    flow = _block_flow(g.info.lowWidth, g.info.lowHeight, 22, -300, 300)
    g.set_blurred_offsets(flow)
    o.set_blurred_offsets(flow)
    _check_modes(g, o, (0, 1, 2), (0.7,), w, pixfmt == 1, msgs, "strip frame-wide flow")
    assert not msgs, "\n".join(msgs)


@pytest.mark.parametrize("w,h,pixfmt,R", [(2560, 1080, 0, 5), (2560, 1080, 0, 16), (3840, 1080, 0, 7), (5120, 2160, 1, 5), (2560, 1080, 0, 3)])
def test_lattices_with_more_tiles_than_sms(hr, oracle, synth, w, h, pixfmt, R):
    """Ultrawide frames: 640x270 / 960x270 lattices = 180 / 270 tiles of 32x32 on 148 SMs, so search CTAs own more
    than one tile (the MULTI instantiation of the search kernel). Raw and blurred offsets, every step's winners."""
    c = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    g, o = _run_pair(hr, oracle, c.frame(2), c.frame(3), h, w, w, R, pixfmt=pixfmt)
    tiles = ((g.info.lowWidth + 31) // 32) * ((g.info.lowHeight + 31) // 32)
    assert g.info.searchCtas < tiles, "this geometry was meant to need several tiles per CTA"
    _assert_flow_equal(g, o)
    assert np.abs(g.get_offsets()[0]).max() > 0
    msgs = []
    _check_modes(g, o, (2,), (0.4,), w, pixfmt == 1, msgs, "ultrawide")
    # twice more on the same context: the epoch-tagged words of the previous launch must read as stale
    g.calc_flow(R)
    g.calc_flow(R)
    _assert_flow_equal(g, o)
    assert not msgs, "\n".join(msgs)


def test_4k_p010_hsv_and_levels(hr, oracle, synth):
    """Config C3's second half: 4K P010 in the HSV flow mode (and the grey / side-by-side-2 modes) with the default and
    a preset level pair."""
    w, h = 3840, 2160
    c = synth.MovingTextureClip(w, h, pixfmt=1)
    g, o = _run_pair(hr, oracle, c.frame(2), c.frame(3), h, w, w, 5, pixfmt=1)
    _assert_flow_equal(g, o)
    msgs = []
    _check_modes(g, o, (3, 4, 6), (0.0, 0.5, np.float32(1.0 / 6.0)), w, True, msgs, "4K P010")
    _check_modes(g, o, (3, 6), (0.5,), w, True, msgs, "4K P010 16/219", 16.0, 219.0)
    flow = _block_flow(480, 270, 23)
    g.set_blurred_offsets(flow)
    o.set_blurred_offsets(flow)
    _check_modes(g, o, (3, 4, 6), (0.3,), w, True, msgs, "4K P010 frame-wide flow")
    assert not msgs, "\n".join(msgs)


@pytest.mark.parametrize("mode", [2, 3])
def test_4k_p010_24_to_144_replay_through_the_c_host(hr, oracle, synth, mode):
    """The filter's call order for a 24 -> 144 stream (6, 1, 6, 6 warps, one at t = 0.99999994)
    through libhopperrender_ofc.so (hrReplayStream = updateFrame / calculateOpticalFlow / warpFrames / downloadFrame),
    source frame by source frame; the last delivered frame of every source frame against the oracle, which replays the
    same calls."""
    from hopperrender_b200 import pacing

    w, h = 3840, 2160
    c = synth.MovingTextureClip(w, h, pixfmt=1)
    frames = [c.frame(k) for k in range(5)]
    ts = pacing.schedule(5, 24.0, 144.0)
    assert [len(t) for t in ts] == [0, 6, 1, 6, 6]
    assert any(0.999 < t < 1.0 for step in ts for t in step)
    lib = hr.load_ofc_library()
    ofc = hr.COpticalFlowCalc()
    ofc.pixelFormat = 1
    assert not lib.initOpticalFlowCalc(C.byref(ofc), h, w, w)
    o = oracle.Oracle(h, w, w, 1)
    oy, ouv = np.zeros((h, w), np.uint16), np.zeros((h // 2, w), np.uint16)
    tol = 256 if mode == 3 else 0
    msgs = []
    total = 0
    for k in range(5):
        got = hr.replay_stream_c(ofc, frames, k, [ts[k]], mode, (oy, ouv)) if k else hr.replay_stream_c(ofc, frames, 0, [[]], mode, (oy, ouv))
        assert got == len(ts[k])
        total += got
        o.update_frame(*frames[k])
        if k == 0:
            continue
        o.calc_flow(ofc.opticalFlowSearchRadius, ofc.deltaScalar, ofc.neighborBiasScalar)
        o.warp(np.float32(ts[k][-1]), mode)
        ey, euv = o.download()
        for nm, a, b in (("Y", oy, ey), ("UV", ouv, euv)):
            r = _diff_report("source frame %d last t=%r mode %d %s" % (k, ts[k][-1], mode, nm), a, b, tol)
            if r:
                msgs.append(r)
        raw = np.empty((2, 270, 480), np.int16)
        assert hr.load_library().hr_get_offsets(C.c_void_p(ofc.impl), C.c_void_p(raw.ctypes.data), None) == 0
        assert np.array_equal(raw, o.get_offsets()[0]), "raw offsets differ at source frame %d" % k
    assert total == 19
    lib.freeOFC(C.byref(ofc))
    assert not msgs, "\n".join(msgs)


def test_frames_taller_than_8k_are_refused(hr):
    with pytest.raises(hr.HrError, match="not supported"):
        hr.HrCuda(4352, 256, 256)
    hr.HrCuda(4320, 64, 64).close()
