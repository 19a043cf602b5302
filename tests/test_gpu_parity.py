"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle
on the same seeded inputs. Integer flow offsets must be bit-exact; NV12 pixels must be equal for
the non-HSV modes (the CUDA kernels round after every float operation exactly like the oracle)
and within +-1 LSB for HSV (atan2f/fmodf implementations differ); P010 likewise identical (the oracle implements the
same P010 definition), HSV within one 8-bit step.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _diff_report(name, got, exp, tol=0):
    d = np.abs(got.astype(np.int64) - exp.astype(np.int64))
    bad = np.argwhere(d > tol)
    if bad.size == 0:
        return None
    first = tuple(int(v) for v in bad[0])
    return "%s: %d of %d differ (max |d|=%d); first at %s got=%d exp=%d; bbox=%s..%s" % (
        name, len(bad), d.size, int(d.max()), first, int(got[first]), int(exp[first]),
        bad.min(axis=0).tolist(), bad.max(axis=0).tolist())


def _run_pair(hr, oracle, f1, f2, H, W, aW, R, dS=8, nS=6, pixfmt=0):
    g = hr.HrCuda(H, W, aW, pixfmt)
    g.set_trace(True)
    g.update_frame(*f1)
    g.update_frame(*f2)
    g.calc_flow(R, dS, nS)
    o = oracle.Oracle(H, W, aW, pixfmt)
    o.update_frame(*f1)
    o.update_frame(*f2)
    o.calc_flow(R, dS, nS)
    return g, o


def _assert_flow_equal(g, o):
    graw, gblur = g.get_offsets()
    oraw, oblur = o.get_offsets()
    msgs = []
    if not np.array_equal(graw, oraw):
        first = 256
        lw, lh = o.lw, o.lh
        m = max(lw, lh)
        ws0 = 1
        while ws0 < m:
            ws0 <<= 1
        ws0 //= 2
        for step in range(o.steps):
            ws = ws0 >> (step // 2)
            a = g.get_step_layers(step)[::ws, ::ws]
            b = o.get_step_layers(step)[::ws, ::ws]
            if not np.array_equal(a, b):
                msgs.append("first differing search step %d (window %d): %s" % (step, ws, _diff_report("layers", a, b)))
                break
        msgs.append(_diff_report("raw offsets", graw, oraw))
    r = _diff_report("blurred offsets", gblur, oblur)
    if r:
        msgs.append(r)
    assert not msgs, "\n".join(m for m in msgs if m)


GEOMS = [
    # (w, h, stride) — SURVEY.md Appendix B rows + a 270-line clip (s = 0) + ragged lattice
    (1920, 1080, 1920),
    (1280, 720, 1280),
    (854, 480, 896),
    (480, 270, 480),
    (1000, 562, 1024),
]


@pytest.mark.parametrize("w,h,stride", GEOMS)
@pytest.mark.parametrize("R", [5, 16])
def test_flow_bit_exact(hr, oracle, synth, w, h, stride, R):
    c = synth.MovingTextureClip(w, h, stride=stride)
    g, o = _run_pair(hr, oracle, c.frame(2), c.frame(3), h, stride, w, R)
    _assert_flow_equal(g, o)


@pytest.mark.parametrize("R,dS,nS", [(2, 8, 6), (6, 8, 6), (8, 0, 0), (9, 12, 10), (13, 4, 8), (32, 8, 6)])
def test_flow_bit_exact_radius_and_scalars(hr, oracle, synth, R, dS, nS):
    c = synth.MovingTextureClip(1920, 1080)
    g, o = _run_pair(hr, oracle, c.frame(5), c.frame(6), 1080, 1920, 1920, R, dS, nS)
    _assert_flow_equal(g, o)


def test_flow_bit_exact_4k(hr, oracle, synth):
    c = synth.MovingTextureClip(3840, 2160)
    g, o = _run_pair(hr, oracle, c.frame(1), c.frame(2), 2160, 3840, 3840, 5)
    _assert_flow_equal(g, o)


def test_flow_adversarial_inputs(hr, oracle, synth):
    # full-range noise with a large deltaScalar: uint32 window sums wrap
    f1, f2 = synth.noise_frame(1080, 1920, 11), synth.noise_frame(1080, 1920, 12)
    g, o = _run_pair(hr, oracle, f1, f2, 1080, 1920, 1920, 5, 12, 6)
    _assert_flow_equal(g, o)
    # static pair -> zero flow; constant grey -> ties resolved to the zero candidate
    c = synth.MovingTextureClip(1280, 720)
    g, o = _run_pair(hr, oracle, c.frame(0), c.frame(0), 720, 1280, 1280, 5)
    _assert_flow_equal(g, o)
    assert not g.get_offsets()[0].any()
    y = np.full((720, 1280), 90, np.uint8)
    uv = np.full((360, 1280), 128, np.uint8)
    g, o = _run_pair(hr, oracle, (y, uv), (y, uv), 720, 1280, 1280, 6)
    _assert_flow_equal(g, o)
    assert not g.get_offsets()[0].any()


def test_flow_is_repeatable_and_stateless(hr, oracle, synth):
    """calculateOpticalFlow zeroes the offsets every call (opticalFlowCalc.c:153): two calls on
    the same pair, and a call after a different radius, give identical results."""
    c = synth.MovingTextureClip(1920, 1080)
    g = hr.HrCuda(1080, 1920, 1920)
    g.update_frame(*c.frame(0))
    g.update_frame(*c.frame(1))
    g.calc_flow(5)
    a = g.get_offsets()
    g.calc_flow(16)
    g.calc_flow(5)
    b = g.get_offsets()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_blur_tap(hr, oracle):
    g = hr.HrCuda(1080, 1920, 1920)
    rng = np.random.default_rng(3)
    raw = rng.integers(-512, 400, size=(2, 270, 480), dtype=np.int16)
    assert np.array_equal(g.blur_flow(raw), oracle.blur_flow(raw))


def _warp_both(g, o, t, mode, black=0.0, white=255.0):
    g.warp(t, mode, black, white)
    gy, guv, _ = g.download()
    assert o.warp(t, mode, black, white) == 0
    oy, ouv = o.download()
    return gy, guv, oy, ouv


# (1918, 1080, 1920): encoded width not a multiple of 4 -> partial columns of the block kernel; 4K: 8-row blocks
@pytest.mark.parametrize("w,h,stride", [(1920, 1080, 1920), (854, 480, 896), (480, 270, 480), (1918, 1080, 1920), (3840, 2160, 3840)])
def test_warp_all_modes(hr, oracle, synth, w, h, stride):
    c = synth.MovingTextureClip(w, h, stride=stride)
    g, o = _run_pair(hr, oracle, c.frame(2), c.frame(3), h, stride, w, 8)
    _assert_flow_equal(g, o)
    msgs = []
    for mode in range(7):
        for t in ((0.0, 0.4, np.float32(0.8), 1.0) if h <= 1080 else (0.4,)):
            gy, guv, oy, ouv = _warp_both(g, o, t, mode)
            tol = 1 if mode == 3 else 0
            for nm, a, b in (("Y", gy, oy), ("UV", guv, ouv)):
                r = _diff_report("mode %d t=%.2f %s" % (mode, t, nm), a[:, :w], b[:, :w], tol)
                if r:
                    msgs.append(r)
    assert not msgs, "\n".join(msgs)


def test_warp_levels_presets(hr, oracle, synth):
    c = synth.MovingTextureClip(1920, 1080)
    g, o = _run_pair(hr, oracle, c.frame(2), c.frame(3), 1080, 1920, 1920, 5)
    msgs = []
    for black, white in ((10.0, 219.0), (16.0, 219.0), (0.0, 200.0), (30.0, 255.0)):   # vf_HopperRender.c:160-170
        for mode in (2, 3, 5):
            gy, guv, oy, ouv = _warp_both(g, o, 0.6, mode, black, white)
            tol = 1 if mode == 3 else 0
            for nm, a, b in (("Y", gy, oy), ("UV", guv, ouv)):
                r = _diff_report("levels %g/%g mode %d %s" % (black, white, mode, nm), a, b, tol)
                if r:
                    msgs.append(r)
    assert not msgs, "\n".join(msgs)


def test_warp_with_large_synthetic_flow(hr, oracle, synth):
    """Flow injected through the tap: vectors up to the R=16 reach (-512/+392), so the flip
    indirection, both mirrors and the clamps are exercised everywhere, not only at the border."""
    c = synth.MovingTextureClip(1920, 1080)
    g, o = _run_pair(hr, oracle, c.frame(0), c.frame(1), 1080, 1920, 1920, 5)
    rng = np.random.default_rng(5)
    coarse = rng.integers(-512, 393, size=(2, 18, 30))
    flow = np.repeat(np.repeat(coarse, 15, axis=1), 16, axis=2).astype(np.int16)
    flow[:, ::7, ::5] += rng.integers(-3, 4, size=flow[:, ::7, ::5].shape).astype(np.int16)
    g.set_blurred_offsets(flow)
    o.set_blurred_offsets(flow)
    msgs = []
    for mode in (0, 1, 2, 4, 5, 6):
        for t in (0.2, 0.6):
            gy, guv, oy, ouv = _warp_both(g, o, t, mode)
            for nm, a, b in (("Y", gy, oy), ("UV", guv, ouv)):
                r = _diff_report("mode %d t=%.1f %s" % (mode, t, nm), a, b, 0)
                if r:
                    msgs.append(r)
    assert not msgs, "\n".join(msgs)


def test_p010_flow_and_warp(hr, oracle, synth):
    c = synth.MovingTextureClip(1920, 1080, pixfmt=1)
    g, o = _run_pair(hr, oracle, c.frame(2), c.frame(3), 1080, 1920, 1920, 5, pixfmt=1)
    _assert_flow_equal(g, o)
    msgs = []
    for mode in range(7):
        gy, guv, oy, ouv = _warp_both(g, o, 0.4, mode)
        # oracle and CUDA implement the same P010 definition: identical, except the HSV hue (atan2f of
        # libm vs CUDA: one 8-bit step, stored << 8). north_star's +-4 LSB10 is the bound against a 16-bit reading of
        # the reference; against our own oracle anything but 0 would hide a regression.
        tol = 256 if mode == 3 else 0
        for nm, a, b in (("Y", gy, oy), ("UV", guv, ouv)):
            r = _diff_report("P010 mode %d %s" % (mode, nm), a, b, tol)
            if r:
                msgs.append(r)
    assert not msgs, "\n".join(msgs)


def test_reference_call_order_and_error_convention(hr, oracle, synth):
    """Replays vf_HopperRender.c:445-496 through the mirrored six-function interface; every
    function returns False (0) on success and True on failure."""
    c = synth.MovingTextureClip(1280, 720)
    ofc = hr.OpticalFlowCalc()
    y, uv = c.frame(0)
    assert hr.updateFrame(ofc, [y, uv]) is True            # not initialised -> failure (opticalFlowCalc.c:97)
    assert hr.calculateOpticalFlow(ofc) is True
    assert hr.warpFrames(ofc, 0.5, 2) is True
    assert hr.initOpticalFlowCalc(ofc, 720, 1280, 1280) is False
    assert (ofc.opticalFlowResScalar, ofc.opticalFlowFrameWidth, ofc.opticalFlowFrameHeight) == (2, 320, 180)
    assert ofc.opticalFlowSearchRadius == 5 and ofc.deltaScalar == 8 and ofc.neighborBiasScalar == 6
    o = oracle.Oracle(720, 1280)
    from hopperrender_b200 import pacing
    p = pacing.Pacer(24.0, 60.0)
    outs = 0
    for k in range(4):
        y, uv = c.frame(k)
        ts = p.next_source_frame()
        assert hr.updateFrame(ofc, [y, uv]) is False
        o.update_frame(y, uv)
        if k >= 1:
            assert hr.calculateOpticalFlow(ofc) is False
            o.calc_flow(ofc.opticalFlowSearchRadius, ofc.deltaScalar, ofc.neighborBiasScalar)
            assert ofc.ofcCalcTime > 0.0
        for t in ts:
            oy, ouv = np.empty_like(y), np.empty_like(uv)
            assert hr.warpFrames(ofc, t, hr.BlendedFrame) is False
            assert hr.downloadFrame(ofc, [oy, ouv]) is False
            assert ofc.warpCalcTime > 0.0
            o.warp(np.float32(t), 2)
            ey, euv = o.download()
            assert np.array_equal(oy, ey) and np.array_equal(ouv, euv)
            outs += 1
    assert outs == 3 + 2 + 3
    assert hr.warpFrames(ofc, 1.5, 2) is True               # opticalFlowCalc.c:209-212
    hr.freeOFC(ofc)
    assert ofc.isInitialized is False


def test_device_resident_update_matches_host_update(hr, synth):
    import torch
    c = synth.MovingTextureClip(1920, 1080)
    f0, f1 = c.frame(0), c.frame(1)
    a = hr.HrCuda(1080, 1920, 1920)
    a.update_frame(*f0)
    a.update_frame(*f1)
    a.calc_flow(5)
    a.warp(0.4, 2)
    ay, auv, _ = a.download()
    for borrow in (False, True):
        b = hr.HrCuda(1080, 1920, 1920)
        d = [(torch.from_numpy(f[0]).cuda(), torch.from_numpy(f[1]).cuda()) for f in (f0, f1)]
        torch.cuda.synchronize()
        for dy, duv in d:
            b.update_frame_device(dy, duv, borrow)
        b.calc_flow(5)
        b.warp(0.4, 2)
        by, buv, _ = b.download()
        assert np.array_equal(a.get_offsets()[1], b.get_offsets()[1])
        assert np.array_equal(ay, by) and np.array_equal(auv, buv)


@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 2160)])
def test_p010_levels_and_large_flow(hr, oracle, synth, w, h):
    """P010 with the level presets (the 16-bit black level is then not an integer: the unfolded subtraction and the
    float clamp of the block kernel), with out-of-range samples (low 6 bits set, values above 1023 << 6: the 16-bit
    clamp) and with a flow that reaches across the frame."""
    c = synth.MovingTextureClip(w, h, pixfmt=1)
    f2, f3 = c.frame(2), c.frame(3)
    rng = np.random.default_rng(11)
    noisy = []
    for y, uv in (f2, f3):
        y, uv = y.copy(), uv.copy()
        y[::3, ::5] |= rng.integers(0, 64, size=y[::3, ::5].shape).astype(np.uint16)
        y[5::97, 7::89] = 65535
        uv[::5, ::3] |= rng.integers(0, 64, size=uv[::5, ::3].shape).astype(np.uint16)
        uv[3::53, 2::61] = 65535
        uv[9::53, 5::61] = 0
        noisy.append((y, uv))
    g, o = _run_pair(hr, oracle, noisy[0], noisy[1], h, w, w, 5, pixfmt=1)
    _assert_flow_equal(g, o)
    msgs = []
    tol = 0
    for black, white in ((0.0, 255.0), (16.0, 219.0), (10.0, 219.0), (30.0, 255.0)):
        for mode in (2, 5):
            gy, guv, oy, ouv = _warp_both(g, o, 0.6, mode, black, white)
            for nm, a, b in (("Y", gy, oy), ("UV", guv, ouv)):
                r = _diff_report("P010 levels %g/%g mode %d %s" % (black, white, mode, nm), a, b, tol)
                if r:
                    msgs.append(r)
    lw, lh = g.info.lowWidth, g.info.lowHeight
    coarse = rng.integers(-512, 393, size=(2, (lh + 14) // 15, (lw + 15) // 16))
    flow = np.repeat(np.repeat(coarse, 15, axis=1), 16, axis=2)[:, :lh, :lw].astype(np.int16)
    g.set_blurred_offsets(flow)
    o.set_blurred_offsets(flow)
    for mode in (0, 1, 2):
        gy, guv, oy, ouv = _warp_both(g, o, 0.3, mode, 16.0, 219.0)
        for nm, a, b in (("Y", gy, oy), ("UV", guv, ouv)):
            r = _diff_report("P010 large flow mode %d %s" % (mode, nm), a, b, tol)
            if r:
                msgs.append(r)
    assert not msgs, "\n".join(msgs)
