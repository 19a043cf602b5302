"""CPU suite: the oracle against an independent numpy restatement and against the analytic
properties of the algorithm (static scene -> zero flow, translation -> constant
offset, tie -> lowest candidate, uint32 wrap on noise)."""
import numpy as np
import pytest

import np_restatement as npr


def _pair(synth, w, h, stride=None, k0=0, pixfmt=0):
    c = synth.MovingTextureClip(w, h, stride=stride, pixfmt=pixfmt)
    return c.frame(k0), c.frame(k0 + 1)


def _oracle_flow(oracle, f1, f2, H, W, aW=None, R=5, dS=8, nS=6, pixfmt=0):
    o = oracle.Oracle(H, W, aW, pixfmt)
    o.update_frame(*f1)
    o.update_frame(*f2)
    o.calc_flow(R, dS, nS)
    return o


@pytest.mark.parametrize("w,h,stride", [(640, 360, 640), (854, 480, 896), (480, 270, 480), (320, 180, 320)])
@pytest.mark.parametrize("R", [5, 8])
def test_oracle_matches_numpy_restatement(oracle, synth, w, h, stride, R):
    f1, f2 = _pair(synth, w, h, stride)
    o = _oracle_flow(oracle, f1, f2, h, stride, w, R)
    raw, blurred = o.get_offsets()
    nraw, nblur, layers = npr.calc_flow(f1[0], f1[1], f2[0], f2[1], R, return_layers=True)
    for step in range(o.steps):
        ws = npr.first_window(o.lw, o.lh) >> (step // 2)
        got = o.get_step_layers(step)[::ws, ::ws]
        assert np.array_equal(got, layers[step][::ws, ::ws]), "winning layers differ at step %d" % step
    assert np.array_equal(raw, nraw)
    assert np.array_equal(blurred, nblur)


def test_oracle_matches_numpy_1080p_r16_scalars(oracle, synth):
    f1, f2 = _pair(synth, 1920, 1080)
    o = _oracle_flow(oracle, f1, f2, 1080, 1920, 1920, 16, 5, 3)
    raw, blurred = o.get_offsets()
    nraw, nblur = npr.calc_flow(f1[0], f1[1], f2[0], f2[1], 16, 5, 3)
    assert np.array_equal(raw, nraw) and np.array_equal(blurred, nblur)


def test_static_scene_gives_zero_flow(oracle, synth):
    f1, _ = _pair(synth, 1280, 720)
    o = _oracle_flow(oracle, f1, f1, 720, 1280)
    raw, blurred = o.get_offsets()
    assert not raw.any() and not blurred.any()


def test_pure_translation_gives_constant_offset(oracle, synth):
    c = synth.MovingTextureClip(1280, 720, velocity=(4, -4), fg_velocity=(4, -4))
    # remove the foreground rectangle's own texture: use the background only
    c.fgY = None
    y1 = np.roll(c.bgY, (0, 0), axis=(0, 1)).astype(np.uint8)
    y2 = np.roll(c.bgY, (-4, 4), axis=(0, 1)).astype(np.uint8)
    uv = np.full((360, 1280), 128, np.uint8)
    o = _oracle_flow(oracle, (y1, uv), (y2, uv), 720, 1280)
    raw, _ = o.get_offsets()
    inner = raw[:, 20:-20, 20:-20]
    # content moved (+4, -4): frame1[p + off] ~ frame2[p] -> off = (-4, +4)
    assert (inner[0] == -4).mean() > 0.97
    assert (inner[1] == 4).mean() > 0.97


def test_constant_frames_tie_goes_to_zero_offset(oracle):
    y = np.full((360, 640), 77, np.uint8)
    uv = np.full((180, 640), 128, np.uint8)
    for R in (5, 6, 16):
        o = _oracle_flow(oracle, (y, uv), (y, uv), 360, 640, R=R)
        raw, _ = o.get_offsets()
        assert not raw.any()
        # with every delta equal, the offset bias |c| decides: the zero candidate z = R/2
        ws = npr.first_window(o.lw, o.lh)
        assert (o.get_step_layers(0)[::ws, ::ws] == R // 2).all()


def test_uint32_wrap_on_noise(oracle, synth):
    # 256x256-point windows of full-range noise: sums exceed 2^32 and must wrap identically
    f1 = synth.noise_frame(1080, 1920, 1)
    f2 = synth.noise_frame(1080, 1920, 2)
    o = _oracle_flow(oracle, f1, f2, 1080, 1920, R=5, dS=12)   # deltaScalar is a runtime knob (0..31)
    raw, blurred = o.get_offsets()
    nraw, nblur, layers = npr.calc_flow(f1[0], f1[1], f2[0], f2[1], 5, 12, return_layers=True)
    y1, y2 = f1[0].astype(np.int64), f2[0].astype(np.int64)
    # analytic lower bound shows the wrap really happens
    approx = np.abs(y1[::4, ::4][:256, :256] - y2[::4, ::4][:256, :256]).sum() << 12
    assert approx > 2 ** 32
    assert np.array_equal(o.get_step_layers(0)[::256, ::256], layers[0][::256, ::256])
    assert np.array_equal(raw, nraw) and np.array_equal(blurred, nblur)


def test_p010_flow_equals_nv12_flow_of_top_bytes(oracle, synth):
    c = synth.MovingTextureClip(640, 360, pixfmt=1)
    f1, f2 = c.frame(3), c.frame(4)
    o16 = _oracle_flow(oracle, f1, f2, 360, 640, pixfmt=1)
    f1b = tuple((a >> 8).astype(np.uint8) for a in f1)
    f2b = tuple((a >> 8).astype(np.uint8) for a in f2)
    o8 = _oracle_flow(oracle, f1b, f2b, 360, 640, pixfmt=0)
    assert np.array_equal(o16.get_offsets()[0], o8.get_offsets()[0])
    assert np.array_equal(o16.get_offsets()[1], o8.get_offsets()[1])


def test_blur_matches_numpy_and_truncates_toward_zero(oracle):
    rng = np.random.default_rng(7)
    raw = rng.integers(-70, 70, size=(2, 45, 61), dtype=np.int16)
    assert np.array_equal(oracle.blur_flow(raw), npr.blur(raw))
    neg = np.full((2, 40, 40), -1, np.int16)
    neg[:, ::3, ::5] = 0
    out = oracle.blur_flow(neg)
    assert out.max() <= 0 and (out == 0).any()      # -63/64 -> 0, never -1 by floor


def test_identity_warp_border_quirk(oracle, synth):
    """Zero flow, t=0, mode 0: the warp's own mirror maps col 0 -> 1 and col W-1 -> W-3
    (warpFrameKernel.cl:10-18)."""
    (y, uv), _ = _pair(synth, 640, 360)
    o = oracle.Oracle(360, 640)
    o.update_frame(y, uv)
    o.update_frame(y, uv)
    o.calc_flow(5)
    assert o.warp(0.0, 0) == 0
    oy, ouv = o.download()
    assert np.array_equal(oy[1:-1, 1:-1], y[1:-1, 1:-1])
    assert np.array_equal(oy[5, 0], y[5, 1]) and np.array_equal(oy[5, 639], y[5, 637])
    assert np.array_equal(oy[0, 7], y[1, 7]) and np.array_equal(oy[359, 7], y[357, 7])
    assert np.array_equal(ouv[1:-1, 2:-2], uv[1:-1, 2:-2])


def test_warp_rejects_t_above_one(oracle):
    o = oracle.Oracle(360, 640)
    assert o.warp(1.0000001, 2) == 1
    assert o.warp(1.0, 2) == 0


def test_default_levels_are_identity_and_presets_are_not(oracle, synth):
    """The literal (IEEE) reading of warpFrameKernel.cl:1-7."""
    (y, uv), (y2, uv2) = _pair(synth, 640, 360)
    o = oracle.Oracle(360, 640, arith=oracle.ARITH_IEEE)
    o.update_frame(y, uv)
    o.update_frame(y2, uv2)
    o.calc_flow(5)
    o.warp(0.4, 2, 0.0, 255.0)
    a = o.download()
    o.warp(0.4, 2, 16.0, 219.0)       # preset, vf_HopperRender.c:163-170
    b = o.download()
    yy = a[0].astype(np.float32)
    exp = np.clip((yy - np.float32(16)) / np.float32(219 - 16) * np.float32(255), 0, 255).astype(np.uint8)
    assert np.array_equal(b[0][:, :640], exp[:, :640])
    assert not np.array_equal(a[0], b[0])


def test_nvidia_opencl_arithmetic_stays_within_two_of_the_ieee_reading(oracle, synth):
    """HRO_ARITH_NVCL (fma contraction + div.full = x * MUFU.RCP(y), oracle/hr_oracle.h) against the
    literal IEEE reading: one LSB from the blend, one from the reciprocal; with default levels the luma
    map v -> trunc(v * rcp(255) * 255) loses one LSB wherever the product falls short of v."""
    assert oracle.have_nvcl(), "tests/golden/mufu_rcp_table.npy is missing"
    (y, uv), (y2, uv2) = _pair(synth, 640, 360)
    outs = {}
    for ar in (oracle.ARITH_IEEE, oracle.ARITH_NVCL):
        o = oracle.Oracle(360, 640, arith=ar)
        o.update_frame(y, uv)
        o.update_frame(y2, uv2)
        o.calc_flow(5)
        for key, args in (("id", (0.4, 2, 0.0, 255.0)), ("preset", (0.6, 2, 16.0, 219.0)), ("m0", (0.4, 0, 0.0, 255.0))):
            o.warp(*args)
            outs[(ar, key)] = o.download()
    for key, tol in (("id", 2), ("preset", 3), ("m0", 0)):
        for pl in (0, 1):
            d = np.abs(outs[(0, key)][pl].astype(int) - outs[(1, key)][pl].astype(int))
            assert d.max() <= tol, (key, pl, int(d.max()))
    # identity levels in NVCL arithmetic: rcp(255) is one ulp short, so most luma values drop by one
    d = outs[(0, "id")][0].astype(int) - outs[(1, "id")][0].astype(int)
    assert (d >= 0).all() and (d == 1).mean() > 0.5


def test_rcp_table_is_within_one_ulp(oracle):
    t = np.load(oracle.RCP_TABLE)
    assert t.dtype == np.float32 and t.size == 1024
    ref = np.float32(1.0) / np.arange(1, 1024, dtype=np.float32)
    ulp = np.abs(t[1:].view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1
    assert t[255] < ref[254]            # the reason the reference's default luma levels are not an identity on NVIDIA devices
