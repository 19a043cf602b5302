"""Parity against the REFERENCE ITSELF: the unmodified reference host + .cl kernels executed on the
B200 through the NVIDIA OpenCL ICD (oracle/_ref, built by oracle/build_ref.py).

This is what pins the oracle: (a) oracle == reference, (b) CUDA path == reference, on the same
inputs. Flow offsets bit-exact. Pixels: OpenCL C lets the device compiler contract a*b+c to an fma
and makes '/' only 2.5-ulp accurate, so the reference's pixels exist per device. The CUDA path and
the oracle's NVCL arithmetic restate what the NVIDIA OpenCL compiler emits for the unmodified kernel
(oracle/hr_oracle.h), so against THIS reference run they must be bit-identical in every mode but
HSV (mode 3: the hue goes through the atan2/fmod built-ins; the CUDA ones turned out bit-identical
to the OpenCL ones on the tested inputs, libm's differ in the last ulp: +-1 on rare pixels allowed).
The oracle's IEEE reading of the source stays within +-2 of it.
"""
import json
import os
import pathlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_opencl
    ok, why = ref_opencl.available()
    if not ok:
        pytest.skip("reference OpenCL run unavailable: " + why)
    return ref_opencl


def _triple(hr, oracle, ref, f1, f2, H, W, aW, R, dS=8, nS=6):
    r = ref.Reference(H, W, aW)
    o = oracle.Oracle(H, W, aW)
    g = hr.HrCuda(H, W, aW)
    for x in (r, o, g):
        x.update_frame(*f1)
        x.update_frame(*f2)
        x.calc_flow(R, dS, nS)
    return r, o, g


def _cmp(name, a, b, tol=0):
    d = np.abs(a.astype(np.int64) - b.astype(np.int64))
    assert d.max() <= tol, "%s: %d of %d differ, max |d| = %d, first at %s" % (name, int((d > tol).sum()), d.size, int(d.max()), np.argwhere(d > tol)[0].tolist())


@pytest.mark.parametrize("w,h,stride,R", [(1920, 1080, 1920, 5), (1920, 1080, 1920, 16), (1280, 720, 1280, 8), (854, 480, 896, 5), (3840, 2160, 3840, 5)])
def test_flow_oracle_and_cuda_equal_reference(hr, oracle, synth, ref, w, h, stride, R):
    c = synth.MovingTextureClip(w, h, stride=stride)
    r, o, g = _triple(hr, oracle, ref, c.frame(2), c.frame(3), h, stride, w, R)
    rraw, rblur = r.get_offsets()
    oraw, oblur = o.get_offsets()
    graw, gblur = g.get_offsets()
    _cmp("oracle raw vs reference", oraw, rraw)
    _cmp("oracle blurred vs reference", oblur, rblur)
    _cmp("cuda raw vs reference", graw, rraw)
    _cmp("cuda blurred vs reference", gblur, rblur)
    # last search step: summed deltas and winning layers at the window representatives (window 2)
    rs, osum = r.get_last_sums(R), o.get_last_sums()
    _cmp("oracle last-step window sums vs reference", osum[:, 0:r.lh - (r.lh % 2):2, ::2], rs[:, 0:r.lh - (r.lh % 2):2, ::2])
    r.close()


def test_flow_scalars_and_noise_equal_reference(hr, oracle, synth, ref):
    f1, f2 = synth.noise_frame(1080, 1920, 21), synth.noise_frame(1080, 1920, 22)
    for R, dS, nS in ((5, 12, 6), (9, 4, 10)):
        r, o, g = _triple(hr, oracle, ref, f1, f2, 1080, 1920, 1920, R, dS, nS)
        for nm, x in (("oracle", o), ("cuda", g)):
            _cmp(nm + " raw vs reference", x.get_offsets()[0], r.get_offsets()[0])
            _cmp(nm + " blurred vs reference", x.get_offsets()[1], r.get_offsets()[1])
        r.close()


@pytest.mark.parametrize("w,h,stride", [(1920, 1080, 1920), (854, 480, 896)])
def test_warp_modes_equal_reference(hr, oracle, synth, ref, w, h, stride):
    c = synth.MovingTextureClip(w, h, stride=stride)
    r, o, g = _triple(hr, oracle, ref, c.frame(2), c.frame(3), h, stride, w, 8)
    for mode in range(7):
        for t, (black, white) in ((0.0, (0.0, 255.0)), (0.4, (0.0, 255.0)), (0.8, (16.0, 219.0))):
            assert not r.warp(t, mode, black, white)
            ry, ruv = r.download()
            o.warp(t, mode, black, white)
            oy, ouv = o.download()
            g.warp(t, mode, black, white)
            gy, guv, _ = g.download()
            tol = 1 if mode == 3 else 0
            for nm, a, b in (("oracle Y", oy, ry), ("oracle UV", ouv, ruv), ("cuda Y", gy, ry), ("cuda UV", guv, ruv)):
                _cmp("mode %d t=%.1f %s vs reference" % (mode, t, nm), a[:, :w], b[:, :w], tol)
                if tol:   # HSV: differences must stay rare
                    d = np.abs(a[:, :w].astype(int) - b[:, :w].astype(int))
                    assert (d > 0).mean() < 0.001, "mode %d t=%.1f %s: %.3f%% of the pixels differ" % (mode, t, nm, 100 * (d > 0).mean())
            # the literal IEEE reading of the kernel source: within +-2 (one LSB from the blend's contraction,
            # one from the reciprocal inside '/'), scaled by the level gain
            if mode in (2, 5, 6):
                i = oracle.Oracle(h, stride, w, arith=oracle.ARITH_IEEE)
                i.update_frame(*c.frame(2))
                i.update_frame(*c.frame(3))
                i.set_blurred_offsets(r.get_offsets()[1])
                i.warp(t, mode, black, white)
                iy, iuv = i.download()
                gain = int(np.ceil(255.0 / (white - black)))
                _cmp("mode %d t=%.1f IEEE oracle Y vs reference" % (mode, t), iy[:, :w], ry[:, :w], 1 + gain)
                _cmp("mode %d t=%.1f IEEE oracle UV vs reference" % (mode, t), iuv[:, :w], ruv[:, :w], 1 + gain)
    r.close()


def test_warp_large_flow_equals_reference(hr, oracle, synth, ref):
    c = synth.MovingTextureClip(1920, 1080)
    r, o, g = _triple(hr, oracle, ref, c.frame(0), c.frame(1), 1080, 1920, 1920, 5)
    rng = np.random.default_rng(5)
    coarse = rng.integers(-512, 393, size=(2, 18, 30))
    flow = np.repeat(np.repeat(coarse, 15, axis=1), 16, axis=2).astype(np.int16)
    for x in (r, o, g):
        x.set_blurred_offsets(flow)
    for mode in (0, 1, 2, 5, 6):
        r.warp(0.6, mode)
        ry, ruv = r.download()
        o.warp(0.6, mode)
        oy, ouv = o.download()
        g.warp(0.6, mode)
        gy, guv, _ = g.download()
        for nm, a, b in (("oracle Y", oy, ry), ("oracle UV", ouv, ruv), ("cuda Y", gy, ry), ("cuda UV", guv, ruv)):
            _cmp("mode %d %s vs reference" % (mode, nm), a, b, 0)
    r.close()


def test_reference_timing_on_b200_is_recorded(hr, synth, ref):
    """Not a parity check: records the reference's own ofcCalcTime / warpCalcTime on this B200
    (the only 'existing GPU kernel to beat') into gpurun_out/ for BASELINE/DESIGN notes."""
    c = synth.MovingTextureClip(1920, 1080)
    r = ref.Reference(1080, 1920, 1920)
    r.update_frame(*c.frame(0))
    flows, warps = [], []
    for k in range(1, 12):
        r.update_frame(*c.frame(k))
        flows.append(r.calc_flow(5))
        r.warp(0.4, 2)
        r.download()
        warps.append(r.s.warpCalcTime)
    out = {"reference_opencl_on_b200": {"ofcCalcTime_ms_median": float(np.median(flows[2:]) * 1e3), "warpCalcTime_ms_median": float(np.median(warps[2:]) * 1e3),
                                         "frame": "1920x1080 NV12", "radius": 5}}
    d = ROOT / "gpurun_out"
    d.mkdir(exist_ok=True)
    (d / "reference_opencl_timing.json").write_text(json.dumps(out))
    r.close()
