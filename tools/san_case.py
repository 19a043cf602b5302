"""A small pass over every kernel for compute-sanitizer (no torch): python tools/san_case.py"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import synth

for w, h, stride, pf in ((1280, 720, 1280, 0), (854, 480, 896, 0), (1918, 1080, 1920, 1)):
    c = synth.MovingTextureClip(w, h, stride=stride, pixfmt=pf)
    ofc = hr.OpticalFlowCalc()
    assert not hr.initOpticalFlowCalc(ofc, h, stride, w, pf)
    dt = np.uint16 if pf else np.uint8
    oy, ouv = np.zeros((h, stride), dt), np.zeros((h // 2, stride), dt)
    assert not hr.updateFrame(ofc, list(c.frame(0)))
    for k in (1, 2):
        assert not hr.updateFrame(ofc, list(c.frame(k)))
        ofc.opticalFlowSearchRadius = 5 if k == 1 else 16
        assert not hr.calculateOpticalFlow(ofc)
        for t, mode in ((0.0, 2), (0.4, 2), (0.8, 2), (0.5, 0), (0.5, 3), (0.5, 5), (0.5, 6), (1.0, 4)):
            assert not hr.warpFrames(ofc, t, mode)
            assert not hr.downloadFrame(ofc, [oy, ouv])
    # a flow that reaches across the frame: mirrors and clamps everywhere
    rng = np.random.default_rng(3)
    lw, lh = ofc.impl.info.lowWidth, ofc.impl.info.lowHeight
    flow = rng.integers(-512, 393, size=(2, lh, lw)).astype(np.int16)
    ofc.impl.set_blurred_offsets(flow)
    for mode in (0, 1, 2):
        assert not hr.warpFrames(ofc, 0.3, mode)
        assert not hr.downloadFrame(ofc, [oy, ouv])
    hr.freeOFC(ofc)
    print("ok", w, h, stride, pf)
