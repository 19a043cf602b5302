"""A small pass over every kernel and every hand-off protocol for compute-sanitizer (no torch):

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/san_case.py [sections]

sections (default: all): host  - the six calls of the host interface, all output modes, three geometries, R = 5 / 16 / 3
                         multi - a lattice with more tiles than SMs (search CTAs own several tiles)
                         s4    - resolution scalar 4 (a narrow 8K strip), NV12 and P010
                         pipe  - pipelined mode: hr_steps_device on device planes (two search lanes, warp streams)
                         bands - a band group of one band (band search kernel, row-range pack; groups of several bands need
                                 one GPU each, see tests/test_gpu_bands.py)
tools/run_sanitizers.sh runs the three tools and keeps their summaries (profiles/r02_sanitizer_*.txt).
"""
import ctypes as C
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

import hr_pkg

hr = hr_pkg.load()
from hopperrender_b200 import pacing, synth

sections = set(sys.argv[1:]) or {"host", "multi", "s4", "pipe", "bands"}
MODES = ((0.0, 2), (0.4, 2), (0.8, 2), (0.5, 0), (0.5, 1), (0.5, 3), (0.5, 5), (0.5, 6), (1.0, 4))


def host_pass(w, h, stride, pf, radii=(5, 16), modes=MODES):
    c = synth.MovingTextureClip(w, h, stride=stride, pixfmt=pf)
    ofc = hr.OpticalFlowCalc()
    assert not hr.initOpticalFlowCalc(ofc, h, stride, w, pf)
    dt = np.uint16 if pf else np.uint8
    oy, ouv = np.zeros((h, stride), dt), np.zeros((h // 2, stride), dt)
    assert not hr.updateFrame(ofc, list(c.frame(0)))
    for k, R in enumerate(radii, 1):
        assert not hr.updateFrame(ofc, list(c.frame(k)))
        ofc.opticalFlowSearchRadius = R
        assert not hr.calculateOpticalFlow(ofc)
        for t, mode in modes:
            assert not hr.warpFrames(ofc, t, mode)
            assert not hr.downloadFrame(ofc, [oy, ouv])
    # a flow that reaches across the frame: mirrors and clamps everywhere
    rng = np.random.default_rng(3)
    lw, lh = ofc.impl.info.lowWidth, ofc.impl.info.lowHeight
    flow = rng.integers(-512, 393, size=(2, lh, lw)).astype(np.int16)
    ofc.impl.set_blurred_offsets(flow)
    for mode in (0, 1, 2, 3, 6):
        assert not hr.warpFrames(ofc, 0.3, mode)
        assert not hr.downloadFrame(ofc, [oy, ouv])
    info = ofc.impl.info
    hr.freeOFC(ofc)
    print("ok host", w, h, stride, pf, radii, "search CTAs", info.searchCtas, flush=True)


if "host" in sections:
    for w, h, stride, pf in ((1280, 720, 1280, 0), (854, 480, 896, 0), (1918, 1080, 1920, 1)):
        host_pass(w, h, stride, pf, radii=(5, 16, 3))
if "multi" in sections:
    host_pass(2560, 1080, 2560, 0, radii=(5, 3), modes=((0.4, 2),))
if "s4" in sections:
    host_pass(256, 4320, 256, 0, radii=(5,))
    host_pass(256, 4320, 256, 1, radii=(16,), modes=((0.4, 2), (0.5, 3), (0.5, 6)))

if "pipe" in sections:
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    w, h = 1280, 720
    c = synth.MovingTextureClip(w, h)
    g = hr.HrCuda(h, w, w, 0)

    def dev(a):
        p = C.c_void_p()
        assert rt.cudaMalloc(C.byref(p), a.nbytes) == 0
        assert rt.cudaMemcpy(p, C.c_void_p(a.ctypes.data), a.nbytes, 1) == 0
        return p.value

    frames = [tuple(dev(p) for p in c.frame(k)) for k in range(6)]
    ts = pacing.schedule(13, 24.0, 60.0)
    nout = sum(len(t) for t in ts)
    outs = [(dev(np.zeros((h, w), np.uint8)), dev(np.zeros((h // 2, w), np.uint8))) for _ in range(nout)]
    g.set_pipeline(True)
    g.steps_device([frames[k % 6] for k in range(13)], ts, outs, radius=5)
    g.synchronize()
    g.steps_device([frames[k % 6] for k in range(13)], ts, outs, radius=16)
    g.synchronize()
    g.set_pipeline(False)
    g.close()
    print("ok pipe", nout, "outputs", flush=True)

if "bands" in sections:
    w, h = 1280, 720
    c = synth.MovingTextureClip(w, h)
    b = hr.BandGroup(h, w, w, 0, (0,))
    for k in range(3):
        b.update_frame(*c.frame(k))
        if k:
            b.calc_flow(5)
            b.warp(0.4, 2)
            b.download()
    b.close()
    print("ok bands", flush=True)
