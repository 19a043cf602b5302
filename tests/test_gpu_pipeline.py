"""Pipelined mode (hr_set_pipeline / hr_step_device): same bits as the serial call sequence."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,pixfmt", [(1280, 720, 0), (1920, 1080, 0), (1920, 1080, 1)])
def test_pipelined_steps_equal_serial(hr, synth, w, h, pixfmt):
    import torch

    clip = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    nfr = 7
    frames = [clip.frame(k) for k in range(nfr)]
    tdt = torch.uint16 if pixfmt else torch.uint8
    dev = [(torch.from_numpy(y).cuda().view(tdt), torch.from_numpy(uv).cuda().view(tdt)) for y, uv in frames]
    ts = [[0.0, 0.4, 0.8], [0.2, 0.6]] * 4

    # serial reference: plain call sequence, host download
    a = hr.HrCuda(h, w, w, pixfmt)
    want, flows = [], []
    a.update_frame(*frames[0])
    for k in range(1, nfr):
        a.update_frame(*frames[k])
        a.calc_flow(5 + (k % 2) * 3, 8, 6)
        flows.append(a.get_offsets()[1].copy())
        for t in ts[k]:
            a.warp(t, 2)
            y, uv, _ = a.download()
            want.append((y.copy(), uv.copy()))
    a.close()

    b = hr.HrCuda(h, w, w, pixfmt)
    b.set_pipeline(True)
    outs = [(torch.zeros((h, w), dtype=tdt, device="cuda"), torch.zeros((h // 2, w), dtype=tdt, device="cuda")) for _ in want]
    oi = 0
    b.step_device(*dev[0], [], [])
    for k in range(1, nfr):
        n = len(ts[k])
        b.step_device(*dev[k], ts[k], outs[oi:oi + n], radius=5 + (k % 2) * 3)
        oi += n
    b.synchronize()
    # the same stream once more, several source frames per call (hr_steps_device; one radius per call)
    c = hr.HrCuda(h, w, w, pixfmt)
    c.set_pipeline(True)
    outs_c = [(torch.zeros((h, w), dtype=tdt, device="cuda"), torch.zeros((h // 2, w), dtype=tdt, device="cuda")) for _ in want]
    c.steps_device([dev[0]], [[]], [])
    oc = 0
    for k in range(1, nfr):     # radius alternates per frame in the serial run: one frame per call here, two calls share a list
        n = len(ts[k])
        c.steps_device([dev[k]], [ts[k]], outs_c[oc:oc + n], radius=5 + (k % 2) * 3)
        oc += n
    c.synchronize()
    for i in range(len(want)):
        assert torch.equal(outs_c[i][0], outs[i][0]) and torch.equal(outs_c[i][1], outs[i][1]), "hr_steps_device output %d" % i
    c.close()
    assert np.array_equal(b.get_offsets()[1], flows[-1])
    for i, (wy, wuv) in enumerate(want):
        gy, guv = outs[i][0].cpu().numpy().view(wy.dtype), outs[i][1].cpu().numpy().view(wuv.dtype)
        assert np.array_equal(gy, wy), "output %d luma differs" % i
        assert np.array_equal(guv, wuv), "output %d chroma differs" % i
    # back to serial mode on the same context: the plain calls still work and agree
    b.set_pipeline(False)
    b.update_frame(*frames[0])
    b.update_frame(*frames[1])
    b.calc_flow(8, 8, 6)
    assert np.array_equal(b.get_offsets()[1], flows[0])
    b.close()


@pytest.mark.parametrize("pixfmt", [0, 1])
def test_download_into_pinned_planes_equals_staged_copy(hr, synth, pixfmt):
    """downloadFrame into pinned and into pageable host planes, repeated downloads of one warp, mode changes in
    between, through the reference-named interface (which runs the context in pipelined mode): same bits."""
    import torch

    w, h = 1920, 1080
    clip = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    npdt = np.uint16 if pixfmt else np.uint8
    tdt = torch.uint16 if pixfmt else torch.uint8
    ofc = hr.OpticalFlowCalc()
    assert not hr.initOpticalFlowCalc(ofc, h, w, w, pixfmt)
    f = [clip.frame(k) for k in range(3)]
    py, puv = torch.zeros((h, w), dtype=tdt).pin_memory(), torch.zeros((h // 2, w), dtype=tdt).pin_memory()
    assert not hr.updateFrame(ofc, list(f[0]))
    for k in (1, 2):
        assert not hr.updateFrame(ofc, list(f[k]))
        assert not hr.calculateOpticalFlow(ofc)
        for t, mode in ((0.0, 2), (0.4, 2), (0.8, 5), (0.6, 0), (0.3, 3)):
            ny, nuv = np.zeros((h, w), npdt), np.zeros((h // 2, w), npdt)
            assert not hr.warpFrames(ofc, t, mode)
            assert not hr.downloadFrame(ofc, [ny, nuv])
            assert not hr.warpFrames(ofc, t, mode)
            assert not hr.downloadFrame(ofc, [py, puv])
            assert ofc.warpCalcTime > 0.0
            assert np.array_equal(py.numpy().view(npdt), ny) and np.array_equal(puv.numpy().view(npdt), nuv), (k, t, mode)
            py.zero_()
            assert not hr.downloadFrame(ofc, [py, puv])            # again, without a new warp
            assert np.array_equal(py.numpy().view(npdt), ny)
            ny2, nuv2 = np.zeros_like(ny), np.zeros_like(nuv)
            assert not hr.downloadFrame(ofc, [ny2, nuv2])
            assert np.array_equal(ny2, ny) and np.array_equal(nuv2, nuv)
    hr.freeOFC(ofc)


def test_c_host_layer_replay_equals_call_by_call(hr, synth):
    """libhopperrender_ofc.so (the compiled drop-in for the reference's opticalFlowCalc.c) driven by hrReplay.c
    delivers the same frames as the six calls made one by one."""
    import ctypes
    import torch

    w, h = 1920, 1080
    clip = synth.MovingTextureClip(w, h)
    frames = [clip.frame(k) for k in range(4)]
    ts = [[0.0, 0.4, 0.8], [0.2, 0.6], [0.0, 0.4, 0.8]]
    ofc = hr.OpticalFlowCalc()
    assert not hr.initOpticalFlowCalc(ofc, h, w, w)
    assert not hr.updateFrame(ofc, list(frames[0]))
    want = None
    for k in range(3):
        assert not hr.updateFrame(ofc, list(frames[k + 1]))
        assert not hr.calculateOpticalFlow(ofc)
        for t in ts[k]:
            y, uv = np.zeros((h, w), np.uint8), np.zeros((h // 2, w), np.uint8)
            assert not hr.warpFrames(ofc, t, 2)
            assert not hr.downloadFrame(ofc, [y, uv])
            want = (y, uv)
    hr.freeOFC(ofc)

    lib = hr.load_ofc_library()
    c = hr.COpticalFlowCalc()
    assert not lib.initOpticalFlowCalc(ctypes.byref(c), h, w, w)
    assert c.isInitialized and c.opticalFlowResScalar == 2 and c.opticalFlowFrameWidth == 480
    oy, ouv = torch.zeros((h, w), dtype=torch.uint8).pin_memory(), torch.zeros((h // 2, w), dtype=torch.uint8).pin_memory()
    assert hr.replay_stream_c(c, frames, 0, [[]], 2, (oy, ouv)) == 0
    assert hr.replay_stream_c(c, frames, 1, ts, 2, (oy, ouv)) == 8
    assert c.ofcCalcTime > 0.0 and c.warpCalcTime > 0.0
    assert np.array_equal(oy.numpy(), want[0]) and np.array_equal(ouv.numpy(), want[1])
    lib.freeOFC(ctypes.byref(c))
    assert not c.isInitialized


def test_work_ahead_never_changes_results(hr, synth):
    """The host interface in pipelined mode starts the next search and the next warp ahead of the calls that ask for
    them. Whatever the caller then asks for — the guessed values or anything else — the frames are those of a
    plain context that guesses nothing."""
    w, h = 1280, 720
    clip = synth.MovingTextureClip(w, h)
    frames = [clip.frame(k) for k in range(6)]
    # (radius, [(t, mode, black, white), ...]) per source frame: regular pacing, a radius change, an irregular t, a
    # repeated t, a mode change and a level change right after a regular run
    script = [
        (5, [(0.0, 2, 0.0, 255.0), (0.4, 2, 0.0, 255.0), (0.8, 2, 0.0, 255.0)]),
        (5, [(0.2, 2, 0.0, 255.0), (0.6, 2, 0.0, 255.0)]),
        (6, [(0.0, 2, 0.0, 255.0), (0.4, 2, 0.0, 255.0), (0.5, 2, 0.0, 255.0), (0.5, 2, 0.0, 255.0)]),
        (6, [(0.1, 2, 0.0, 255.0), (0.3, 2, 0.0, 255.0), (0.5, 0, 0.0, 255.0), (0.7, 2, 16.0, 219.0), (0.9, 2, 16.0, 219.0)]),
        (9, [(0.25, 5, 0.0, 255.0), (0.5, 5, 0.0, 255.0), (0.75, 5, 0.0, 255.0)]),
    ]
    plain = hr.HrCuda(h, w, w)
    plain.update_frame(*frames[0])
    ofc = hr.OpticalFlowCalc()
    assert not hr.initOpticalFlowCalc(ofc, h, w, w)           # pipelined, guessing
    assert not hr.updateFrame(ofc, list(frames[0]))
    for k, (radius, warps) in enumerate(script):
        plain.update_frame(*frames[k + 1])
        plain.calc_flow(radius, 8, 6)
        assert not hr.updateFrame(ofc, list(frames[k + 1]))
        ofc.opticalFlowSearchRadius = radius
        assert not hr.calculateOpticalFlow(ofc)
        if k == 2:                                            # the flow is asked for twice, with other knobs in between
            ofc.deltaScalar = 4
            assert not hr.calculateOpticalFlow(ofc)
            ofc.deltaScalar = 8
            assert not hr.calculateOpticalFlow(ofc)
        assert np.array_equal(ofc.impl.get_offsets()[1], plain.get_offsets()[1]), "flow of pair %d" % k
        for t, mode, black, white in warps:
            plain.warp(t, mode, black, white)
            py, puv, _ = plain.download()
            ofc.outputBlackLevel, ofc.outputWhiteLevel = black, white
            gy, guv = np.zeros((h, w), np.uint8), np.zeros((h // 2, w), np.uint8)
            assert not hr.warpFrames(ofc, t, mode)
            assert not hr.downloadFrame(ofc, [gy, guv])
            assert np.array_equal(gy, py) and np.array_equal(guv, puv), "pair %d t=%g mode %d" % (k, t, mode)
            if t == 0.4:                                      # the same frame once more
                assert not hr.downloadFrame(ofc, [gy, guv])
                assert np.array_equal(gy, py) and np.array_equal(guv, puv)
    hr.freeOFC(ofc)
    plain.close()


def test_many_frames_per_call(hr, synth):
    """hr_steps_device with a run of source frames per call equals one hr_step_device call per frame."""
    import torch

    w, h = 1280, 720
    clip = synth.MovingTextureClip(w, h)
    frames = [clip.frame(k) for k in range(9)]
    dev = [(torch.from_numpy(y).cuda(), torch.from_numpy(uv).cuda()) for y, uv in frames]
    ts = [[]] + [[0.0, 0.4, 0.8], [0.2, 0.6]] * 4
    nout = sum(len(t) for t in ts)
    mk = lambda: [(torch.zeros((h, w), dtype=torch.uint8, device="cuda"), torch.zeros((h // 2, w), dtype=torch.uint8, device="cuda")) for _ in range(nout)]
    a, oa = hr.HrCuda(h, w, w), mk()
    a.set_pipeline(True)
    i = 0
    for k in range(9):
        a.step_device(*dev[k], ts[k], oa[i:i + len(ts[k])], radius=7)
        i += len(ts[k])
    a.synchronize()
    b, ob = hr.HrCuda(h, w, w), mk()
    b.set_pipeline(True)
    b.steps_device(dev[:4], ts[:4], ob[:sum(len(t) for t in ts[:4])], radius=7)
    first = sum(len(t) for t in ts[:4])
    b.steps_device(dev[4:], ts[4:], ob[first:], radius=7)
    b.synchronize()
    assert np.array_equal(a.get_offsets()[1], b.get_offsets()[1])
    for i in range(nout):
        assert torch.equal(oa[i][0], ob[i][0]) and torch.equal(oa[i][1], ob[i][1]), "output %d" % i
    a.close()
    b.close()


@pytest.mark.parametrize("road", ["pinned", "pageable", "pageable-64k"])
@pytest.mark.parametrize("w,h,stride,pixfmt", [(1000, 562, 1024, 0), (854, 480, 896, 0), (1918, 1080, 1920, 1), (3840, 2160, 3840, 0), (7680, 4320, 7680, 1)])
def test_lattice_rows_first_upload(hr, synth, monkeypatch, w, h, stride, pixfmt, road):
    """updateFrame through the host layer uploads the rows the search reads first (pitched copies) and the rest behind
    them while the search runs: ragged heights (a partial last row group), resolution scalars 1 to 4, NV12 and P010, the
    frame in device memory and everything computed from it equal to a plain context's. Pinned planes go row group by
    row group straight to the copy engine; pageable ones are gathered into the pinned ring by the copying threads first
    (csrc/hr_staging.h, with 64 KB chunks every part of a frame goes round the eight slots several times — or, where a row
    group is larger than a chunk, the frame goes up in one piece)."""
    import torch
    if road == "pageable-64k":
        monkeypatch.setenv("HR_STAGE_CHUNK_KB", "64")
    clip = synth.MovingTextureClip(w, h, stride=stride, pixfmt=pixfmt)
    dt = np.uint16 if pixfmt else np.uint8
    plain = hr.HrCuda(h, stride, w, pixfmt)
    ofc = hr.OpticalFlowCalc()
    assert not hr.initOpticalFlowCalc(ofc, h, stride, w, pixfmt)
    lib = hr.load_library()
    for k in range(4):
        f = clip.frame(k)
        plain.update_frame(*f)
        if road == "pinned":
            f = tuple(torch.from_numpy(p.view(np.int16) if pixfmt else p).pin_memory() for p in f)
        assert lib.hr_debug_host_pointer_kind(hr._ptr(f[0])) == (1 if road == "pinned" else 0)
        assert not hr.updateFrame(ofc, list(f))
        if k == 0:
            continue
        plain.calc_flow(5, 8, 6)
        assert not hr.calculateOpticalFlow(ofc)
        a, b = ofc.impl.get_offsets(), plain.get_offsets()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), "flow of pair %d" % k
        for t in (0.3, 0.7):
            plain.warp(t, 2)
            py, puv, _ = plain.download()
            gy, guv = np.zeros((h, stride), dt), np.zeros((h // 2, stride), dt)
            assert not hr.warpFrames(ofc, t, 2)
            assert not hr.downloadFrame(ofc, [gy, guv])
            assert np.array_equal(gy, py) and np.array_equal(guv, puv), "pair %d t=%g" % (k, t)
    hr.freeOFC(ofc)
    plain.close()


@pytest.mark.parametrize("w,h,pixfmt,chunk_kb,threads", [(1920, 1080, 0, 512, 4), (1920, 1080, 0, 64, 3), (1920, 1080, 0, 128, 1), (3840, 2160, 1, 512, 4), (1280, 720, 1, 64, 2), (1920, 1080, 0, 512, 0)])
def test_pageable_planes_through_the_staging_ring(hr, synth, monkeypatch, w, h, pixfmt, chunk_kb, threads):
    """Pageable planes go through the pinned ring of csrc/hr_staging.h (copying threads + copy engine, chunk by chunk;
    small chunks make a frame go round the eight slots several times), pinned planes straight to the copy engine, and
    with HR_STAGE_THREADS=0 pageable planes are left to the driver: the same frames come back on every road, whether
    the two planes are one allocation or two."""
    import torch

    monkeypatch.setenv("HR_STAGE_CHUNK_KB", str(chunk_kb))
    monkeypatch.setenv("HR_STAGE_THREADS", str(threads))
    clip = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    npdt = np.uint16 if pixfmt else np.uint8
    tdt = torch.uint16 if pixfmt else torch.uint8
    frames = [clip.frame(k) for k in range(4)]
    ts = [[0.0, 0.4, 0.8], [0.2, 0.6], [0.0, 0.4, 0.8]]

    def run(kind):
        def host_pair(src=None):
            if kind == "pinned":
                y, uv = torch.zeros((h, w), dtype=tdt).pin_memory(), torch.zeros((h // 2, w), dtype=tdt).pin_memory()
                if src is not None:
                    y.numpy().view(npdt)[:] = src[0]
                    uv.numpy().view(npdt)[:] = src[1]
                return [y, uv]
            if kind == "one_allocation":      # UV plane right behind the Y plane, as mpv's pool lays NV12 out
                both = np.zeros((h + h // 2, w), npdt)
                y, uv = both[:h], both[h:]
            else:
                y, uv = np.zeros((h, w), npdt), np.zeros((h // 2, w), npdt)
            if src is not None:
                y[:] = src[0]
                uv[:] = src[1]
            return [y, uv]

        ofc = hr.OpticalFlowCalc()
        assert not hr.initOpticalFlowCalc(ofc, h, w, w, pixfmt)
        got = []
        src = [host_pair(f) for f in frames]
        assert not hr.updateFrame(ofc, src[0])
        for k in range(3):
            assert not hr.updateFrame(ofc, src[k + 1])
            assert not hr.calculateOpticalFlow(ofc)
            for t in ts[k]:
                out = host_pair()
                assert not hr.warpFrames(ofc, t, 2)
                assert not hr.downloadFrame(ofc, out)
                got.append([np.array(p.numpy().view(npdt) if hasattr(p, "numpy") else p) for p in out])
        hr.freeOFC(ofc)
        return got

    want = run("pinned")
    for kind in ("two_allocations", "one_allocation"):
        got = run(kind)
        for i, (a, b) in enumerate(zip(want, got)):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (kind, i)
