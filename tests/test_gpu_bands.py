"""Spatial bands: a frame stream split over N contexts must produce exactly the frames
of the single-context path — same flow on every rank, every output row bit-identical.

placement "distinct": one band per GPU (peer access over NVLink); SKIPPED with a reason on a box with fewer GPUs than
bands (run it under `gpurun --gpus 2` / `--gpus 4`; log kept under profiles/). placement "one-gpu": all band contexts
on GPU 0 — says so in its id — which exercises the row bookkeeping and the mailbox kernels in stream order, not
NVLink. The two-process CUDA-IPC form is tests/test_gpu_bands_ipc.py."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev():
    cu = ctypes.CDLL("libcuda.so.1")
    cu.cuInit(0)
    n = ctypes.c_int(0)
    cu.cuDeviceGetCount(ctypes.byref(n))
    return n.value


@pytest.mark.parametrize("placement", ["distinct", "one-gpu"])
@pytest.mark.parametrize("w,h,pixfmt,world", [(1920, 1080, 0, 2), (1920, 1080, 0, 4), (3840, 2160, 1, 2), (1280, 720, 0, 3)])
def test_banded_stream_equals_single_context(hr, synth, w, h, pixfmt, world, placement):
    nd = _ndev()
    if placement == "distinct":
        if nd < world:
            pytest.skip("%d bands on distinct GPUs need %d GPUs, this box has %d (run under gpurun --gpus %d)" % (world, world, nd, world))
        devices = list(range(world))
    else:
        if world > 2 and pixfmt == 0 and h == 1080:
            pytest.skip("one-GPU protocol check: the 2- and 3-band cases are enough")
        devices = [0] * world
    c = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    single = hr.HrCuda(h, w, w, pixfmt, 0)
    bands = hr.BandGroup(h, w, w, pixfmt, devices)
    assert bands.rows[0][0] == 0 and bands.rows[-1][1] == h
    for k in range(4):
        f = c.frame(k)
        single.update_frame(*f)
        bands.update_frame(*f)
        if k == 0:
            continue
        single.calc_flow(5)
        bands.calc_flow(5)
        sraw, sblur = single.get_offsets()
        for ctx in bands.ctx:                      # the replicated search gives the same flow everywhere
            raw, blur = ctx.get_offsets()
            assert np.array_equal(raw, sraw) and np.array_equal(blur, sblur)
        for t, mode in ((0.0, 2), (0.4, 2), (0.8, 0), (0.6, 5)):
            single.warp(t, mode)
            sy, suv, _ = single.download()
            bands.warp(t, mode)
            by, buv = bands.download()
            assert np.array_equal(by, sy), "luma differs (frame %d t=%.1f mode %d)" % (k, t, mode)
            assert np.array_equal(buv, suv), "chroma differs (frame %d t=%.1f mode %d)" % (k, t, mode)
    bands.close()
    single.close()


def test_band_configuration_errors(hr):
    g = hr.HrCuda(1080, 1920, 1920)
    with pytest.raises(hr.HrError):
        g.band_configure(0, 2, [(0, 500), (500, 1080)])        # 500 is not a multiple of 2^(s+1) = 8
    with pytest.raises(hr.HrError):
        g.band_configure(0, 2, [(0, 544), (544, 1000)])        # does not cover the frame
    g.band_configure(0, 2, [(0, 544), (544, 1080)])
    y = np.zeros((544, 1920), np.uint8)
    uv = np.zeros((272, 1920), np.uint8)
    with pytest.raises(hr.HrError):
        g.band_upload(y, uv)                                   # peer 1 is not connected
    g.close()
