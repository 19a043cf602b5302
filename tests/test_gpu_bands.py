"""Spatial bands: a frame stream split over N GPUs must produce exactly the frames of the single-context
path — the same flow on every rank (each GPU searches only its own lattice tiles; tile totals, edge windows and the flow
travel between the searches over NVLink), every output row bit-identical.

The searches of a group wait for one another on the device, so a group needs one GPU per band: those tests are SKIPPED
with a reason on a box with fewer GPUs (run them under `gpurun --gpus 2` / `--gpus 4`; logs under profiles/). What runs
anywhere: a group of ONE band (the band kernel, the row-range pack, the halo bookkeeping with an empty halo) and the
configuration errors. The two-process CUDA-IPC form is tests/test_gpu_bands_ipc.py."""
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


def _stream_equals_single(hr, synth, w, h, pixfmt, devices, radii, max_radius):
    c = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    single = hr.HrCuda(h, w, w, pixfmt, 0)
    bands = hr.BandGroup(h, w, w, pixfmt, devices, max_radius=max_radius)
    assert bands.rows[0][0] == 0 and bands.rows[-1][1] == h
    for k in range(len(radii) + 1):
        f = c.frame(k)
        single.update_frame(*f)
        bands.update_frame(*f)
        if k == 0:
            continue
        R = radii[k - 1]
        single.calc_flow(R)
        bands.calc_flow(R)
        sraw, sblur = single.get_offsets()
        assert np.abs(sraw).max() > 0
        for r, ctx in enumerate(bands.ctx):        # every GPU ends up with the whole flow
            raw, blur = ctx.get_offsets()
            assert np.array_equal(raw, sraw), "raw offsets differ on rank %d (frame %d, R %d)" % (r, k, R)
            assert np.array_equal(blur, sblur), "blurred offsets differ on rank %d (frame %d, R %d)" % (r, k, R)
        for t, mode in ((0.0, 2), (0.4, 2), (0.8, 0), (0.6, 5), (0.5, 3)):
            single.warp(t, mode)
            sy, suv, _ = single.download()
            bands.warp(t, mode)
            by, buv = bands.download()
            assert np.array_equal(by, sy), "luma differs (frame %d t=%.1f mode %d)" % (k, t, mode)
            assert np.array_equal(buv, suv), "chroma differs (frame %d t=%.1f mode %d)" % (k, t, mode)
    halos = [ctx.band_halo() for ctx in bands.ctx]
    bands.close()
    single.close()
    return halos


@pytest.mark.parametrize("w,h,pixfmt,world,radii,max_radius", [
    (1920, 1080, 0, 2, (5, 5, 16), 16),
    (1920, 1080, 0, 4, (5, 8, 5), 8),
    (3840, 2160, 1, 2, (5, 5), 5),
    (1280, 720, 0, 3, (5, 6, 5), 6),
    (7680, 4320, 1, 2, (5,), 5),
    (7680, 4320, 1, 8, (5, 7), 7),       # one lattice tile row per GPU (the last one: the last two)
    (3840, 2160, 0, 8, (16,), 16),       # halos that span several bands
])
def test_banded_stream_equals_single_context(hr, synth, w, h, pixfmt, world, radii, max_radius):
    nd = _ndev()
    if nd < world:
        pytest.skip("%d bands need %d GPUs, this box has %d (run under gpurun --gpus %d)" % (world, world, nd, world))
    halos = _stream_equals_single(hr, synth, w, h, pixfmt, list(range(world)), radii, max_radius)
    # halo exchange, not a gather: at a small radius a rank holds its band and a few rows of its neighbours only
    if max_radius == 5:
        from hopperrender_b200 import sharding
        for (lo, hi, nbytes), (r0, r1) in zip(halos, sharding.band_rows(h, world)):
            assert r0 - lo <= 36 and hi - r1 <= 36
            assert nbytes > 0


@pytest.mark.parametrize("w,h,pixfmt", [(1920, 1080, 0), (1280, 720, 1)])
def test_a_group_of_one_band_equals_the_plain_path(hr, synth, w, h, pixfmt):
    halos = _stream_equals_single(hr, synth, w, h, pixfmt, [0], (5, 16, 7), 16)
    assert halos[0][:2] == (0, h) and halos[0][2] == 0


def test_band_configuration_errors(hr):
    g = hr.HrCuda(1080, 1920, 1920)
    with pytest.raises(hr.HrError):
        g.band_configure(0, 2, [(0, 544), (544, 1080)])        # 544 is not a whole number of lattice tile rows (128 frame rows)
    with pytest.raises(hr.HrError):
        g.band_configure(0, 2, [(0, 512), (512, 1000)])        # does not cover the frame
    with pytest.raises(hr.HrError):
        g.band_set_max_radius(5)                               # not configured yet
    g.band_configure(0, 2, [(0, 512), (512, 1080)])
    g.band_set_max_radius(5)
    assert g.band_halo()[:2] == (0, 546)
    y = np.zeros((512, 1920), np.uint8)
    uv = np.zeros((256, 1920), np.uint8)
    with pytest.raises(hr.HrError):
        g.band_upload(y, uv)                                   # peer 1 is not connected
    g.close()
    with pytest.raises(ValueError):
        hr.BandGroup(1080, 1920, 1920, 0, (0, 0))              # one GPU per band
