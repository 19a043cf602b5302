"""The reference's filter source WITH this repository's patch series (patches/*.patch, applied by oracle/build_ref.py to a
scratch copy) on top of the CUDA path — the three rows SURVEY.md §8f leaves to the filter host:

  N3  P010 negotiated by the image format (patches/0001): 16-bit frames in, 16-bit frames out,
  N4  the headless control channel (patches/0002): the applet's integer codes from a FIFO named by $HOPPERRENDER_CONTROL
      and from `vf-command ... hr <code>`, with the radius pinned,
  N2  IMGFMT_CUDA in and out (patches/0003): source frames stay where a CUDA decoder left them, outputs are warped into
      device images of a hardware pool; not one byte crosses PCIe (hr_debug_host_transfer_bytes stands still),
  N1  the output pool in page-locked memory (patches/0004): downloadFrame copies straight into the image that travels on.

oracle/filter_host_sim.c stands in for mpv's filter runtime and for the slice of libavutil's hardware-frame API the patch
uses. Every output is compared with direct C-ABI calls at the blend positions of the pacing replay: identical."""
import ctypes as C
import os
import pathlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent
SIM = ROOT / "oracle" / "_ref" / "libhr_filter_sim_patched.so"
IMGFMT_NV12, IMGFMT_P010, IMGFMT_CUDA = 1, 2, 3


@pytest.fixture(scope="module")
def sim():
    if not SIM.exists():
        pytest.skip("oracle/_ref/libhr_filter_sim_patched.so not built (python oracle/build_ref.py where /root/reference exists)")
    L = C.CDLL(str(SIM))
    L.hr_sim_create.restype = C.c_void_p
    L.hr_sim_create.argtypes = [C.c_int, C.c_double]
    L.hr_sim_push_fmt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    L.hr_sim_push_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    L.hr_sim_pop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.hr_sim_pop_image.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                   C.POINTER(C.c_double)]
    L.hr_sim_command_text.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.hr_sim_destroy.argtypes = [C.c_void_p]
    return L


def _cudart():
    for name in ("/usr/local/cuda/lib64/libcudart.so", "libcudart.so.12", "libcudart.so"):
        try:
            rt = C.CDLL(name)
            rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
            return rt
        except OSError:
            continue
    pytest.skip("libcudart not found")


def _pop_all(L, s, w, h, dtype):
    outs = []
    while True:
        y = np.empty((h, w), dtype)
        uv = np.empty((h // 2, w), dtype)
        pts = C.c_double()
        if L.hr_sim_pop(s, C.c_void_p(y.ctypes.data), C.c_void_p(uv.ctypes.data), C.byref(pts), None):
            return outs
        outs.append((y, uv, pts.value))


def test_p010_is_negotiated_by_the_image_format(hr, synth, sim):
    from hopperrender_b200 import pacing
    w, h, fps = 1280, 720, 24.0
    clip = synth.MovingTextureClip(w, h, pixfmt=1)
    s = sim.hr_sim_create(2, 60.0)
    direct = hr.HrCuda(h, w, w, 1)
    pacer = pacing.Pacer(fps, 60.0)
    total = 0
    for k in range(4):
        y, uv = clip.frame(k)
        assert sim.hr_sim_push_fmt(s, C.c_void_p(y.ctypes.data), C.c_void_p(uv.ctypes.data), w, h, IMGFMT_P010, k / fps, fps) >= 0
        outs = _pop_all(sim, s, w, h, np.uint16)
        direct.update_frame(y, uv)
        ts = pacer.next_source_frame()
        if k == 0:
            assert len(outs) == 1                      # the first source frame passes as it is (vf_HopperRender.c:490-495)
            continue
        direct.calc_flow(5)
        assert len(outs) == len(ts)
        for (oy, ouv, _), t in zip(outs, ts):
            direct.warp(t, 2)
            ey, euv, _ = direct.download()
            assert np.array_equal(oy, ey) and np.array_equal(ouv, euv), "source frame %d t=%r" % (k, t)
            total += 1
    assert total == 3 + 2 + 3
    sim.hr_sim_destroy(s)


def test_control_codes_from_a_fifo_and_from_vf_command(hr, synth, sim, tmp_path):
    """Codes of the applet channel (reference vf_HopperRender.c:126-183) without the applet: 916 pins the search radius
    to 16 (added code), 5 selects the HSV flow output (mode 3), `hr 4` through the command hook goes back to the blend."""
    from hopperrender_b200 import pacing
    w, h, fps = 1280, 720, 24.0
    fifo = tmp_path / "hopperrender_control"
    os.mkfifo(fifo)
    os.environ["HOPPERRENDER_CONTROL"] = str(fifo)
    try:
        s = sim.hr_sim_create(2, 60.0)
    finally:
        del os.environ["HOPPERRENDER_CONTROL"]
    wfd = os.open(fifo, os.O_WRONLY | os.O_NONBLOCK)
    clip = synth.MovingTextureClip(w, h)
    direct = hr.HrCuda(h, w, w)
    pacer = pacing.Pacer(fps, 60.0)
    radius, mode = 5, 2
    for k in range(5):
        if k == 2:
            os.write(wfd, b"916\n5\n")                 # before source frame 2 is processed
            radius, mode = 16, 3
        if k == 4:
            sim.hr_sim_command_text(s, b"hr", b"4")    # vf-command <label> hr 4
            mode = 2
        y, uv = clip.frame(k)
        assert sim.hr_sim_push_fmt(s, C.c_void_p(y.ctypes.data), C.c_void_p(uv.ctypes.data), w, h, IMGFMT_NV12, k / fps, fps) >= 0
        outs = _pop_all(sim, s, w, h, np.uint8)
        direct.update_frame(y, uv)
        ts = pacer.next_source_frame()
        if k == 0:
            continue
        direct.calc_flow(radius)
        assert len(outs) == len(ts)
        for (oy, ouv, _), t in zip(outs, ts):
            direct.warp(t, mode)
            ey, euv, _ = direct.download()
            assert np.array_equal(oy, ey) and np.array_equal(ouv, euv), "source frame %d radius %d mode %d" % (k, radius, mode)
    os.close(wfd)
    sim.hr_sim_destroy(s)


@pytest.mark.parametrize("pixfmt", [0, 1])
def test_imgfmt_cuda_in_and_out_without_host_copies(hr, synth, sim, pixfmt):
    import torch
    from hopperrender_b200 import pacing
    w, h, fps = 1280, 720, 24.0                        # 1280 samples: the pitch a CUDA frame pool gives (256-byte aligned) is the width
    tdt = torch.uint16 if pixfmt else torch.uint8
    ndt = np.uint16 if pixfmt else np.uint8
    clip = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    frames = [clip.frame(k) for k in range(5)]
    dev = [(torch.from_numpy(y).cuda().view(tdt), torch.from_numpy(uv).cuda().view(tdt)) for y, uv in frames]   # "decoder output"
    torch.cuda.synchronize()
    direct = hr.HrCuda(h, w, w, pixfmt)
    s = sim.hr_sim_create(2, 60.0)
    before = hr.host_transfer_bytes()
    pacer = pacing.Pacer(fps, 60.0)
    got = []
    for k in range(5):
        dy, duv = dev[k]
        n = sim.hr_sim_push_device(s, C.c_void_p(dy.data_ptr()), C.c_void_p(duv.data_ptr()), w, h, w * (2 if pixfmt else 1),
                                   IMGFMT_P010 if pixfmt else IMGFMT_NV12, k / fps, fps)
        assert n >= 0, "the filter marked itself failed"
        ts = pacer.next_source_frame()
        assert n == (1 if k == 0 else len(ts))
        for i in range(n):
            fmt, sub, p0, p1, pitch, pts = C.c_int(), C.c_int(), C.c_void_p(), C.c_void_p(), C.c_int(), C.c_double()
            assert sim.hr_sim_pop_image(s, C.byref(fmt), C.byref(sub), C.byref(p0), C.byref(p1), C.byref(pitch), C.byref(pts)) == 0
            assert fmt.value == IMGFMT_CUDA and sub.value == (IMGFMT_P010 if pixfmt else IMGFMT_NV12)
            if k == 0:
                assert p0.value == dy.data_ptr()       # the first frame of a stream passes as it is
                continue
            assert p0.value not in (dy.data_ptr(), dev[k - 1][0].data_ptr()), "an output must not overwrite a source of the pair"
            assert pitch.value == w * (2 if pixfmt else 1)
            # look at the device image (a device-to-device copy into a tensor of ours, then to the host for the comparison only)
            oy, ouv = torch.empty((h, w), dtype=tdt, device="cuda"), torch.empty((h // 2, w), dtype=tdt, device="cuda")
            assert _cudart().cudaMemcpy(C.c_void_p(oy.data_ptr()), p0, oy.numel() * oy.element_size(), 3) == 0
            assert _cudart().cudaMemcpy(C.c_void_p(ouv.data_ptr()), p1, ouv.numel() * ouv.element_size(), 3) == 0
            got.append((k, ts[i], oy.cpu().numpy().view(ndt), ouv.cpu().numpy().view(ndt)))
    sim.hr_sim_pop_image(None, None, None, None, None, None, None)          # release the last held image
    after = hr.host_transfer_bytes()
    assert after == before, "the IMGFMT_CUDA path moved %d / %d bytes across PCIe" % (after[0] - before[0], after[1] - before[1])
    assert len(got) == 3 + 2 + 3 + 2
    # the same frames through the host interface
    j = 0
    for k in range(5):
        direct.update_frame(*frames[k])
        if k == 0:
            continue
        direct.calc_flow(5)
        while j < len(got) and got[j][0] == k:
            _, t, oy, ouv = got[j]
            direct.warp(t, 2)
            ey, euv, _ = direct.download()
            assert np.array_equal(oy, ey) and np.array_equal(ouv, euv), "source frame %d t=%r" % (k, t)
            j += 1
    assert j == len(got)
    sim.hr_sim_destroy(s)


@pytest.mark.parametrize("pixfmt,w,h", [(0, 1280, 720), (1, 960, 562)])
def test_output_pool_is_page_locked(hr, synth, sim, pixfmt, w, h):
    """patches/0004: the filter's output images (reference vf_HopperRender.c:385) come from hr_host_alloc through
    mp_image_pool_set_allocator / mp_image_from_buffer, are recycled by the pool, and hold the frames the direct calls
    give (960 x 562 P010: an odd number of chroma rows to lay out behind the luma plane)."""
    from hopperrender_b200 import pacing
    fps = 24.0
    lib = hr.load_library()
    sim.hr_sim_pool_images_allocated.restype = C.c_int
    ndt = np.uint16 if pixfmt else np.uint8
    bps = 2 if pixfmt else 1
    clip = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    s = sim.hr_sim_create(2, 60.0)
    before = sim.hr_sim_pool_images_allocated()
    stride = (w * bps + 63) // 64 * 64 // bps                                  # the pool's strides are 64-byte aligned (video/mp_image.h:35)
    direct = hr.HrCuda(h, stride, w, pixfmt)
    pacer = pacing.Pacer(fps, 60.0)
    checked = 0
    for k in range(6):
        y, uv = clip.frame(k)
        assert sim.hr_sim_push_fmt(s, C.c_void_p(y.ctypes.data), C.c_void_p(uv.ctypes.data), w, h, IMGFMT_P010 if pixfmt else IMGFMT_NV12, k / fps, fps) >= 0
        ys, uvs = np.zeros((h, stride), ndt), np.zeros((h // 2, stride), ndt)
        ys[:, :w], uvs[:, :w] = y, uv
        direct.update_frame(ys, uvs)
        ts = pacer.next_source_frame()
        if k:
            direct.calc_flow(5)
        n = 0
        while True:
            fmt, sub, p0, p1, pitch, pts = C.c_int(), C.c_int(), C.c_void_p(), C.c_void_p(), C.c_int(), C.c_double()
            if sim.hr_sim_pop_image(s, C.byref(fmt), C.byref(sub), C.byref(p0), C.byref(p1), C.byref(pitch), C.byref(pts)):
                break
            n += 1
            if k == 0:
                continue                                                        # the first source frame passes as it is
            assert lib.hr_debug_host_pointer_kind(p0) == 1 and lib.hr_debug_host_pointer_kind(p1) == 1, "output image is not page-locked"
            assert pitch.value == stride * bps and p0.value % 64 == 0 and p1.value % 64 == 0
            oy = np.ctypeslib.as_array(C.cast(p0, C.POINTER(C.c_uint16 if pixfmt else C.c_uint8)), (h, stride))
            ouv = np.ctypeslib.as_array(C.cast(p1, C.POINTER(C.c_uint16 if pixfmt else C.c_uint8)), (h // 2, stride))
            direct.warp(ts[n - 1], 2)
            ey, euv, _ = direct.download()
            assert np.array_equal(oy[:, :w], ey[:, :w]) and np.array_equal(ouv[:, :w], euv[:, :w]), "source frame %d t=%r" % (k, ts[n - 1])
            checked += 1
        if k:
            assert n == len(ts)
    assert checked == 3 + 2 + 3 + 2 + 3
    # thirteen outputs from at most four images: the three of one source frame wait in the harness's queue and one more is
    # held by hr_sim_pop_image while the filter asks for the next
    assert 1 <= sim.hr_sim_pool_images_allocated() - before <= 4
    sim.hr_sim_pop_image(s, None, None, None, None, None, None)                # drop the image the harness still holds
    sim.hr_sim_destroy(s)
