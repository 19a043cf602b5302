"""CPU suite: the C-ABI library loads, exports every symbol include/*.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _declared_functions():
    text = (ROOT / "include" / "hopperrender_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hr_[a-z0-9_]+)\s*\(", text)))


def test_header_and_bindings_agree(hr):
    declared = _declared_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(hr.ABI.keys()) == declared


def test_library_exports_every_declared_symbol(hr):
    lib = hr.load_library()
    for name in _declared_functions():
        assert hasattr(lib, name), name
    assert lib.hr_abi_version() == 1


def test_every_entry_point_cites_the_reference():
    text = (ROOT / "include" / "hopperrender_cuda.h").read_text()
    assert text.count("HR/opticalFlowCalc") >= 10


def test_no_cpu_fallback_without_a_device(hr):
    from conftest import HAVE_GPU
    if HAVE_GPU:
        pytest.skip("a CUDA device is present")
    lib = hr.load_library()
    h = ctypes.c_void_p()
    assert lib.hr_create(ctypes.byref(h), 1080, 1920, 1920, 0, -1) != 0
    assert not h.value
    assert b"no CUDA device" in lib.hr_last_error(None)
    ofc = hr.OpticalFlowCalc()
    assert hr.initOpticalFlowCalc(ofc, 1080, 1920, 1920) is True     # failure, reference convention
    assert ofc.isInitialized is False


def test_product_package_never_touches_the_oracle():
    """The oracle is a checker: nothing under the product package may import, link or call it."""
    pkg = ROOT / "mpv-frame-interpolator_b200"
    for p in pkg.rglob("*"):
        if p.suffix in (".py", ".c", ".h", ".cu", ".cuh"):
            t = p.read_text(errors="ignore")
            assert "oracle" not in t.lower() and "hro_" not in t, p


def test_pacing_matches_appendix_d(hr):
    from hopperrender_b200 import pacing
    s = pacing.schedule(11, 24.0, 60.0)
    assert [len(x) for x in s] == [0, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2]
    assert [round(t, 6) for t in s[1]] == [0.0, 0.4, 0.8] and [round(t, 6) for t in s[2]] == [0.2, 0.6]
    s = pacing.schedule(5, 24.0, 144.0)
    assert [len(x) for x in s] == [0, 6, 1, 6, 6]          # 6*(1/6) accumulates to 0.9999999999999999
    assert s[2][0] < 1.0 and s[2][0] > 0.999999
    s = pacing.schedule(7, 24000.0 / 1001.0, 60.0)
    assert [len(x) for x in s][1:] == [3, 3, 2, 3, 2, 3]


def test_page_locked_allocation_fails_loudly_without_a_device(hr):
    """hr_host_alloc / allocHostPlanes (patches/0004): no device, no page-locked memory — an error / NULL, which the
    patched filter answers with mp_image_alloc; a malloc'd pointer is reported as pageable."""
    from conftest import HAVE_GPU
    if HAVE_GPU:
        pytest.skip("a CUDA device is present")
    lib = hr.load_library()
    p = ctypes.c_void_p(1)
    assert lib.hr_host_alloc(ctypes.byref(p), 1 << 20) != 0 and not p.value
    assert lib.hr_host_alloc(None, 1 << 20) != 0 and lib.hr_host_alloc(ctypes.byref(p), 0) != 0
    assert lib.hr_host_free(None) == 0
    buf = ctypes.create_string_buffer(4096)
    assert lib.hr_debug_host_pointer_kind(ctypes.cast(buf, ctypes.c_void_p)) == 0
    ofc = hr.load_ofc_library()
    assert not ofc.allocHostPlanes(1 << 20)
    ofc.freeHostPlanes(None, None)


def test_patch_series_applies_to_the_reference_filter(tmp_path):
    """patches/*.patch apply in order, without fuzz or offsets, to the reference's vf_HopperRender.c (where the reference
    tree is present: this container, not the GPU box) and the result calls the host-layer functions the series needs."""
    import shutil
    import subprocess
    ref = pathlib.Path("/root/reference/video/filter/HopperRender/vf_HopperRender.c")
    if not ref.exists() or not shutil.which("patch"):
        pytest.skip("reference tree or patch(1) not present")
    work = tmp_path / "video" / "filter" / "HopperRender"
    work.mkdir(parents=True)
    shutil.copy(ref, work / "vf_HopperRender.c")
    series = sorted((ROOT / "patches").glob("*.patch"))
    assert [p.name[:4] for p in series] == ["0001", "0002", "0003", "0004"]
    for p in series:
        r = subprocess.run(["patch", "-p1", "-d", str(tmp_path), "-i", str(p)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "fuzz" not in r.stdout and "offset" not in r.stdout, r.stdout
    text = (work / "vf_HopperRender.c").read_text()
    header = (ROOT / "mpv-frame-interpolator_b200" / "mpv" / "video" / "filter" / "HopperRender" / "opticalFlowCalc.h").read_text()
    for fn in ("updateFrameDevice", "warpFramesToDevice", "finishFrames", "allocHostPlanes", "freeHostPlanes", "hrControlPoll"):
        assert fn + "(" in text or fn + ")" in text or fn + ";" in text, fn
    for fn in ("updateFrameDevice", "warpFramesToDevice", "finishFrames", "allocHostPlanes", "freeHostPlanes"):
        assert fn in header, fn
    assert "mp_image_pool_set_allocator(priv->imagePool" in text and "IMGFMT_P010" in text and "IMGFMT_CUDA" in text
