"""Headless control surface (mpv/video/filter/HopperRender/hrControl.c): the applet's integer codes
(reference video/filter/HopperRender/vf_HopperRender.c:112-180) applied to the same state, without the applet.
Host-only C: runs without a GPU."""
import ctypes as C
import os

import pytest


@pytest.fixture(scope="module")
def ctl(hr):
    return hr.load_ofc_library()


def _fresh(hr):
    ofc = hr.COpticalFlowCalc()
    ofc.outputBlackLevel, ofc.outputWhiteLevel = 0.0, 255.0
    ofc.deltaScalar, ofc.neighborBiasScalar, ofc.opticalFlowSearchRadius = 8, 6, 5
    st = hr.HrControlState()
    st.interpolationActive, st.frameOutputMode = 1, 2
    return ofc, st


def test_codes_map_like_the_applet_channel(hr, ctl):
    ofc, st = _fresh(hr)
    ap = lambda code: ctl.hrControlApply(C.byref(ofc), C.byref(st), code)
    assert ap(0) == 0 and st.interpolationActive == 0 and st.restartCounters == 1          # :127-133
    assert ap(1) == 0 and st.interpolationActive == 1                                       # :134-136
    for code in range(2, 9):                                                                # :138-158, enum FrameOutput :21
        assert ap(code) == 0 and st.frameOutputMode == code - 2
    for code, (b, w) in {9: (0, 255), 10: (10, 219), 11: (16, 219)}.items():                # :159-170
        assert ap(code) == 0 and (ofc.outputBlackLevel, ofc.outputWhiteLevel) == (b, w)
    assert ap(100) == 0 and ofc.outputBlackLevel == 0.0                                     # :173-174
    assert ap(355) == 0 and ofc.outputBlackLevel == 255.0
    assert ap(400) == 0 and ofc.outputWhiteLevel == 0.0                                     # :175-176
    assert ap(619) == 0 and ofc.outputWhiteLevel == 219.0
    assert ap(700) == 0 and ofc.deltaScalar == 0 and ap(731) == 0 and ofc.deltaScalar == 31  # :177-178
    assert ap(806) == 0 and ofc.neighborBiasScalar == 6                                     # :179-180
    # codes of no range leave everything alone, like the reference's default branch
    before = (ofc.outputBlackLevel, ofc.outputWhiteLevel, ofc.deltaScalar, ofc.neighborBiasScalar, st.frameOutputMode)
    for code in (12, 99, 356, 399, 656, 699, 732, 799, 832, 899, 933, -1, 100000):
        assert ap(code) == 1
    assert before == (ofc.outputBlackLevel, ofc.outputWhiteLevel, ofc.deltaScalar, ofc.neighborBiasScalar, st.frameOutputMode)
    # added: pinned radius
    assert ap(916) == 0 and st.pinnedRadius == 16 and ofc.opticalFlowSearchRadius == 16
    assert ap(901) == 1 and st.pinnedRadius == 16
    assert ap(900) == 0 and st.pinnedRadius == 0 and ofc.opticalFlowSearchRadius == 16


def test_parse_rule(ctl):
    assert ctl.hrControlParse(b"4\n") == 4
    assert ctl.hrControlParse(b"619 trailing") == 619
    assert ctl.hrControlParse(b"\n") == -1 and ctl.hrControlParse(b"x4") == -1 and ctl.hrControlParse(b"") == -1 and ctl.hrControlParse(b"-3") == -1


def test_poll_reads_lines_from_a_pipe(hr, ctl):
    ofc, st = _fresh(hr)
    r, w = os.pipe()
    os.set_blocking(r, False)
    assert ctl.hrControlPoll(r, C.byref(ofc), C.byref(st)) == 0            # nothing there
    os.write(w, b"5\n116\n\nnoise\n703\n")
    assert ctl.hrControlPoll(r, C.byref(ofc), C.byref(st)) == 3
    assert st.frameOutputMode == 3 and ofc.outputBlackLevel == 16.0 and ofc.deltaScalar == 3
    os.close(w)
    assert ctl.hrControlPoll(r, C.byref(ofc), C.byref(st)) == 0            # end of file
    os.close(r)
    assert ctl.hrControlPoll(r, C.byref(ofc), C.byref(st)) == -1           # a real read error


def test_poll_keeps_a_line_cut_by_the_end_of_a_read(hr, ctl):
    """A code split over two reads ('455' arriving as '45' and '5\\n') must not be taken for two codes."""
    ofc, st = _fresh(hr)
    r, w = os.pipe()
    os.set_blocking(r, False)
    os.write(w, b"3\n45")
    assert ctl.hrControlPoll(r, C.byref(ofc), C.byref(st)) == 1 and st.frameOutputMode == 1
    assert ofc.outputWhiteLevel == 255.0                                     # '45' is not a code yet
    os.write(w, b"5\n70")
    assert ctl.hrControlPoll(r, C.byref(ofc), C.byref(st)) == 1 and ofc.outputWhiteLevel == 55.0 and st.frameOutputMode == 1
    os.write(w, b"4")
    assert ctl.hrControlPoll(r, C.byref(ofc), C.byref(st)) == 0 and ofc.deltaScalar == 8
    os.close(w)                                                              # end of file finishes the last line
    assert ctl.hrControlPoll(r, C.byref(ofc), C.byref(st)) == 1 and ofc.deltaScalar == 4
    os.close(r)
    # a read that ends exactly on the buffer size boundary and a long run of junk
    ofc, st = _fresh(hr)
    r, w = os.pipe()
    os.set_blocking(r, False)
    os.write(w, b"x" * 700 + b"\n" + b"6\n" * 300 + b"81")
    assert ctl.hrControlPoll(r, C.byref(ofc), C.byref(st)) == 300 and st.frameOutputMode == 4
    os.write(w, b"2\n")
    assert ctl.hrControlPoll(r, C.byref(ofc), C.byref(st)) == 1 and ofc.neighborBiasScalar == 12
    os.close(w)
    os.close(r)


def test_status_text(hr, ctl):
    ofc, st = _fresh(hr)
    ofc.frameWidth, ofc.frameHeight, ofc.opticalFlowResScalar = 1920, 1080, 2
    ofc.ofcCalcTime = 40e-6
    buf = C.create_string_buffer(512)
    n = ctl.hrControlStatus(buf, 512, C.byref(ofc), 1 / 60.0, 1 / 24.0, 1.0, 30e-6)
    text = buf.value.decode()
    assert n == len(text) and text.startswith("Search Radius: 5\nCalc Res: 480x270\nTarget Time: 016.67 ms (60.0 fps)")
    assert "OFC Time: 000.04 ms" in text and "Warp Time: 000.03 ms" in text
