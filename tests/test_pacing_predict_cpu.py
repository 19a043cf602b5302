"""The predictor of the filter's next blending scalar (csrc/hr_pacing_predict.h) against the filter's own pacing
arithmetic (pacing.Pacer = vf_HopperRender.c:357-375,481 in host doubles): the library warps an output ahead only when
the float it guessed is the float the filter then asks for, bit for bit, so the hit rate is what the work-ahead is
worth. No GPU: the tap runs on the host."""
import numpy as np
import pytest


def _sequence(hr_pkg_mod, src, dst, frames):
    from hopperrender_b200 import pacing
    p = pacing.Pacer(src, dst)
    p.next_source_frame()
    per_frame = [p.next_source_frame() for _ in range(frames)]
    flat = [np.float32(t) for ts in per_frame for t in ts]
    first = []
    for ts in per_frame:
        first += [True] + [False] * (len(ts) - 1)
    return flat, first


@pytest.mark.parametrize("src,dst,min_hit", [(24, 60, 0.99), (24, 144, 0.99), (24, 120, 0.99), (30, 60, 0.99), (25, 60, 0.99), (25, 50, 0.99),
                                             (23.976, 59.94, 0.99), (24000 / 1001, 60, 0.9), (24, 59.951, 0.9), (29.97, 143.998, 0.9), (50, 60, 0.99)])
def test_predictor_hits(hr, src, dst, min_hit):
    flat, first = _sequence(hr, src, dst, 1500)
    got = hr.debug_predict_pacing(flat)
    start = len(flat) // 5                      # the trackers have converged by then
    hits = sum(1 for i in range(start, len(flat)) if got[i][2] and got[i][0].tobytes() == flat[i].tobytes() and got[i][1] == first[i])
    rate = hits / (len(flat) - start)
    assert rate >= min_hit, "%s -> %s: %.3f of the scalars guessed bit for bit" % (src, dst, rate)


def test_predictor_recovers_from_a_seek_and_a_speed_change(hr):
    a, fa = _sequence(hr, 24, 60, 400)
    b, fb = _sequence(hr, 24, 60, 400)          # the filter resets its scalar to 0 on a seek
    c, fc = _sequence(hr, 24 * 1.25, 60, 400)   # playback speed 1.25
    flat, first = a + b + c, fa + fb + fc
    got = hr.debug_predict_pacing(flat)
    for lo, hi in ((300, len(a)), (len(a) + 300, len(a) + len(b)), (len(a) + len(b) + 300, len(flat))):
        hits = sum(1 for i in range(lo, hi) if got[i][2] and got[i][0].tobytes() == flat[i].tobytes() and got[i][1] == first[i])
        assert hits / (hi - lo) >= 0.99, (lo, hi, hits / (hi - lo))


def test_predictor_never_guesses_outside_the_unit_interval(hr):
    rng = np.random.default_rng(3)
    flat = [np.float32(x) for x in rng.random(500)]
    for t, _, have in hr.debug_predict_pacing(flat):
        assert (not have) or (0.0 <= float(t) < 1.0)
